"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures.

bit-exact : ternary/binary weight codes, packed formats, STE mask, int8 activation codes and scales
tolerance : forward y        rel <= 1e-4  (int32 accumulation is exact; fp32 dequant epilogue)
            grad_x/W/alpha   rel <= 1e-2  (bf16 tensor-core backward), grad_bias rel <= 1e-4
"""
import hashlib
import math

import numpy as np
import pytest
import torch

from conftest import bits_to_f32
from oracle import onebit_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import onebit_b200
    return onebit_b200


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def rel_err(got, ref):
    return float(np.abs(np.asarray(got, np.float64) - np.asarray(ref, np.float64)).max() /
                 (np.abs(ref).max() + 1e-30))


# ------------------------------------------------------------------ weight quantiser (bit-exact)
@pytest.mark.parametrize("bw", [1, 2])
def test_weight_codes_edges(ob, kat_edges, bw):
    """KAT-1 ties/edges through the dense quantiser kernel: W_hat, STE mask and d/d-alpha."""
    for case in ("", "alpha2_"):
        W = torch.tensor(kat_edges["W2" if case else "W"]).cuda().requires_grad_(True)
        a = torch.tensor(kat_edges["alpha2" if case else "alpha"]).cuda().requires_grad_(True)
        ref = kat_edges[f"{case}bw{bw}"]
        what = ob.quantize_weight(W, a, bw)
        what.backward(torch.tensor(kat_edges["g"]).cuda())
        assert what.detach().cpu().tolist() == ref["W_hat"]
        assert W.grad.cpu().tolist() == ref["grad_W"]
        assert math.isclose(a.grad.item(), ref["grad_alpha"], rel_tol=1e-6, abs_tol=1e-6)


def test_weight_codes_seeded_layers(ob, kat_seeded):
    """KAT-2: packed codes of the seeded reference layers hash to the reference's codes, both layouts."""
    from onebit_b200 import quant as obq
    for key, ref in kat_seeded.items():
        torch.manual_seed(0)
        layer = ob.QuantizedLinear(ref["in"], ref["out"])
        with torch.no_grad():
            layer.alpha.copy_(torch.tensor(float(bits_to_f32(ref["alpha_bits"]))))
        layer = layer.cuda()
        for bw in (1, 2):
            packed, packed_t = layer.packed_weight(bw)
            codes = obq.unpack_codes(packed, 0).cpu().numpy()
            assert sha16(codes) == ref[f"sha_q{bw}"], (key, bw)
            assert np.array_equal(orc.unpack_codes(packed.cpu().numpy(), "i8"), codes)
            assert np.array_equal(orc.unpack_codes(packed_t.cpu().numpy(), "bf16"), codes.T)
        assert layer.packed_weight(2)[0] is layer.packed_weight(2)[0]        # cached while the weights are unchanged


def _fresh_codes(obq, layer, bw):
    packed, _ = obq.pack_weight(layer.weight, layer.alpha, bw)
    return obq.unpack_codes(packed, 0).cpu().numpy()


@pytest.mark.parametrize("update", ["mul_", "data_mul_", "adamw_foreach", "adamw_fused", "sgd_after_backward"])
def test_packed_cache_follows_every_kind_of_weight_update(ob, update):
    """The cached 2-bit codes must be re-quantised after ANY update of the latent weights - including the ones that do not
    bump ``Tensor._version`` (fused AdamW, ``.data`` edits): quant.py:124 re-quantises on every forward."""
    from onebit_b200 import quant as obq
    torch.manual_seed(7)
    layer = ob.QuantizedLinear(256, 128).cuda()
    x = torch.randn(40, 256, device="cuda")
    before = obq.unpack_codes(layer.packed_weight(2)[0], 0).cpu().numpy()
    assert layer.packed_weight(2)[0] is layer.packed_weight(2)[0]
    if update == "mul_":
        with torch.no_grad():
            layer.weight.mul_(-1.0)
    elif update == "data_mul_":
        layer.weight.data.mul_(-1.0)
        obq.invalidate_packed_weights()              # raw .data writes are invisible to autograd: the documented call
    elif update == "sgd_after_backward":
        layer(x, 2).square().mean().backward()
        layer.weight.data.add_(layer.weight.grad, alpha=-50.0)   # manual update, no optimiser, no version bump
    else:
        opt = torch.optim.AdamW(layer.parameters(), lr=0.05, fused=(update == "adamw_fused"))
        for _ in range(2):
            opt.zero_grad()
            layer(x, 2).square().mean().backward()
            opt.step()
    packed_now = layer.packed_weight(2)[0]
    after = obq.unpack_codes(packed_now, 0).cpu().numpy()
    assert np.array_equal(after, _fresh_codes(obq, layer, 2)), update
    assert not np.array_equal(after, before), update
    # and the forward really uses them
    y = layer(x, 2)
    ref = orc.linear_forward(x.cpu().numpy(), layer.weight.detach().cpu().numpy(), layer.alpha.detach().cpu().numpy(),
                             layer.bias.detach().cpu().numpy(), 2, act_bits=8)
    assert rel_err(y.detach().cpu().numpy(), ref) < 1e-4


def test_absmean(ob):
    from onebit_b200 import quant as obq
    W = torch.randn(1024, 256, generator=torch.Generator().manual_seed(3))
    got = obq.weight_absmean(W.cuda()).item()
    assert math.isclose(got, W.abs().double().mean().item(), rel_tol=1e-6)


# ------------------------------------------------------------------ activation quantiser (bit-exact)
@pytest.mark.parametrize("shape", [(3, 37, 128), (4, 249, 256), (513, 1024), (300, 2048), (77, 320), (1, 64)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_act_quant_bit_exact(ob, shape, dtype):
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g)
    x.view(-1, shape[-1])[0] = 0.0                          # all-zero token -> amax clamp
    if x.numel() > shape[-1]:
        x.view(-1, shape[-1])[1, 3] = 60.0                  # outlier token
    x = x.to(dtype)
    q, s = ob.act_quant_int8(x.cuda())
    q_ref, s_ref = orc.act_quant(x.float().numpy())
    assert np.array_equal(q.cpu().numpy(), q_ref)
    assert np.array_equal(s.cpu().numpy(), s_ref)


def test_act_quant_golden(ob, kat_layer, kat_layer_stats):
    q, s = ob.act_quant_int8(torch.from_numpy(kat_layer["x"]).cuda())
    assert np.array_equal(q.cpu().numpy(), kat_layer["act_q"]) and np.array_equal(s.cpu().numpy(), kat_layer["act_s"])
    x = torch.randn(4, 249, 256, generator=torch.Generator().manual_seed(1234))
    q, s = ob.act_quant_int8(x.cuda())
    assert sha16(q.cpu().numpy()) == kat_layer_stats["act"]["sha_q"]
    assert sha16(s.cpu().numpy()) == kat_layer_stats["act"]["sha_s"]


# ------------------------------------------------------------------ forward GEMM
@pytest.mark.parametrize("M,N,K", [(1, 64, 64), (128, 256, 256), (300, 256, 256), (996, 1024, 256),
                                   (996, 256, 1024), (2500, 512, 2048), (777, 192, 320)])
@pytest.mark.parametrize("block_n", [0, 64, 128, 256, 1128, 1256])   # 1xxx = CTA-pair kernel
def test_forward_gemm_exact_integer_dot(ob, M, N, K, block_n):
    """The int8 x ternary contraction is exact: compare against an integer matmul, tight tolerance."""
    from onebit_b200 import _cabi, quant as obq
    g = torch.Generator().manual_seed(M + N + K)
    q = torch.randint(-128, 128, (M, K), generator=g, dtype=torch.int8).cuda()
    scale = (torch.rand(M, generator=g) * 50 + 10).cuda()
    codes = torch.randint(-1, 2, (N, K), generator=g, dtype=torch.int8)
    packed = torch.from_numpy(orc.pack_codes(codes.numpy(), "i8")).cuda()
    alpha = torch.tensor(0.0625).cuda()
    bias = torch.randn(N, generator=g).cuda()
    ref = (q.double() @ codes.cuda().double().t()) * (alpha.double() / scale.double())[:, None] + bias.double()
    _cabi.lib.ob_debug_set(_cabi.DBG_FORCE_BLOCK_N, block_n)
    try:
        y = obq.gemm_fwd(q, scale, packed, alpha, bias, N, torch.float32, _cabi.OB_ALPHA_EFF)
        yb = obq.gemm_fwd(q, scale, packed, alpha, None, N, torch.bfloat16, _cabi.OB_ALPHA_EFF)
        torch.cuda.synchronize()
    finally:
        _cabi.lib.ob_debug_set(_cabi.DBG_FORCE_BLOCK_N, 0)
    assert rel_err(y.cpu().numpy(), ref.cpu().numpy()) < 1e-6
    assert rel_err(yb.float().cpu().numpy(), (ref - bias.double()).cpu().numpy()) < 1e-2


@pytest.mark.parametrize("M", [1, 3, 8, 17, 64])
@pytest.mark.parametrize("N,K", [(256, 256), (1024, 256), (256, 1024), (2048, 2048), (64, 64), (192, 320)])
def test_small_batch_gemv_kernel(ob, M, N, K):
    """M <= 64 token rows take the weight-streaming DP4A kernel (csrc/ob_gemv.cu): exact integer contraction, and the SAME
    bits as the tcgen05 kernel it stands in for (fp32 and bf16 outputs, with and without bias)."""
    from onebit_b200 import _cabi, quant as obq
    g = torch.Generator().manual_seed(7 * M + N + K)
    q = torch.randint(-128, 128, (M, K), generator=g, dtype=torch.int8).cuda()
    scale = (torch.rand(M, generator=g) * 50 + 10).cuda()
    codes = torch.randint(-1, 2, (N, K), generator=g, dtype=torch.int8)
    packed = torch.from_numpy(orc.pack_codes(codes.numpy(), "i8")).cuda()
    alpha = torch.tensor(-0.0625).cuda()                     # raw (negative) alpha: |alpha| + 1e-8 inside
    bias = torch.randn(N, generator=g).cuda()
    a_eff = np.float32(abs(np.float32(-0.0625))) + np.float32(1e-8)
    ref = (q.double() @ codes.cuda().double().t()) * (float(a_eff) / scale.double())[:, None] + bias.double()
    outs = {}
    for mode in (2, 1):                                      # 2: forced DP4A kernel, 1: forced tcgen05 kernel
        _cabi.lib.ob_debug_set(_cabi.DBG_SMALL_M, mode)
        try:
            outs[mode] = (obq.gemm_fwd(q, scale, packed, alpha, bias, N, torch.float32),
                          obq.gemm_fwd(q, scale, packed, alpha, None, N, torch.bfloat16))
            torch.cuda.synchronize()
        finally:
            _cabi.lib.ob_debug_set(_cabi.DBG_SMALL_M, 0)
    assert rel_err(outs[2][0].cpu().numpy(), ref.cpu().numpy()) < 1e-6
    assert torch.equal(outs[2][0], outs[1][0])
    assert torch.equal(outs[2][1], outs[1][1])


# ------------------------------------------------------------------ whole layer vs golden fixtures
def _golden_layer(ob, k):
    layer = ob.QuantizedLinear(128, 192)
    with torch.no_grad():
        layer.weight.copy_(torch.from_numpy(k["W"]))
        layer.alpha.copy_(torch.tensor(float(k["alpha"])))
        layer.bias.copy_(torch.from_numpy(k["bias"]))
    return layer.cuda()


@pytest.mark.parametrize("bw", [1, 2])
def test_layer_against_reference_fixture(ob, kat_layer, bw):
    """Fixture produced by the unmodified reference layer wrapped with the Oracle-B activation quantiser."""
    k = kat_layer
    layer = _golden_layer(ob, k)
    x = torch.from_numpy(k["x"]).cuda().requires_grad_(True)
    y = layer(x, bw)
    y.backward(torch.from_numpy(k["gy"]).cuda())
    tag = f"bw{bw}_B"
    assert rel_err(y.detach().cpu().numpy(), k[f"{tag}_y"]) < 1e-4
    assert rel_err(x.grad.cpu().numpy(), k[f"{tag}_gx"]) < 1e-2
    gw = layer.weight.grad.cpu().numpy()
    assert rel_err(gw, k[f"{tag}_gW"]) < 1e-2
    assert np.array_equal(gw != 0, k[f"{tag}_gW"] != 0)                     # STE mask bit-exact
    assert rel_err(layer.bias.grad.cpu().numpy(), k[f"{tag}_gb"]) < 1e-4
    assert math.isclose(layer.alpha.grad.item(), float(k[f"{tag}_galpha"]), rel_tol=1e-2)
    # against the PURE reference (fp32 activations, Oracle-A) only the int8 activation rounding separates us
    # (the fixture holds an outlier token whose absmax step is 40/127, hence the loose bound)
    assert rel_err(y.detach().cpu().numpy(), k[f"bw{bw}_A_y"]) < 6e-2


def test_layer_fp32_bypass_matches_reference(ob, kat_layer):
    k = kat_layer
    layer = _golden_layer(ob, k)
    x = torch.from_numpy(k["x"]).cuda().requires_grad_(True)
    y = layer(x, 32)
    y.backward(torch.from_numpy(k["gy"]).cuda())
    assert rel_err(y.detach().cpu().numpy(), k["bw32_A_y"]) < 1e-5
    assert rel_err(layer.weight.grad.cpu().numpy(), k["bw32_A_gW"]) < 1e-5
    assert layer.alpha.grad is None


@pytest.mark.parametrize("bw", [1, 2])
def test_layer_stats_kat3(ob, kat_layer_stats, bw):
    """KAT-3 at the model's real shape (256 -> 1024, 996 tokens)."""
    torch.manual_seed(0)
    layer = ob.QuantizedLinear(256, 1024).cuda()
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(4, 249, 256, generator=g).cuda().requires_grad_(True)
    gy = torch.randn(4, 249, 1024, generator=g).cuda()
    y = layer(x, bw)
    y.backward(gy)
    ref = kat_layer_stats[f"bw{bw}_B"]
    assert math.isclose(y.double().sum().item(), ref["sum_y"], rel_tol=2e-3)
    assert math.isclose(y.abs().mean().item(), ref["mean_abs_y"], rel_tol=1e-4)
    assert math.isclose(x.grad.abs().mean().item(), ref["mean_abs_gx"], rel_tol=1e-2)
    assert math.isclose(layer.weight.grad.abs().mean().item(), ref["mean_abs_gW"], rel_tol=1e-2)
    assert math.isclose((layer.weight.grad != 0).float().mean().item(), ref["nnz_frac_gW"], rel_tol=1e-6)
    assert math.isclose(layer.alpha.grad.item(), ref["galpha"], rel_tol=1e-2)
    assert math.isclose(layer.bias.grad.double().sum().item(), ref["sum_gb"], rel_tol=1e-4)


# ------------------------------------------------------------------ layer vs oracle on seeded inputs
@pytest.mark.parametrize("M,K,N", [(996, 256, 256), (996, 256, 1024), (996, 1024, 256), (249, 256, 256),
                                   (4100, 512, 2048), (70, 320, 192)])
@pytest.mark.parametrize("bw", [1, 2])
def test_layer_vs_oracle(ob, M, K, N, bw):
    torch.manual_seed(K + N)
    layer = ob.QuantizedLinear(K, N)
    with torch.no_grad():
        layer.bias.normal_(0, 0.1)
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, K, generator=g)
    gy = torch.randn(M, N, generator=g)
    W, a, b = (t.detach().numpy().copy() for t in (layer.weight, layer.alpha, layer.bias))
    layer = layer.cuda()
    xd = x.cuda().requires_grad_(True)
    y = layer(xd, bw)
    y.backward(gy.cuda())
    y_ref = orc.linear_forward(x.numpy(), W, a, b, bw, 8)
    g_ref = orc.linear_backward(gy.numpy(), x.numpy(), W, a, b, bw, 8)
    assert rel_err(y.detach().cpu().numpy(), y_ref) < 1e-4
    assert rel_err(xd.grad.cpu().numpy(), g_ref["x"]) < 1e-2
    assert rel_err(layer.weight.grad.cpu().numpy(), g_ref["weight"]) < 1e-2
    assert np.array_equal(layer.weight.grad.cpu().numpy() != 0, g_ref["weight"] != 0)
    assert rel_err(layer.bias.grad.cpu().numpy(), g_ref["bias"]) < 1e-4
    # grad_alpha = sum over N*K of grad_W_hat * term is cancellation-prone: its error scales with ||grad_W_hat||_2
    # (bf16 operand rounding, 2^-9 per element), not with its own value -> absolute bound of 1e-2 of that norm
    g_hat_norm = float(np.linalg.norm(g_ref["weight"].astype(np.float64))) * 2 ** 0.5      # ~half the entries are masked
    assert math.isclose(layer.alpha.grad.item(), float(g_ref["alpha"]), rel_tol=1e-2, abs_tol=1e-2 * g_hat_norm)


@pytest.mark.parametrize("M,rows2,K,N", [(25536, None, 256, 1024), (25536, None, 1024, 256),
                                         (76608, 25536 + 7 * 399, 256, 1024), (76608, 2 * 25536 + 399, 256, 256)])
def test_backward_at_bench_token_counts(ob, M, rows2, K, N):
    """Backward GEMMs at the token counts the bench runs (M = 25 536 = 64 x 399 per pass; 76 608 rows when the three co-training
    passes are stacked, with a ragged 2-bit / 1-bit split as the stochastic-precision pass produces): grad_x, grad_W (token
    splits + finaliser), grad_alpha, grad_bias against float64 numpy of the same math (oracle, seconds on the host)."""
    torch.manual_seed(K + N + 1)
    layer = ob.QuantizedLinear(K, N)
    with torch.no_grad():
        layer.bias.normal_(0, 0.1)
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, K, generator=g)
    gy = torch.randn(M, N, generator=g) * (1.0 / 64.0)
    W, a, b = (t.detach().numpy().copy() for t in (layer.weight, layer.alpha, layer.bias))
    layer = layer.cuda()
    xd = x.cuda().requires_grad_(True)
    y = layer(xd, 2) if rows2 is None else layer.forward_grouped(xd, rows2)
    y.backward(gy.cuda())
    groups = [(0, M, 2)] if rows2 is None else [(0, rows2, 2), (rows2, M, 1)]
    gw_ref, ga_ref, gb_ref = np.zeros((N, K)), 0.0, np.zeros(N)
    for r0, r1, bw in groups:
        y_ref = orc.linear_forward(x[r0:r1].numpy(), W, a, b, bw, 8)
        g_ref = orc.linear_backward(gy[r0:r1].numpy(), x[r0:r1].numpy(), W, a, b, bw, 8)
        assert rel_err(y[r0:r1].detach().cpu().numpy(), y_ref) < 1e-4
        assert rel_err(xd.grad[r0:r1].cpu().numpy(), g_ref["x"]) < 1e-2
        gw_ref += g_ref["weight"].astype(np.float64)
        ga_ref += float(g_ref["alpha"])
        gb_ref += g_ref["bias"].astype(np.float64)
    gw = layer.weight.grad.cpu().numpy()
    assert rel_err(gw, gw_ref) < 1e-2
    assert np.array_equal(gw != 0, gw_ref != 0)                      # STE mask
    assert rel_err(layer.bias.grad.cpu().numpy(), gb_ref) < 1e-4
    g_hat_norm = float(np.linalg.norm(gw_ref)) * 2 ** 0.5
    assert math.isclose(layer.alpha.grad.item(), ga_ref, rel_tol=1e-2, abs_tol=1e-2 * g_hat_norm)


@pytest.mark.parametrize("M,N,K", [(300, 256, 256), (4100, 1024, 256), (4100, 256, 1024), (1000, 320, 512), (2049, 576, 320),
                                   (64, 256, 256), (25536, 1024, 256)])
@pytest.mark.parametrize("bw", [1, 2])
def test_grad_w_from_int8_codes_matches_bf16_copy(ob, M, N, K, bw):
    """ob_bwd_dw_q8 (CTA pairs, int8 codes converted to bf16 in shared memory) against ob_bwd_dw on the bf16 copy of q made by the
    prep kernel, and both against float64 of the same bf16 operands: the conversion is exact, so the only difference is the
    fp32 summation order of the token splits."""
    from onebit_b200 import _cabi
    lib, check = _cabi.lib, _cabi.check
    torch.manual_seed(M + N + K)
    layer = ob.QuantizedLinear(K, N).cuda()
    q, s = ob.act_quant_int8(torch.randn(M, K, device="cuda"))
    g = torch.randn(M, N, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    dys = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    qb = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
    colsum = torch.empty(lib.ob_bwd_colsum_blocks(M), N, device="cuda")
    check(lib.ob_bwd_prep(g.data_ptr(), 0, s.data_ptr(), q.data_ptr(), M, N, K, dys.data_ptr(), qb.data_ptr(), colsum.data_ptr(), st))
    assert torch.equal(qb.float(), q.float())
    nbytes = lib.ob_bwd_dw_workspace_bytes(M, N, K)
    ws = torch.empty(nbytes, device="cuda", dtype=torch.uint8)
    res = []
    for fn, src in ((lib.ob_bwd_dw, qb), (lib.ob_bwd_dw_q8, q)):
        gw = torch.full((N, K), float("nan"), device="cuda")
        ga, gb = torch.empty((), device="cuda"), torch.empty(N, device="cuda")
        check(fn(dys.data_ptr(), src.data_ptr(), colsum.data_ptr(), layer.weight.data_ptr(), layer.alpha.data_ptr(), 1, bw, M, N, K,
                 gw.data_ptr(), ga.data_ptr(), gb.data_ptr(), ws.data_ptr(), nbytes, st))
        res.append((gw.cpu().double(), ga.item(), gb.cpu()))
    (gw0, ga0, gb0), (gw1, ga1, gb1) = res
    mask = (layer.weight.double() / (layer.alpha.abs().double() + 1e-8)).abs() <= 1.0
    ref = ((dys.double().t() @ qb.double()) * mask).cpu()
    scale = ref.abs().max().item()
    assert not torch.isnan(gw1).any()
    assert (gw1 - ref).abs().max().item() <= 2e-5 * scale
    assert (gw0 - gw1).abs().max().item() <= 2e-5 * scale
    assert torch.equal(gw0 != 0, gw1 != 0)
    g_hat_norm = float(torch.linalg.norm(ref)) * 2 ** 0.5
    assert math.isclose(ga0, ga1, rel_tol=1e-4, abs_tol=1e-5 * g_hat_norm)
    assert torch.equal(gb0, gb1)


@pytest.mark.parametrize("M,rows2,N,K", [(4100, 1500, 1024, 256), (4100, 4100, 256, 1024), (4100, 0, 256, 256), (2049, 1, 320, 512),
                                         (76608, 25536 + 7 * 399, 256, 256), (1000, 999, 256, 256)])
def test_grad_w_over_two_bitwidth_groups_in_one_launch(ob, M, rows2, N, K):
    """ob_bwd_dw_q8_groups (one GEMM launch whose token splits stop at the 2-bit / 1-bit boundary + one finaliser) against the
    sum of two single-group calls on the row ranges; grad_bias against the column sums of all rows."""
    from onebit_b200 import _cabi
    lib, check = _cabi.lib, _cabi.check
    torch.manual_seed(M + rows2 + N)
    layer = ob.QuantizedLinear(K, N).cuda()
    q, s = ob.act_quant_int8(torch.randn(M, K, device="cuda"))
    g = torch.randn(M, N, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    dys = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    colsum = torch.empty(lib.ob_bwd_colsum_blocks(M), N, device="cuda")
    check(lib.ob_bwd_prep(g.data_ptr(), 0, s.data_ptr(), q.data_ptr(), M, N, K, dys.data_ptr(), None, colsum.data_ptr(), st))
    nbytes = lib.ob_bwd_dw_workspace_bytes(M, N, K)
    ws = torch.empty(nbytes, device="cuda", dtype=torch.uint8)
    gw = torch.full((N, K), float("nan"), device="cuda")
    ga, gb = torch.empty((), device="cuda"), torch.empty(N, device="cuda")
    check(lib.ob_bwd_dw_q8_groups(dys.data_ptr(), q.data_ptr(), colsum.data_ptr(), layer.weight.data_ptr(), layer.alpha.data_ptr(), 1,
                                  rows2, M, N, K, gw.data_ptr(), ga.data_ptr(), gb.data_ptr(), ws.data_ptr(), nbytes, st))
    gw_ref, ga_ref = torch.zeros(N, K, dtype=torch.float64), 0.0
    for r0, r1, bw in ((0, rows2, 2), (rows2, M, 1)):
        if r1 <= r0:
            continue
        gw_g, ga_g = torch.empty(N, K, device="cuda"), torch.empty((), device="cuda")
        nb = lib.ob_bwd_dw_workspace_bytes(r1 - r0, N, K)
        ws_g = torch.empty(nb, device="cuda", dtype=torch.uint8)
        check(lib.ob_bwd_dw_q8(dys.data_ptr() + 2 * r0 * N, q.data_ptr() + r0 * K, None, layer.weight.data_ptr(), layer.alpha.data_ptr(),
                               1, bw, r1 - r0, N, K, gw_g.data_ptr(), ga_g.data_ptr(), None, ws_g.data_ptr(), nb, st))
        gw_ref += gw_g.cpu().double()
        ga_ref += ga_g.item()
    scale = gw_ref.abs().max().item()
    assert not torch.isnan(gw).any()
    assert (gw.cpu().double() - gw_ref).abs().max().item() <= 2e-5 * scale
    assert torch.equal(gw.cpu() != 0, gw_ref != 0)
    assert math.isclose(ga.item(), ga_ref, rel_tol=1e-4, abs_tol=1e-5 * float(torch.linalg.norm(gw_ref)))
    assert rel_err(gb.cpu().numpy(), g.double().sum(0).cpu().numpy()) < 1e-5


def test_size_independent_properties_at_bench_size(ob):
    """BASELINE config sizes (M = 65536 tokens, 2048 x 2048): linearity in the activation scale and a
    checksum identity  sum_n y[m,n] = (q[m,:] . colsum(Q)) * alpha/s[m] + sum(b)  instead of a CPU re-run."""
    from onebit_b200 import quant as obq
    M, K, N = 65536, 2048, 2048
    torch.manual_seed(0)
    layer = ob.QuantizedLinear(K, N).cuda()
    x = torch.randn(M, K, device="cuda")
    y = layer(x, 2)
    y2 = layer(x * 2.0, 2)                                  # absmax quantiser is scale-equivariant
    assert torch.allclose(y2 - layer.bias, 2.0 * (y - layer.bias), rtol=1e-5, atol=1e-5)
    q, s = ob.act_quant_int8(x)
    codes = obq.unpack_codes(layer.packed_weight(2)[0], 0)
    colsum = codes.double().sum(0)                          # [K]
    a_eff = layer.alpha.abs().double() + 1e-8
    chk = (q.double() @ colsum) * (a_eff / s.double()) + layer.bias.double().sum()
    got = y.double().sum(1)
    assert (got - chk).abs().max().item() < 1e-3 * chk.abs().max().item()


@pytest.mark.parametrize("K,N", [(100, 70), (80, 256), (256, 5004 % 1000)])
@pytest.mark.parametrize("bw", [1, 2])
def test_unaligned_feature_counts_are_padded_not_emulated(ob, K, N, bw):
    """Feature counts that are not multiples of 64 run on the same kernels through zero padding (no other code path):
    outputs and every gradient still match the oracle."""
    torch.manual_seed(K + N)
    layer = ob.QuantizedLinear(K, N)
    with torch.no_grad():
        layer.bias.normal_(0, 0.1)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(37, K, generator=g)
    gy = torch.randn(37, N, generator=g)
    W, a, b = (t.detach().numpy().copy() for t in (layer.weight, layer.alpha, layer.bias))
    layer = layer.cuda()
    xd = x.cuda().requires_grad_(True)
    y = layer(xd, bw)
    assert y.shape == (37, N)
    y.backward(gy.cuda())
    y_ref = orc.linear_forward(x.numpy(), W, a, b, bw, 8)
    g_ref = orc.linear_backward(gy.numpy(), x.numpy(), W, a, b, bw, 8)
    assert rel_err(y.detach().cpu().numpy(), y_ref) < 1e-4
    assert rel_err(xd.grad.cpu().numpy(), g_ref["x"]) < 1e-2
    assert layer.weight.grad.shape == (N, K) and rel_err(layer.weight.grad.cpu().numpy(), g_ref["weight"]) < 1e-2
    assert rel_err(layer.bias.grad.cpu().numpy(), g_ref["bias"]) < 1e-4
    g_hat_norm = float(np.linalg.norm(g_ref["weight"].astype(np.float64))) * 2 ** 0.5
    assert math.isclose(layer.alpha.grad.item(), float(g_ref["alpha"]), rel_tol=1e-2, abs_tol=1e-2 * g_hat_norm)


# ------------------------------------------------------------------ fused FFN mid-section
@pytest.mark.parametrize("p", [0.0, 0.1])
@pytest.mark.parametrize("bw", [1, 2])
def test_fused_swish_dropout_quant_matches_unfused(ob, p, bw):
    """lin2(dropout(swish(h))) through the fused kernels == the same chain through separate torch ops + our layer
    (same int8 codes up to expf rounding; outputs and gradients within the layer tolerances)."""
    torch.manual_seed(5)
    layer = ob.QuantizedLinear(1024, 256).cuda()
    with torch.no_grad():
        layer.bias.normal_(0, 0.1)
    g = torch.Generator().manual_seed(6)
    h0 = (torch.randn(3, 111, 1024, generator=g) * 2).cuda()
    gy = torch.randn(3, 111, 256, generator=g).cuda()
    keep = (torch.rand(3, 111, 1024, generator=g) > p).cuda() if p > 0 else None

    def run(fused):
        layer.zero_grad()
        h = h0.clone().requires_grad_(True)
        if fused:
            y = layer.forward_swish_dropout(h, bw, p, True, keep=keep)
        else:
            z = h * torch.sigmoid(h)
            if keep is not None:
                z = z * keep.float() * (1.0 / (1.0 - p))
            y = layer(z, bw)
        y.backward(gy)
        return y.detach(), h.grad.clone(), layer.weight.grad.clone(), layer.alpha.grad.clone(), layer.bias.grad.clone()

    yf, ghf, gwf, gaf, gbf = run(True)
    yu, ghu, gwu, gau, gbu = run(False)
    scale = yu.abs().max().item()
    assert (yf - yu).abs().max().item() < 2e-3 * scale           # a code may flip by one step where expf rounds differently
    assert (yf - yu).abs().mean().item() < 1e-5 * scale
    assert (ghf - ghu).abs().max().item() < 1e-2 * ghu.abs().max().item()
    assert (gwf - gwu).abs().max().item() < 1e-2 * gwu.abs().max().item()
    assert torch.equal(gwf != 0, gwu != 0)
    assert torch.allclose(gbf, gbu, rtol=1e-5, atol=1e-5)
    if p > 0:
        dropped = ~keep
        assert ghf[dropped].abs().max().item() == 0.0              # no gradient through dropped activations


# ------------------------------------------------------------------ LayerNorm in front of the routed projections
@pytest.mark.parametrize("shape", [(4, 249, 256), (25536, 256), (3, 7, 128), (1000, 1024), (5, 512)])
def test_layernorm_matches_torch(ob, shape):
    from onebit_b200.norm import layer_norm
    g = torch.Generator().manual_seed(sum(shape))
    C = shape[-1]
    x0 = (torch.randn(*shape, generator=g) * 3 + 1).cuda()
    w = (torch.randn(C, generator=g) * 0.5 + 1).cuda().requires_grad_(True)
    b = torch.randn(C, generator=g).cuda().requires_grad_(True)
    gy = torch.randn(*shape, generator=g).cuda()
    outs = []
    for fn in (lambda t: layer_norm(t, w, b, 1e-5), lambda t: torch.nn.functional.layer_norm(t, (C,), w, b, 1e-5)):
        w.grad = b.grad = None
        x = x0.clone().requires_grad_(True)
        y = fn(x)
        y.backward(gy)
        outs.append((y.detach(), x.grad, w.grad.clone(), b.grad.clone()))
    (y1, gx1, gw1, gb1), (y2, gx2, gw2, gb2) = outs
    assert torch.allclose(y1, y2, rtol=1e-5, atol=1e-5)
    assert torch.allclose(gx1, gx2, rtol=1e-4, atol=1e-5)
    rows = x0.numel() // C
    assert (gw1 - gw2).abs().max().item() < 1e-5 * rows ** 0.5 * gw2.abs().max().clamp_min(1).item()
    assert (gb1 - gb2).abs().max().item() < 1e-5 * rows ** 0.5 * gb2.abs().max().clamp_min(1).item()


# ------------------------------------------------------------------ fused relative-position attention chain
@pytest.mark.parametrize("B,H,T", [(2, 4, 49), (3, 2, 249), (1, 4, 399), (2, 1, 33), (1, 2, 700)])
@pytest.mark.parametrize("p", [0.0, 0.1])
def test_rel_attention_probs_matches_reference_chain(ob, B, H, T, p):
    """Against the reference's own op sequence (conformer.py:96-128) written with torch ops, incl. padded utterances
    (fully masked rows -> zeros) and the relative shift."""
    import math as _m
    from onebit_b200.attention import rel_attention_probs
    from onebit_b200.conformer import MHSA
    g = torch.Generator().manual_seed(B * 1000 + T)
    ac0 = (torch.randn(B, H, T, T, generator=g) * 3).cuda()
    bd0 = (torch.randn(B, H, T, T, generator=g) * 3).cuda()
    lens = torch.randint(1, T + 1, (B,), generator=g)
    lens[0] = T
    km = (torch.arange(T)[None, :] < lens[:, None]).cuda()
    mask = km[:, :, None] & km[:, None, :]
    keep = (torch.rand(B, H, T, T, generator=g) > p).cuda() if p > 0 else None
    gy = torch.randn(B, H, T, T, generator=g).cuda()
    scale = 1.0 / _m.sqrt(64)

    def reference(ac, bd):
        scores = (ac + MHSA.rel_shift(bd)) / _m.sqrt(64)
        scores = scores.masked_fill(mask[:, None, :, :] == 0, float("-inf"))
        a = torch.nan_to_num(torch.softmax(scores, dim=-1), nan=0.0)
        return a if keep is None else a * keep.float() * (1.0 / (1.0 - p))

    outs = []
    for fn in (lambda a, b: rel_attention_probs(a, b, mask, scale, p, True, keep=keep), reference):
        ac, bd = ac0.clone().requires_grad_(True), bd0.clone().requires_grad_(True)
        out = fn(ac, bd)
        out.backward(gy)
        outs.append((out.detach(), ac.grad, bd.grad))
    (o1, ga1, gb1), (o2, ga2, gb2) = outs
    assert torch.allclose(o1, o2, rtol=1e-5, atol=1e-6)
    assert torch.allclose(ga1, ga2, rtol=1e-4, atol=1e-6)
    assert torch.allclose(gb1, gb2, rtol=1e-4, atol=1e-6)
    if (~km).any():
        assert o1.transpose(1, 2)[~km].abs().max().item() == 0.0        # padded query rows are exactly zero


# ------------------------------------------------------------------ in-kernel dropout streams (Philox4x32-10)
@pytest.mark.parametrize("M,K", [(37, 256), (333, 1024), (64, 2048), (5, 512)])
def test_swish_dropout_rng_path_equals_explicit_mask(ob, M, K):
    """The mask the kernels generate from (seed, offset, threshold) is the oracle's Philox mask: forward codes/scales and
    backward gradients are bit-identical to the explicit-mask path fed with the oracle mask."""
    from onebit_b200._cabi import lib, check
    seed, offset, thr = 0x1234567890ABCDEF, 4 * M + (1 << 33), int(round(0.1 * 2 ** 16))
    keep = torch.from_numpy(orc.dropout_keep_flat(M * K, seed, offset, thr).reshape(M, K)).cuda()
    g = torch.Generator().manual_seed(M + K)
    h = (torch.randn(M, K, generator=g) * 2).cuda()
    gz = torch.randn(M, K, generator=g).cuda()
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    for kp, rng in ((keep, (0, 0, 0)), (None, (seed, offset, thr))):
        q = torch.empty(M, K, dtype=torch.int8, device="cuda")
        s = torch.empty(M, device="cuda")
        gh = torch.empty(M, K, device="cuda")
        kptr = None if kp is None else kp.data_ptr()
        check(lib.ob_swish_drop_quant(h.data_ptr(), kptr, 1 / 0.9, *rng, M, K, q.data_ptr(), s.data_ptr(), st))
        check(lib.ob_swish_drop_bwd(gz.data_ptr(), h.data_ptr(), kptr, 1 / 0.9, *rng, M * K, gh.data_ptr(), st))
        outs.append((q, s, gh))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert 0.85 < keep.float().mean().item() < 0.95
    assert torch.equal(outs[1][2] == 0, ~keep | (gz == 0))


@pytest.mark.parametrize("B,H,T", [(2, 3, 49), (1, 2, 399), (1, 1, 700), (1, 1, 1100)])
def test_rel_attention_rng_path_equals_explicit_mask(ob, B, H, T):
    from onebit_b200._cabi import lib, check
    seed, offset, thr = 99, 12 + (7 << 32), int(round(0.25 * 2 ** 16))
    keep = torch.from_numpy(orc.dropout_keep_relattn(B, H, T, seed, offset, thr)).cuda()
    g = torch.Generator().manual_seed(T)
    ac, bd, gd = ((torch.randn(B, H, T, T, generator=g) * 2).cuda() for _ in range(3))
    mask = torch.ones(B, T, T, dtype=torch.bool, device="cuda")
    mask[0, :, T - 3:] = False
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    for kp, rng in ((keep, (0, 0, 0)), (None, (seed, offset, thr))):
        y, ad, d_ac, d_bd = (torch.empty_like(ac) for _ in range(4))
        kptr = None if kp is None else kp.data_ptr()
        check(lib.ob_relattn_softmax_fwd(ac.data_ptr(), bd.data_ptr(), mask.data_ptr(), kptr, 4 / 3, *rng, 0.125, B, H, T, T,
                                         y.data_ptr(), ad.data_ptr(), st))
        check(lib.ob_relattn_softmax_bwd(gd.data_ptr(), y.data_ptr(), kptr, 4 / 3, *rng, 0.125, B, H, T, T, d_ac.data_ptr(),
                                         d_bd.data_ptr(), st))
        outs.append((y, ad, d_ac, d_bd))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert 0.70 < keep.float().mean().item() < 0.80


@pytest.mark.parametrize("M,C,with_mask", [(37, 256, True), (1000, 64, False), (5, 8, True)])
def test_residual_dropout_tail(ob, M, C, with_mask):
    """x + scale * dropout(y) * rowmask: the in-kernel Philox mask equals the oracle's, forward and backward."""
    from onebit_b200._cabi import lib, check
    seed, offset, thr = 77, 40, int(round(0.1 * 2 ** 16))
    inv_keep = 65536.0 / (65536 - thr)
    keep = torch.from_numpy(orc.dropout_keep_groups8(M * C, seed, offset, thr).reshape(M, C)).cuda().float()
    g = torch.Generator().manual_seed(M + C)
    x, y, gr = (torch.randn(M, C, generator=g).cuda() for _ in range(3))
    rm = (torch.rand(M, generator=g) > 0.3).float().cuda() if with_mask else None
    rmv = rm[:, None] if with_mask else 1.0
    out, gy = torch.empty_like(x), torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    rp = None if rm is None else rm.data_ptr()
    check(lib.ob_residual_dropout_fwd(x.data_ptr(), y.data_ptr(), rp, 0.5, inv_keep, seed, offset, thr, M, C, out.data_ptr(), st))
    check(lib.ob_residual_dropout_bwd(gr.data_ptr(), rp, 0.5, inv_keep, seed, offset, thr, M, C, gy.data_ptr(), st))
    assert torch.allclose(out, x + 0.5 * inv_keep * keep * y * rmv, rtol=1e-6, atol=1e-6)
    assert torch.allclose(gy, 0.5 * inv_keep * keep * gr * rmv, rtol=1e-6, atol=1e-6)
    # no dropout: plain residual
    check(lib.ob_residual_dropout_fwd(x.data_ptr(), y.data_ptr(), rp, 1.0, 1.0, 0, 0, 0, M, C, out.data_ptr(), st))
    assert torch.allclose(out, x + y * rmv, rtol=1e-6, atol=1e-6)
    # autograd wrapper
    from onebit_b200.residual import residual_dropout
    xa, ya = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    fm = None if rm is None else rm.bool().view(M, 1)
    o = residual_dropout(xa, ya, fm, 0.5, 0.0, True)
    o.backward(gr)
    assert torch.allclose(o, x + 0.5 * y * rmv, atol=1e-6) and torch.equal(xa.grad, gr)
    assert torch.allclose(ya.grad, 0.5 * gr * rmv, atol=1e-6)


def test_layer_dropout_is_seeded_and_advances(ob):
    """forward_swish_dropout / rel_attention_probs draw their stream from the device generator: same seed -> same result,
    consecutive calls -> different masks, and the backward drops exactly what the forward dropped."""
    from onebit_b200.attention import rel_attention_probs
    layer = ob.QuantizedLinear(256, 256).cuda()
    h = torch.randn(4, 50, 256, device="cuda")
    gy = torch.randn(4, 50, 256, device="cuda")
    gen = torch.cuda.default_generators[torch.cuda.current_device()]

    def run():
        hh = h.clone().requires_grad_(True)
        y = layer.forward_swish_dropout(hh, 2, 0.1, True)
        y.backward(gy)
        return y.detach(), hh.grad

    torch.manual_seed(1234)
    off0 = gen.get_offset()
    y1, g1 = run()
    assert gen.get_offset() == off0 + 4
    y2, g2 = run()
    torch.manual_seed(1234)
    y3, g3 = run()
    assert torch.equal(y1, y3) and torch.equal(g1, g3)
    assert not torch.equal(y1, y2)
    seed = (1234 ^ ob.quant._DROP_DOMAIN) & (2 ** 64 - 1)
    keep = torch.from_numpy(orc.dropout_keep_flat(h.numel(), seed, off0, int(round(0.1 * 2 ** 16)))).cuda().view_as(h)
    assert (g1[~keep] == 0).all() and (g1[keep] != 0).float().mean().item() > 0.99
    # eval mode / p = 0: no stream is drawn
    off = gen.get_offset()
    layer.forward_swish_dropout(h, 2, 0.1, False)
    layer.forward_swish_dropout(h, 2, 0.0, True)
    assert gen.get_offset() == off
    ac, bd = torch.randn(2, 2, 40, 40, device="cuda"), torch.randn(2, 2, 40, 40, device="cuda")
    mask = torch.ones(2, 40, 40, dtype=torch.bool, device="cuda")
    torch.manual_seed(5)
    a1 = rel_attention_probs(ac, bd, mask, 0.125, 0.1, True)
    a2 = rel_attention_probs(ac, bd, mask, 0.125, 0.1, True)
    torch.manual_seed(5)
    a3 = rel_attention_probs(ac, bd, mask, 0.125, 0.1, True)
    assert torch.equal(a1, a3) and not torch.equal(a1, a2)
    frac = (a1 == 0).float().mean().item()
    assert 0.07 < frac < 0.13
    assert torch.allclose(rel_attention_probs(ac, bd, mask, 0.125, 0.1, False).sum(-1), torch.ones(2, 2, 40, device="cuda"),
                          atol=1e-5)


# ------------------------------------------------------------------ edge cases of the layer interface
def test_layer_edge_cases(ob):
    torch.manual_seed(2)
    layer = ob.QuantizedLinear(128, 64).cuda()
    # empty batch
    y = layer(torch.empty(0, 7, 128, device="cuda"), 2)
    assert y.shape == (0, 7, 64)
    # a single token, 1-D input
    x1 = torch.randn(128, device="cuda")
    y1 = layer(x1, 1)
    assert y1.shape == (64,)
    ref = orc.linear_forward(x1.cpu().numpy()[None], layer.weight.detach().cpu().numpy(), layer.alpha.item(),
                             layer.bias.detach().cpu().numpy(), 1, 8)[0]
    assert rel_err(y1.detach().cpu().numpy(), ref) < 1e-4
    # non-contiguous input (a transposed view) and a bf16 input
    xt = torch.randn(128, 33, device="cuda").t()
    assert torch.equal(layer(xt, 2).detach(), layer(xt.contiguous(), 2).detach())
    xb = xt.contiguous().to(torch.bfloat16)
    yb = layer(xb, 2).detach()
    assert yb.dtype == torch.bfloat16
    assert rel_err(yb.float().cpu().numpy(), layer(xb.float(), 2).detach().cpu().numpy()) < 2e-2
    # wrong feature size, wrong bitwidth
    with pytest.raises(ValueError):
        layer(torch.randn(4, 64, device="cuda"), 2)
    with pytest.raises(ValueError, match="bitwidth must be one of"):
        layer(torch.randn(4, 128, device="cuda"), 8)
    # no-bias layer, gradient only w.r.t. the weight
    nb = ob.QuantizedLinear(128, 64, bias=False).cuda()
    x = torch.randn(10, 128, device="cuda")
    nb(x, 2).sum().backward()
    assert nb.weight.grad is not None and nb.alpha.grad is not None and nb.bias is None
    # all-zero input: scale clamp path, output == bias
    z = layer(torch.zeros(3, 128, device="cuda"), 2).detach()
    assert torch.allclose(z, layer.bias.detach().expand(3, 64))
