"""Fusions around the layer inside the routed modules (conformer.py:34-45, 105-138), through the
C ABI, against the unfused kernels they replace and against torch / the oracle.

bit-exact : LayerNorm+quantiser codes and scales == quantiser(LayerNorm); GEMM with the tail epilogue == GEMM then
            residual_dropout kernel (same Philox lanes, oracle-checked mask); fused module forward == unfused module forward
tolerance : fused backward prep vs the separate backward kernels: bf16 operands may differ by one rounding step on a
            < 1e-3 fraction of elements (expression contraction), column sums 1e-5; module gradients 1e-2 (bf16 backward bound)
"""
import numpy as np
import pytest
import torch

from oracle import onebit_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import onebit_b200
    return onebit_b200


def _st():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("M,C", [(1, 128), (37, 256), (996, 256), (513, 512), (300, 1024)])
def test_layernorm_quant_equals_layernorm_then_quant(ob, M, C):
    from onebit_b200 import _cabi, fused
    from onebit_b200.norm import layer_norm
    g = torch.Generator().manual_seed(M + C)
    x = (torch.randn(M, C, generator=g) * 3 + 0.5).cuda()
    x[0] = 0.0                                                   # constant row: normalised values are exactly beta
    w, b = (torch.randn(C, generator=g) * 0.5 + 1).cuda(), (torch.randn(C, generator=g) * 0.1).cuda()
    q, s, stats = fused._ln_quant(x, w, b, 1e-5)
    y = layer_norm(x, w, b, 1e-5)
    q_ref, s_ref = ob.act_quant_int8(y)
    assert torch.equal(q, q_ref) and torch.equal(s, s_ref)
    # and both equal the oracle quantiser applied to the library LayerNorm, which equals torch's to fp32 rounding
    q_o, s_o = orc.act_quant(y.cpu().numpy())
    assert np.array_equal(q.cpu().numpy(), q_o) and np.array_equal(s.cpu().numpy(), s_o)
    y_t = torch.nn.functional.layer_norm(x, (C,), w, b, 1e-5)
    assert (y - y_t).abs().max().item() < 2e-5
    mean, rstd = x.double().mean(1), 1.0 / torch.sqrt(x.double().var(1, unbiased=False) + 1e-5)
    assert torch.allclose(stats[0].double(), mean, atol=1e-5) and torch.allclose(stats[1].double(), rstd, rtol=1e-5)
    del _cabi


@pytest.mark.parametrize("p", [0.0, 0.1])
@pytest.mark.parametrize("M,N,K,row_base", [(300, 256, 256, 0), (996, 256, 1024, 0), (513, 256, 256, 1400), (70, 192, 320, 64),
                                            (4100, 320, 512, 0), (5000, 256, 1024, 100)])
@pytest.mark.parametrize("tile", [0, 1256])     # 0: the launcher's choice; 1256: the 256-wide CTA-pair tile, whose tail takes the residual by TMA
def test_gemm_tail_epilogue_equals_gemm_then_tail_kernel(ob, p, M, N, K, row_base, tile):
    """out = x + scale * dropout(layer(q)) * frame_mask from the GEMM epilogue: same bits as the GEMM followed by the
    residual_dropout kernel on the rows [row_base, row_base + M) of a taller tensor, and the mask is the oracle's."""
    from onebit_b200 import _cabi, quant as obq
    lib = _cabi.lib
    lib.ob_debug_set(_cabi.DBG_FORCE_BLOCK_N, tile)
    try:
        _tail_epilogue_case(ob, p, M, N, K, row_base)
    finally:
        lib.ob_debug_set(_cabi.DBG_FORCE_BLOCK_N, 0)


def _tail_epilogue_case(ob, p, M, N, K, row_base):
    from onebit_b200 import _cabi, quant as obq
    lib = _cabi.lib
    g = torch.Generator().manual_seed(M + N + K)
    total = row_base + M
    q = torch.randint(-128, 128, (M, K), generator=g, dtype=torch.int8).cuda()
    scale = (torch.rand(M, generator=g) * 50 + 10).cuda()
    codes = torch.randint(-1, 2, (N, K), generator=g, dtype=torch.int8)
    packed = torch.from_numpy(orc.pack_codes(codes.numpy(), "i8")).cuda()
    alpha, bias = torch.tensor(0.05).cuda(), torch.randn(N, generator=g).cuda()
    resid_all = torch.randn(total, N, generator=g).cuda()
    rowmask = (torch.rand(total, generator=g) > 0.2).float().cuda()
    thr = int(round(p * 65536))
    inv_keep = 65536.0 / (65536 - thr)
    seed, offset, sc = 0x1234ABCD5678, 40, 0.5
    factor = float(np.float32(sc) * np.float32(inv_keep))
    out = torch.empty(M, N, device="cuda")
    _cabi.check(lib.ob_gemm_tern_i8_fwd_tail(q.data_ptr(), scale.data_ptr(), packed.data_ptr(), alpha.data_ptr(), _cabi.OB_ALPHA_RAW,
                                             bias.data_ptr(), M, N, K, resid_all[row_base:].data_ptr(), rowmask.data_ptr(), factor, seed,
                                             offset, thr, row_base, out.data_ptr(), _st()))
    y = obq.gemm_fwd(q, scale, packed, alpha, bias, N, torch.float32)
    y_all = torch.zeros(total, N, device="cuda")
    y_all[row_base:] = y
    ref_all = torch.empty(total, N, device="cuda")
    _cabi.check(lib.ob_residual_dropout_fwd(resid_all.data_ptr(), y_all.data_ptr(), rowmask.data_ptr(), sc, inv_keep, seed, offset, thr,
                                            total, N, ref_all.data_ptr(), _st()))
    assert torch.equal(out, ref_all[row_base:])
    keep = torch.from_numpy(orc.dropout_keep_groups8(total * N, seed, offset, thr).reshape(total, N)).cuda() if thr else 1.0
    want = resid_all + y_all * keep * rowmask[:, None] * factor
    assert torch.allclose(out, want[row_base:], rtol=1e-6, atol=1e-6)


def _bf16_close(a, b, frac=1e-3):
    """bf16 tensors equal up to one rounding step on at most ``frac`` of the elements."""
    a, b = a.float(), b.float()
    diff = (a - b).abs()
    assert (diff <= 2.0 ** -7 * b.abs() + 1e-30).all()
    assert (diff > 0).float().mean().item() <= frac


@pytest.mark.parametrize("p", [0.0, 0.1])
def test_fused_prep_tail_equals_tail_backward_then_prep(ob, p):
    from onebit_b200 import _cabi
    lib = _cabi.lib
    M, N, K, row_base = 777, 256, 1024, 96
    g = torch.Generator().manual_seed(3)
    total = row_base + M
    gout_all = torch.randn(total, N, generator=g).cuda()
    rowmask = (torch.rand(total, generator=g) > 0.2).float().cuda()
    scale = (torch.rand(M, generator=g) * 50 + 10).cuda()
    q = torch.randint(-128, 128, (M, K), generator=g, dtype=torch.int8).cuda()
    thr = int(round(p * 65536))
    inv_keep, seed, offset, sc = 65536.0 / (65536 - thr), 99, 8, 0.5
    factor = float(np.float32(sc) * np.float32(inv_keep))
    nblk = lib.ob_bwd_colsum_blocks(M)
    dys, qb, cs = (torch.empty(M, N, device="cuda", dtype=torch.bfloat16), torch.empty(M, K, device="cuda", dtype=torch.bfloat16),
                   torch.empty(nblk, N, device="cuda"))
    _cabi.check(lib.ob_bwd_prep_fused(gout_all[row_base:].data_ptr(), _cabi.OB_PREP_TAIL, rowmask.data_ptr(), None, factor, seed, offset,
                                      thr, row_base, scale.data_ptr(), q.data_ptr(), M, N, K, dys.data_ptr(), qb.data_ptr(),
                                      cs.data_ptr(), _st()))
    gy_all = torch.empty(total, N, device="cuda")
    _cabi.check(lib.ob_residual_dropout_bwd(gout_all.data_ptr(), rowmask.data_ptr(), sc, inv_keep, seed, offset, thr, total, N,
                                            gy_all.data_ptr(), _st()))
    dys2, qb2, cs2 = torch.empty_like(dys), torch.empty_like(qb), torch.empty_like(cs)
    _cabi.check(lib.ob_bwd_prep(gy_all[row_base:].data_ptr(), _cabi.OB_F32, scale.data_ptr(), q.data_ptr(), M, N, K, dys2.data_ptr(),
                                qb2.data_ptr(), cs2.data_ptr(), _st()))
    assert torch.equal(dys, dys2) and torch.equal(qb, qb2)
    assert torch.allclose(cs.sum(0), gy_all[row_base:].sum(0), rtol=1e-5, atol=1e-4)
    assert torch.equal(qb.float(), q.float())


@pytest.mark.parametrize("p", [0.0, 0.1])
def test_fused_prep_swish_equals_swish_backward_then_prep(ob, p):
    from onebit_b200 import _cabi
    lib = _cabi.lib
    M, N, K, row_base = 500, 1024, 256, 32          # N = FFN width (this layer's output), K = its input width
    g = torch.Generator().manual_seed(4)
    total = row_base + M
    gz_all = torch.randn(total, N, generator=g).cuda()
    h_all = (torch.randn(total, N, generator=g) * 2).cuda()
    scale = (torch.rand(M, generator=g) * 50 + 10).cuda()
    q = torch.randint(-128, 128, (M, K), generator=g, dtype=torch.int8).cuda()
    thr = int(round(p * 65536))
    inv_keep, seed, offset = 65536.0 / (65536 - thr), 7, 12
    nblk = lib.ob_bwd_colsum_blocks(M)
    dys, cs = torch.empty(M, N, device="cuda", dtype=torch.bfloat16), torch.empty(nblk, N, device="cuda")
    _cabi.check(lib.ob_bwd_prep_fused(gz_all[row_base:].data_ptr(), _cabi.OB_PREP_SWISH, None, h_all[row_base:].data_ptr(), inv_keep, seed,
                                      offset, thr, row_base, scale.data_ptr(), q.data_ptr(), M, N, K, dys.data_ptr(), None,
                                      cs.data_ptr(), _st()))
    gh_all = torch.empty(total, N, device="cuda")
    _cabi.check(lib.ob_swish_drop_bwd(gz_all.data_ptr(), h_all.data_ptr(), None, inv_keep, seed, offset, thr, total * N,
                                      gh_all.data_ptr(), _st()))
    dys2, cs2 = torch.empty_like(dys), torch.empty_like(cs)
    _cabi.check(lib.ob_bwd_prep(gh_all[row_base:].data_ptr(), _cabi.OB_F32, scale.data_ptr(), q.data_ptr(), M, N, K, dys2.data_ptr(), None,
                                cs2.data_ptr(), _st()))
    _bf16_close(dys, dys2)
    assert torch.allclose(cs.sum(0), gh_all[row_base:].sum(0), rtol=1e-4, atol=1e-3)
    # and against torch autograd of dropout(swish(h)) with the oracle's mask
    keep = torch.ones(total, N)
    if thr:
        keep = torch.from_numpy(orc.dropout_keep_flat(total * N, seed, offset, thr).astype(np.float32).reshape(total, N))
    hh = h_all.cpu().double().requires_grad_(True)
    (torch.nn.functional.silu(hh) * keep.double() * inv_keep * gz_all.cpu().double()).sum().backward()
    assert torch.allclose(gh_all.cpu().double(), hh.grad, rtol=2e-5, atol=2e-6)


def test_layernorm_bwd3_sums_inputs_and_adds_residual(ob):
    from onebit_b200 import fused
    M, C = 700, 256
    g = torch.Generator().manual_seed(8)
    x = torch.randn(M, C, generator=g).cuda()
    w, b = (torch.randn(C, generator=g) * 0.5 + 1).cuda(), torch.zeros(C).cuda()
    dys = [torch.randn(M, C, generator=g).cuda() for _ in range(3)]
    resid = torch.randn(M, C, generator=g).cuda()
    _, _, stats = fused._ln_quant(x, w, b, 1e-5)
    for n in (1, 2, 3):
        for r in (None, resid):
            dx, dg, db = fused._ln_backward(dys[:n], x, stats, w, r)
            xd = x.double().requires_grad_(True)
            wd, bd = w.double().requires_grad_(True), b.double().requires_grad_(True)
            torch.nn.functional.layer_norm(xd, (C,), wd, bd, 1e-5).backward(sum(d.double() for d in dys[:n]))
            want = xd.grad + (0 if r is None else r.double())
            assert torch.allclose(dx.double(), want, rtol=1e-4, atol=1e-5)
            assert torch.allclose(dg.double(), wd.grad, rtol=1e-4, atol=1e-4) and torch.allclose(db.double(), bd.grad, rtol=1e-4, atol=1e-4)


def test_non_finite_activation_rows_propagate(ob):
    """A NaN / Inf in a token row must surface as a non-finite output row, not be clamped into int8 codes."""
    torch.manual_seed(0)
    layer = ob.QuantizedLinear(256, 128).cuda()
    x = torch.randn(200, 256, device="cuda")
    x[5, 17] = float("nan")
    x[9, 3] = float("inf")
    y = layer(x, 2)
    bad = ~torch.isfinite(y).all(dim=1)
    assert bad[5] and bad[9] and int(bad.sum()) == 2


def _block_modules(ob, seed):
    from onebit_b200.asr_model import FeedForwardModule, MHSA, RelPositionalEncoding
    torch.manual_seed(seed)
    ffn = FeedForwardModule(256, 1024, 0.0).cuda().train()
    mhsa = MHSA(256, 4, 0.0).cuda().train()
    pos = RelPositionalEncoding(256, 0.0).cuda()
    with torch.no_grad():
        for m in (ffn, mhsa):
            for n, p in m.named_parameters():
                if n.endswith("bias"):
                    p.normal_(0, 0.05)
    return ffn, mhsa, pos


@pytest.mark.parametrize("stacked", [False, True])
def test_fused_modules_equal_unfused_modules(ob, stacked):
    """FeedForwardModule and MHSA through the fused chains vs the same modules through the separate kernels
    (OB_TORCH_NONROUTED=fuse): identical forward bits, gradients within the bf16 bound - plain 2-bit call and a stacked
    2-bit / 1-bit batch with a ragged frame mask."""
    from onebit_b200 import matmul
    from onebit_b200.asr_model import StackedBits
    ffn, mhsa, pos = _block_modules(ob, 11)
    B, T = 6, 57
    g = torch.Generator().manual_seed(2)
    x0 = torch.randn(B, T, 256, generator=g).cuda()
    gy = torch.randn(B, T, 256, generator=g).cuda()
    lens = torch.tensor([57, 40, 57, 13, 57, 30]).cuda()
    valid = torch.arange(T, device="cuda")[None, :] < lens[:, None]
    mask = valid[:, :, None] & valid[:, None, :]
    bits = StackedBits(4, 3) if stacked else 2
    _, pe = pos(x0)
    res = {}
    for mode in ("fused", "unfused"):
        if mode == "unfused":
            matmul.DISABLED.add("fuse")
        try:
            outs = []
            for mod, call in ((ffn, lambda m, x: m(x, bits, mask)), (mhsa, lambda m, x: m(x, mask, bits, pe))):
                mod.zero_grad(set_to_none=True)
                x = x0.clone().requires_grad_(True)
                y = call(mod, x)
                y.backward(gy)
                outs.append((y.detach(), x.grad, {n: p.grad.clone() for n, p in mod.named_parameters() if p.grad is not None}))
            res[mode] = outs
        finally:
            matmul.DISABLED.discard("fuse")
    for (y_f, gx_f, gp_f), (y_u, gx_u, gp_u) in zip(res["fused"], res["unfused"]):
        assert torch.equal(y_f, y_u)
        assert (gx_f - gx_u).abs().max().item() <= 1e-2 * gx_u.abs().max().item()
        assert gp_f.keys() == gp_u.keys()
        for n in gp_u:
            assert (gp_f[n] - gp_u[n]).abs().max().item() <= 1e-2 * gp_u[n].abs().max().item() + 1e-6, n
