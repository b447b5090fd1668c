"""fp32 tensor-core GEMM (ob_gemm_f32, 3 x tf32 split) and the attention core built on it, against fp64 / the reference's op
sequence in fp32.  Tolerance: 2e-5 of max|result| (measured <= 1e-5; plain tf32 is ~8e-4, torch's fp32 SIMT path ~5e-7)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 2e-5


@pytest.fixture(scope="module")
def ob():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import onebit_b200
    return onebit_b200


def rel(got, ref64):
    return ((got.double() - ref64).abs().max() / ref64.abs().max().clamp_min(1e-30)).item()


def R(*shape, seed=0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return torch.randn(*shape, generator=g).cuda()


CASES = {
    "nt_square": lambda: (R(256, 256), R(256, 256, seed=1)),
    "nt_tiny": lambda: (R(1, 8), R(3, 8, seed=1)),
    "nt_ragged_batched": lambda: (R(3, 4, 399, 64), R(3, 4, 399, 64, seed=1)),
    "nt_wide_n": lambda: (R(1000, 256), R(5004, 256, seed=1)),
    "nn_b_mn_major": lambda: (R(2, 4, 399, 400)[..., :399], R(2, 4, 399, 64, seed=1).transpose(-1, -2)),
    "tn_both_mn_major": lambda: (R(1000, 64).t(), R(1000, 256, seed=1).t()),
    "tn_batched": lambda: (R(2, 4, 399, 400)[..., :399].transpose(-1, -2), R(2, 4, 399, 64, seed=1).transpose(-1, -2)),
    "tk_a_mn_major": lambda: (R(500, 300).t(), R(200, 500, seed=1)),
    "deep_k_split": lambda: (R(25536, 512).t(), R(25536, 256, seed=1).t()),
    "k_not_multiple_of_32": lambda: (R(77, 45), R(130, 45, seed=1)),
    "broadcast_b": lambda: (R(3, 4, 100, 64), R(1, 4, 100, 64, seed=1)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_bmm_nt_matches_fp64(ob, name):
    from onebit_b200.matmul import bmm_nt
    a, b = CASES[name]()
    ref = torch.matmul(a.double(), b.double().transpose(-1, -2))
    got = bmm_nt(a, b)
    assert got.shape == ref.shape
    assert rel(got, ref) < TOL
    assert rel(bmm_nt(a, b, passes=1), ref) < 3e-3                    # plain tf32, for comparison


def test_bmm_nt_strided_views_bias_scale_accumulate(ob):
    from onebit_b200.matmul import bmm_nt
    qkv = R(3, 399, 256)
    q = qkv.view(3, 399, 4, 64).permute(0, 2, 1, 3)                   # [B, H, T, d] view of [B, T, H*d]
    ref = torch.matmul(q.double(), q.double().transpose(-1, -2))
    assert rel(bmm_nt(q, q), ref) < TOL
    out = torch.empty(3, 399, 256, device="cuda")
    probs = torch.softmax(R(3, 4, 399, 399, seed=2), -1)
    bmm_nt(probs, q.transpose(-1, -2), out=out.view(3, 399, 4, 64).permute(0, 2, 1, 3))     # written through a strided view
    ref = torch.matmul(probs.double(), q.double()).permute(0, 2, 1, 3).reshape(3, 399, 256)
    assert rel(out, ref) < TOL
    x, w, bias = R(300, 256), R(512, 256, seed=1), R(512, seed=2)
    ref = torch.nn.functional.linear(x.double(), w.double(), bias.double()) * 1.0
    assert rel(bmm_nt(x, w, bias=bias), ref) < TOL
    ref2 = 0.5 * torch.matmul(x.double(), w.double().t())
    y = bmm_nt(x, w, scale=0.5)
    assert rel(y, ref2) < TOL
    bmm_nt(x, w, out=y, scale=0.5, accumulate=True)
    assert rel(y, 2 * ref2) < TOL
    # sum over a batch axis: output shared by the batch (stride 0) + accumulate
    a, b = R(5, 1, 64, 40), R(5, 1, 32, 40, seed=1)
    acc = torch.zeros(1, 1, 64, 32, device="cuda")
    bmm_nt(a, b, out=acc.expand(5, 1, 64, 32), accumulate=True)
    assert rel(acc[0, 0], torch.matmul(a.double(), b.double().transpose(-1, -2)).sum(0)[0]) < TOL


@pytest.mark.parametrize("name", ["nt_wide_n", "tk_a_mn_major", "deep_k_split", "nn_wide_k"])
def test_cta_pair_tiles_equal_single_cta_tiles(ob, name):
    """The 256-wide tiles run on CTA pairs (cta_group::2) by default; ob_debug_set(8, 0) selects single CTAs.  Same MMA
    sequence per output element, so the results are bitwise equal (tools/gpu_f32pair.py measures the same on more shapes)."""
    from onebit_b200._cabi import lib
    from onebit_b200.matmul import bmm_nt
    a, b = CASES[name]() if name in CASES else (R(700, 516), R(516, 256, seed=1).t())
    try:
        assert lib.ob_debug_set(8, 0) == 0
        single = bmm_nt(a, b)
        assert lib.ob_debug_set(8, 1) == 0
        pair = bmm_nt(a, b)
        torch.cuda.synchronize()
    finally:
        lib.ob_debug_set(8, 1)
    assert torch.equal(single, pair)
    assert rel(pair, a.double() @ b.double().transpose(-1, -2)) <= TOL


@pytest.mark.parametrize("name", ["nt_ragged_batched", "nn_b_mn_major", "tn_batched", "tn_both_mn_major", "k_not_multiple_of_32",
                                  "broadcast_b", "nt_tiny"])
def test_a_operand_through_tensor_memory_equals_shared_memory_operand(ob, name):
    """The single-CTA split products can keep the A operand in tensor memory (tcgen05.mma with A from TMEM, hi / lo written by
    tcgen05.st; default for the 128-wide tiles, ob_debug_set(11, 2) also for the 64-wide ones, 0 = off).  Same operand values
    and MMA order per output element, so the results are bitwise equal in every layout, incl. the transposed A that is then
    read from un-swizzled boxes."""
    from onebit_b200._cabi import lib
    from onebit_b200.matmul import bmm_nt
    a, b = CASES[name]()
    try:
        assert lib.ob_debug_set(11, 0) == 0
        smem = bmm_nt(a, b)
        assert lib.ob_debug_set(11, 2) == 0
        tmem = bmm_nt(a, b)
        torch.cuda.synchronize()
    finally:
        lib.ob_debug_set(11, 1)
    assert torch.equal(smem, tmem)
    assert rel(tmem, torch.matmul(a.double(), b.double().transpose(-1, -2))) < TOL


def test_bmm_nt_rejects_bad_arguments(ob):
    from onebit_b200.matmul import bmm_nt
    with pytest.raises(RuntimeError):
        bmm_nt(torch.randn(4, 4), torch.randn(4, 4))                  # CPU tensors: no fallback
    with pytest.raises(ValueError):
        bmm_nt(R(4, 8), R(4, 12))
    with pytest.raises(ValueError):
        bmm_nt(R(4, 8), R(4, 8), accumulate=True)


@pytest.mark.parametrize("B,H,T,p", [(2, 4, 49, 0.0), (3, 4, 249, 0.1), (2, 2, 399, 0.1), (1, 4, 33, 0.0)])
def test_rel_attention_matches_reference_sequence(ob, B, H, T, p):
    """The attention core (three products + chain forward, six products + chain backward) against the reference's op
    sequence (conformer.py:113-129) in torch fp32 with the same dropout mask."""
    from onebit_b200.attention import rel_attention
    from onebit_b200.conformer import MHSA
    d = 64
    W = H * d
    g = torch.Generator().manual_seed(B * 100 + T)
    mk = lambda *s: torch.randn(*s, generator=g).cuda()  # noqa: E731
    q0, k0, v0, pos0 = mk(B, T, W), mk(B, T, W), mk(B, T, W), mk(1, T, W)
    u0, w0 = mk(H, d) * 0.5, mk(H, d) * 0.5
    lens = torch.randint(1, T + 1, (B,), generator=g)
    lens[0] = T
    km = (torch.arange(T)[None, :] < lens[:, None]).cuda()
    mask = km[:, :, None] & km[:, None, :]
    keep = (torch.rand(B, H, T, T, generator=g) > p).cuda() if p > 0 else None
    go = mk(B, T, W)

    def reference(q, k, v, pos, u, w):
        split = lambda t, b: t.view(b, -1, H, d).transpose(1, 2)  # noqa: E731
        qh, kh, vh, ph = split(q, B), split(k, B), split(v, B), split(pos, 1)
        ac = torch.matmul(qh + u.view(1, H, 1, d), kh.transpose(-2, -1))
        bd = MHSA.rel_shift(torch.matmul(qh + w.view(1, H, 1, d), ph.transpose(-2, -1)))
        s = ((ac + bd) / math.sqrt(d)).masked_fill(mask[:, None] == 0, float("-inf"))
        a = torch.nan_to_num(torch.softmax(s, dim=-1), nan=0.0)
        if keep is not None:
            a = a * keep.float() * (1.0 / (1.0 - p))
        return torch.matmul(a, vh).transpose(1, 2).contiguous().view(B, T, W)

    outs = []
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for fn in (lambda *a: rel_attention(*a, mask, H, p, True, keep=keep), reference):
            leaves = [t.clone().requires_grad_(True) for t in (q0, k0, v0, pos0, u0, w0)]
            out = fn(*leaves)
            out.backward(go)
            outs.append([out.detach()] + [t.grad for t in leaves])
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    names = ["out", "g_q", "g_k", "g_v", "g_pos", "g_u", "g_w"]
    for name, a, b in zip(names, *outs):
        err = ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
        assert err < 5e-5, (name, err)
    if (~km).any():
        assert outs[0][0][~km].abs().max().item() == 0.0              # padded query rows are exactly zero


@pytest.mark.parametrize("shape,n_out", [((4, 249, 256), 5004), ((1000, 4864), 256), ((3, 7, 64), 10)])
def test_linear_matches_torch_fp32(ob, shape, n_out):
    from onebit_b200.matmul import linear
    x0, w0, b0, gy = R(*shape), R(n_out, shape[-1], seed=1) * 0.1, R(n_out, seed=2), R(*shape[:-1], n_out, seed=3)
    outs = []
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for fn in (linear, torch.nn.functional.linear):
            x, w, b = (t.clone().requires_grad_(True) for t in (x0, w0, b0))
            y = fn(x, w, b)
            y.backward(gy)
            outs.append((y.detach(), x.grad, w.grad, b.grad))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    for name, a, b in zip(("y", "g_x", "g_w", "g_b"), *outs):
        assert a.shape == b.shape and a.is_contiguous()
        err = ((a - b).abs().max() / b.abs().max()).item()
        assert err < TOL, (name, err)


@pytest.mark.parametrize("B,T,C,ks", [(3, 130, 64, 31), (2, 399, 256, 31), (4, 17, 128, 7), (1, 64, 64, 15)])
def test_conv_module_middle_matches_torch(ob, B, T, C, ks):
    """swish(BatchNorm(depthwise(GLU(a)))) and every gradient against torch's own ops (conformer.py:141-167)."""
    import torch.nn.functional as F
    from onebit_b200.convmod import glu_dwconv_bn_swish
    a0 = R(B, T, 2 * C)
    w0, b0 = R(C, 1, ks, seed=1) * 0.3, R(C, seed=2) * 0.1
    ga0, be0 = R(C, seed=3) * 0.5 + 1.0, R(C, seed=4) * 0.2
    gy = R(B, T, C, seed=5)

    def reference(a, w, b, gamma, beta):
        t = F.glu(a.transpose(1, 2), dim=1)
        t = F.conv1d(t, w, b, padding=ks // 2, groups=C)
        t = F.batch_norm(t, None, None, gamma, beta, True, 0.1, 1e-5)
        return (t * torch.sigmoid(t)).transpose(1, 2)

    outs = []
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for fn in (lambda *t: glu_dwconv_bn_swish(*t, 1e-5), reference):
            leaves = [t.clone().requires_grad_(True) for t in (a0, w0, b0, ga0, be0)]
            y = fn(*leaves)
            y.backward(gy)
            outs.append([y.detach()] + [t.grad for t in leaves])
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    for name, x, y in zip(("s", "g_a", "g_w", "g_bias", "g_gamma", "g_beta"), *outs):
        assert x.shape == y.shape, name
        # a bias in front of BatchNorm has no gradient (the mean is removed): both sides are round-off of a zero sum, so
        # it is compared on the scale of the other per-channel sums
        scale = outs[1][5].abs().max() if name == "g_bias" else y.abs().max().clamp_min(1e-30)
        err = ((x - y).abs().max() / scale).item()
        assert err < (2e-4 if name.startswith("g_") else 2e-5), (name, err)


def test_conv_module_channel_last_equals_torch_path(ob):
    """The whole ConvModule: B200 channel-last path (CUDA) against the module's torch path on a float64 CPU copy."""
    from onebit_b200.conformer import ConvModule
    torch.manual_seed(3)
    mod = ConvModule(256, 31, 0.0).cuda()
    ref = ConvModule(256, 31, 0.0).double()
    ref.load_state_dict({k: v.double().cpu() for k, v in mod.state_dict().items()})
    x0 = R(2, 150, 256)
    gy = R(2, 150, 256, seed=1)
    x = x0.clone().requires_grad_(True)
    y = mod(x)
    y.backward(gy)
    xr = x0.double().cpu().requires_grad_(True)
    yr = ref(xr)
    yr.backward(gy.double().cpu())
    assert (y.detach().cpu().double() - yr.detach()).abs().max().item() < 2e-5 * yr.abs().max().item()
    assert (x.grad.cpu().double() - xr.grad).abs().max().item() < 2e-4 * xr.grad.abs().max().item()
    bn_scale = ref.bn.bias.grad.abs().max().item()
    for (n, p), (_, pr) in zip(mod.named_parameters(), ref.named_parameters()):
        # the depthwise bias sits in front of BatchNorm: its gradient is a zero sum, compared on the scale of g_beta
        scale = bn_scale if n == "dw.bias" else pr.grad.abs().max().clamp_min(1e-30).item()
        assert (p.grad.cpu().double() - pr.grad).abs().max().item() < 2e-4 * scale, n


def test_conv_module_middle_groups(ob):
    """groups = 3: a stack of three passes, BatchNorm statistics per pass == three separate calls."""
    from onebit_b200.convmod import glu_dwconv_bn_swish
    B, T, C, ks = 6, 77, 64, 31
    a0, w0, b0 = R(B, T, 2 * C), R(C, 1, ks, seed=1) * 0.3, R(C, seed=2) * 0.1
    ga0, be0, gy = R(C, seed=3) * 0.5 + 1.0, R(C, seed=4) * 0.2, R(B, T, C, seed=5)
    outs = []
    for stacked in (True, False):
        leaves = [t.clone().requires_grad_(True) for t in (a0, w0, b0, ga0, be0)]
        a, rest = leaves[0], leaves[1:]
        if stacked:
            y = glu_dwconv_bn_swish(a, *rest, 1e-5, 3)
        else:
            y = torch.cat([glu_dwconv_bn_swish(a[2 * g:2 * g + 2], *rest, 1e-5, 1) for g in range(3)], dim=0)
        y.backward(gy)
        outs.append([y.detach()] + [t.grad for t in leaves])
    for name, x, y in zip(("s", "g_a", "g_w", "g_bias", "g_gamma", "g_beta"), *outs):
        scale = outs[1][5].abs().max() if name == "g_bias" else y.abs().max().clamp_min(1e-30)
        assert ((x - y).abs().max() / scale).item() < 1e-5, name


@pytest.mark.parametrize("B,T,F", [(3, 200, 80), (2, 37, 23), (1, 3, 3)])
def test_frontend_conv1_relu_matches_torch(ob, B, T, F):
    """relu(conv2d(x, w, b, stride 2)) of the one-channel first front-end layer, forward and parameter gradients."""
    from onebit_b200.frontend import conv1_relu
    x = R(B, T, F)
    w0, b0 = R(256, 1, 3, 3, seed=1) * 0.3, R(256, seed=2) * 0.1
    outs = []
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for ours in (True, False):
            w, b = w0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
            y = conv1_relu(x, w, b) if ours else torch.relu(torch.nn.functional.conv2d(x[:, None], w, b, stride=2))
            gy = R(*y.shape, seed=3)
            y.backward(gy)
            outs.append((y.detach(), w.grad, b.grad))
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    for name, a, b in zip(("y", "g_w", "g_b"), *outs):
        assert a.shape == b.shape, name
        err = ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
        assert err < (1e-5 if name == "y" else 2e-4), (name, err)
    assert outs[0][0].is_contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("M,N", [(1, 4), (37, 256), (25536, 512), (4160, 5004), (76608, 256)])
def test_column_sums_kernel(M, N):
    """ob_colsum (bias gradients of the non-routed linears) vs float64, and bitwise run-to-run (fixed summation order)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from onebit_b200.matmul import column_sums
    x = torch.randn(M, N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(M + N))
    got = column_sums(x)
    want = x.double().sum(0)
    assert (got.double() - want).abs().max().item() <= 1e-5 * (x.double().abs().sum(0).max().item() + 1e-30)
    assert torch.equal(got, column_sums(x))
