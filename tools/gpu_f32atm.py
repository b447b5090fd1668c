"""A operand through tensor memory in the fp32 tensor-core GEMM (ob_debug_set key 11): bitwise A/B against the shared-memory
operand path over the attention product shapes and layouts, accuracy against fp64, device time per launch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import onebit_b200 as ob  # noqa: F401,E402
from onebit_b200._cabi import lib  # noqa: E402
from onebit_b200.matmul import bmm_nt  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
R = lambda *s: torch.randn(*s, device=dev, generator=g)  # noqa: E731


def heads(t, H):
    B, T, W = t.shape
    return t.view(B, T, H, W // H).permute(0, 2, 1, 3)


def scores_like(B, H, T):
    ld = (T + 3) // 4 * 4
    return torch.empty(B, H, T, ld, device=dev)[..., :T]


def cases(B, H, T, d=64):
    W = H * d
    q, k, v = R(B, T, W), R(B, T, W), R(B, T, W)
    pos = R(1, T, W)
    p = scores_like(B, H, T)
    p.copy_(torch.softmax(R(B, H, T, T), -1))
    go = R(B, T, W)
    yield "scores q.k^T      (A K, B K, N=T)", heads(q, H), heads(k, H), lambda: scores_like(B, H, T)
    yield "scores q.pos^T    (B broadcast)  ", heads(q, H), heads(pos, H), lambda: scores_like(B, H, T)
    yield "probs.v           (A K, B MN, N=d)", p, heads(v, H).transpose(-1, -2), lambda: heads(torch.empty_like(q), H)
    yield "probs^T.dO        (A MN, B MN)    ", p.transpose(-1, -2), heads(go, H).transpose(-1, -2), lambda: heads(torch.empty_like(q), H)
    yield "dS^T.q -> sum over batch (A MN)   ", p.transpose(-1, -2), heads(q, H).transpose(-1, -2), lambda: heads(torch.empty_like(q), H)


def timeit(fn, n=10):
    """device time per call from a CUDA-graph replay (the Python wrapper costs more than the kernel)"""
    fn()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(gr, stream=side):
            for _ in range(n):
                fn()
    torch.cuda.current_stream().wait_stream(side)
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * n) * 1e3


bad = 0
for (B, H, T) in [(2, 4, 399)]:
    for name, a, b, mk in cases(B, H, T):
        ref = torch.matmul(a.double(), b.double().transpose(-1, -2))
        outs = []
        for atm in (0, 1):
            lib.ob_debug_set(11, atm)
            o = mk()
            o.fill_(float("nan"))
            bmm_nt(a, b, out=o)
            torch.cuda.synchronize()
            outs.append(o.clone())
        e = ((outs[1].double() - ref).abs().max() / ref.abs().max()).item()
        same = torch.equal(outs[0], outs[1])
        ok = same and e < 2e-5
        bad += 0 if ok else 1
        print(f"B={B} H={H} T={T:4d} {name}: bitwise {'==' if same else '!='}  err vs fp64 {e:.2e} {'OK' if ok else 'FAIL'}", flush=True)
# plain matrices that take the single-CTA kernels (narrow N or shallow K), ragged sizes
for (M, N, K) in [(1000, 64, 512), (777, 100, 96), (130, 128, 64), (5000, 40, 1000)]:
    a, b = R(M, K), R(N, K)
    ref = a.double() @ b.double().t()
    outs = []
    for atm in (0, 1):
        lib.ob_debug_set(11, atm)
        outs.append(bmm_nt(a, b).clone())
    e = ((outs[1].double() - ref).abs().max() / ref.abs().max()).item()
    same = torch.equal(outs[0], outs[1])
    bad += 0 if (same and e < 2e-5) else 1
    print(f"M={M} N={N} K={K}: bitwise {'==' if same else '!='}  err vs fp64 {e:.2e}", flush=True)
print("A/B:", "OK" if bad == 0 else f"{bad} FAILED", flush=True)

B, H, T = 192, 4, 399
for name, a, b, mk in cases(B, H, T):
    o = mk()
    ts = []
    for atm, passes in ((0, 3), (1, 3), (0, 1)):
        lib.ob_debug_set(11, atm)
        ts.append(timeit(lambda: bmm_nt(a, b, out=o, passes=passes)))
    print(f"B={B} H={H} T={T} {name}: 3 passes smem {ts[0]:6.1f}  3 passes TMEM {ts[1]:6.1f}  1 pass (plain tf32) {ts[2]:6.1f} us", flush=True)
lib.ob_debug_set(11, 1)
