"""Experiments with ob_gemm_f32 on a B200: accuracy of every operand-layout combination and timings vs torch.matmul."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import onebit_b200 as ob  # noqa: E402
from onebit_b200._cabi import lib  # noqa: E402
from onebit_b200.matmul import bmm_nt  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"


def err(got, ref64):
    return ((got.double() - ref64).abs().max() / ref64.abs().max()).item()


def check(name, a, b, **kw):
    ref = torch.matmul(a.double(), b.double().transpose(-1, -2))
    if kw.get("bias") is not None:
        ref = ref + kw["bias"].double()
    ref = ref * 1.0
    f32 = torch.matmul(a, b.transpose(-1, -2))
    if kw.get("bias") is not None:
        f32 = f32 + kw["bias"]
    out = {}
    for passes in (3, 1):
        for mode in ((0, 1) if passes == 3 else (0,)):
            lib.ob_debug_set(6, mode)
            try:
                got = bmm_nt(a, b, passes=passes, **kw)
                torch.cuda.synchronize()
                out[f"p{passes}m{mode}"] = err(got, ref)
            except Exception as e:  # noqa: BLE001
                out[f"p{passes}m{mode}"] = f"ERR {e}"
    lib.ob_debug_set(6, 0)
    print(f"{name:44s} torch fp32 {err(f32, ref):.2e} | " + " ".join(f"{k} {v:.2e}" if isinstance(v, float) else f"{k} {v}" for k, v in out.items()),
          flush=True)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


g = torch.Generator(device=dev).manual_seed(0)
R = lambda *s: torch.randn(*s, device=dev, generator=g)  # noqa: E731
what = sys.argv[1] if len(sys.argv) > 1 else "all"

if what in ("all", "acc"):
    check("NT 256x256x256", R(256, 256), R(256, 256))
    check("NT 128x128x32", R(128, 32), R(128, 32))
    check("NT ragged 399x399x64 batch 3x4", R(3, 4, 399, 64), R(3, 4, 399, 64))
    check("NT bias 1000x5004x256", R(1000, 256), R(5004, 256), bias=R(5004))
    check("NN (B mn-major) 399x64x399 batch 2x4", R(2, 4, 399, 400)[..., :399], R(2, 4, 399, 64).transpose(-1, -2))
    check("TN (A,B mn-major) 64x256x1000", R(1000, 64).t(), R(1000, 256).t())
    check("TN batch 399x64x399", R(2, 4, 399, 400)[..., :399].transpose(-1, -2), R(2, 4, 399, 64).transpose(-1, -2))
    check("TK (A mn-major, B k-major) 300x200x500", R(500, 300).t(), R(200, 500))
    qkv = R(3, 399, 4 * 64)
    q = qkv.view(3, 399, 4, 64).permute(0, 2, 1, 3)          # [B,H,T,d] view of [B,T,H*d]
    check("NT strided views q.k^T", q, q)
    pos = R(4, 399, 64)
    check("NT broadcast B over batch", q, pos.unsqueeze(0))
    check("TN deep K (split-K) 512x256x25536", R(25536, 512).t(), R(25536, 256).t())
    check("TN deep K (split-K) 5004x256x9000 bias", R(9000, 5004).t(), R(9000, 256).t(), bias=R(256))
    check("NN 2000x256x5004", R(2000, 5004), R(5004, 256).t())
    # accumulate
    a, b = R(2, 4, 399, 64), R(2, 4, 399, 64)
    out = torch.zeros(2, 4, 399, 400, device=dev)[..., :399]
    bmm_nt(a, b, out=out)
    bmm_nt(a, b, out=out, accumulate=True)
    ref = 2 * torch.matmul(a.double(), b.double().transpose(-1, -2))
    print("accumulate x2 err", err(out, ref), flush=True)

if what in ("all", "time"):
    B, H, T, d = 64, 4, 399, 64
    q, k = R(B, H, T, d), R(B, H, T, d)
    sc = torch.empty(B, H, T, 400, device=dev)[..., :T]
    att = R(B, H, T, 400)[..., :T]
    v = R(B, H, T, d)
    o = torch.empty(B, H, T, d, device=dev)
    x, w = R(25536, 256), R(5004, 256)
    y = torch.empty(25536, 5004, device=dev)
    w2 = R(512, 256)
    y2 = torch.empty(25536, 512, device=dev)
    for epi, name in ((1, "direct row stores"), (2, "TMA box stores"), (0, "auto")):
        lib.ob_debug_set(7, epi)
        print(f"passes=3, epilogue: {name}")
        print(f"  scores q.k^T  ours {timeit(lambda: bmm_nt(q, k, out=sc)):8.1f} us")
        print(f"  dO.v^T        ours {timeit(lambda: bmm_nt(v, k, out=sc)):8.1f} us")
        print(f"  attn.v        ours {timeit(lambda: bmm_nt(att, v.transpose(-1, -2), out=o)):8.1f} us")
        print(f"  attn^T.do     ours {timeit(lambda: bmm_nt(att.transpose(-1, -2), v.transpose(-1, -2), out=o)):8.1f} us")
        print(f"  vocab fwd     ours {timeit(lambda: bmm_nt(x, w, out=y)):8.1f} us")
        print(f"  vocab dx      ours {timeit(lambda: bmm_nt(y, w.t(), out=x)):8.1f} us")
        print(f"  pw1 fwd       ours {timeit(lambda: bmm_nt(x, w2, out=y2)):8.1f} us", flush=True)
    lib.ob_debug_set(7, 0)
    print(f"torch: scores {timeit(lambda: torch.matmul(q, k.transpose(-1, -2))):8.1f} us, attn.v {timeit(lambda: torch.matmul(att, v)):8.1f} us, "
          f"vocab fwd {timeit(lambda: torch.nn.functional.linear(x, w)):8.1f} us")
