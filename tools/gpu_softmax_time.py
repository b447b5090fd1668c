"""Device time of the attention softmax chain kernels at the stacked step's shape (192 x 4 x 399 x 399), graph replay."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import onebit_b200 as ob  # noqa: F401,E402
from onebit_b200._cabi import lib  # noqa: E402

B, H, T = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (192, 4, 399)
ld = (T + 3) // 4 * 4
dev = "cuda"
mk = lambda: torch.randn(B, H, T, ld, device=dev)  # noqa: E731
ac, bd, g = [mk(), mk()], [mk(), mk()], [mk(), mk()]
y, ad, da, db = mk(), mk(), mk(), mk()
mask = torch.ones(B, T, T, device=dev, dtype=torch.bool)
thr = int(round(0.1 * 65536))
ik = 65536.0 / (65536 - thr)


def t(fn, n=10):
    fn(0)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(gr, stream=side):
            for i in range(n):
                fn(i)
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


cur = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
fw = t(lambda i: lib.ob_relattn_softmax_fwd(ac[i % 2].data_ptr(), bd[i % 2].data_ptr(), mask.data_ptr(), None, ik, 7, 4 * i, thr, 0.125,
                                            B, H, T, ld, y.data_ptr(), ad.data_ptr(), cur()))
bw = t(lambda i: lib.ob_relattn_softmax_bwd(g[i % 2].data_ptr(), y.data_ptr(), None, ik, 7, 4 * i, thr, 0.125, B, H, T, ld,
                                            da.data_ptr(), db.data_ptr(), cur()))
n = float(B) * H * T * T
print(f"B={B} H={H} T={T}: softmax fwd {fw:7.1f} us ({16 * n / fw * 1e-3:6.0f} GB/s)   bwd {bw:7.1f} us ({16 * n / bw * 1e-3:6.0f} GB/s)   "
      f"checksum {y.double().sum().item():.6f} {ad.double().sum().item():.6f}")
