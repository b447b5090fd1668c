"""Does the attention core run faster on utterance chunks whose score matrices stay in the 126 MB L2?  Forward + backward of
rel_attention at the stacked step's shape, whole batch vs chunks of C utterances, device time from a CUDA-graph replay."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import onebit_b200 as ob  # noqa: F401,E402
from onebit_b200.attention import rel_attention  # noqa: E402

B, H, T, d = 192, 4, 399, 64
W = H * d
dev = "cuda"
torch.manual_seed(0)
q, k, v = (torch.randn(B, T, W, device=dev, requires_grad=True) for _ in range(3))
pos = torch.randn(1, T, W, device=dev, requires_grad=True)
u, w = (torch.randn(H, d, device=dev, requires_grad=True) * 0.01 for _ in range(2))
u, w = u.detach().requires_grad_(True), w.detach().requires_grad_(True)
mask = torch.ones(B, T, T, device=dev, dtype=torch.bool)
go = torch.randn(B, T, W, device=dev)


def run(C):
    outs = []
    for b0 in range(0, B, C):
        o = rel_attention(q[b0:b0 + C], k[b0:b0 + C], v[b0:b0 + C], pos, u, w, mask[b0:b0 + C], H, p=0.0, training=True)   # no dropout: its stream cannot be drawn during graph capture
        outs.append(o)
    out = torch.cat(outs, 0) if len(outs) > 1 else outs[0]
    out.backward(go)
    for t in (q, k, v, pos, u, w):
        t.grad = None


def timed(C, reps=3):
    run(C)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            run(C)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for C in (192, 64, 32, 16, 8, 4):
    print(f"chunk of {C:3d} utterances ({C * H * T * T * 4 / 1e6:6.1f} MB per score tensor): fwd + bwd {timed(C):7.2f} ms per block", flush=True)
