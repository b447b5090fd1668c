"""Two launches of ob_gemm_f32 for ncu: the vocabulary projection (deep tiles) and the attention scores (output-bound)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import onebit_b200  # noqa: E402,F401
from onebit_b200.matmul import bmm_nt  # noqa: E402

dev = "cuda"
x, w = torch.randn(25536, 256, device=dev), torch.randn(5004, 256, device=dev)
y = torch.empty(25536, 5004, device=dev)
q, k = torch.randn(64, 4, 399, 64, device=dev), torch.randn(64, 4, 399, 64, device=dev)
sc = torch.empty(64, 4, 399, 400, device=dev)[..., :399]
att = torch.randn(64, 4, 399, 400, device=dev)[..., :399]
o = torch.empty(64, 4, 399, 64, device=dev)
for _ in range(2):
    bmm_nt(x, w, out=y)
    bmm_nt(q, k, out=sc)
    bmm_nt(att, k.transpose(-1, -2), out=o)
torch.cuda.synchronize()
print("done")
