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

# convolution-module kernels and the attention chain at the stacked (3 passes) training shape, for the same capture
from onebit_b200._cabi import lib  # noqa: E402
from onebit_b200.convmod import glu_dwconv_bn_swish  # noqa: E402
from onebit_b200.attention import rel_attention_probs  # noqa: E402

a = torch.randn(192, 399, 512, device=dev, requires_grad=True)
dw_w, dw_b = torch.randn(256, 1, 31, device=dev, requires_grad=True), torch.zeros(256, device=dev, requires_grad=True)
gam, bet = torch.ones(256, device=dev, requires_grad=True), torch.zeros(256, device=dev, requires_grad=True)
for _ in range(2):
    s = glu_dwconv_bn_swish(a, dw_w, dw_b, gam, bet, 1e-5, 3)
    s.backward(torch.randn_like(s))
ac, bd = torch.randn(64, 4, 399, 399, device=dev, requires_grad=True), torch.randn(64, 4, 399, 399, device=dev, requires_grad=True)
mask = torch.ones(64, 399, 399, dtype=torch.bool, device=dev)
for _ in range(2):
    p = rel_attention_probs(ac, bd, mask, 0.125, 0.1, True)
    p.backward(torch.randn_like(p))
torch.cuda.synchronize()
print("done 2")
