"""One launch each of the probs . v product (3 passes, shared-memory operand path) and of its plain-tf32 form at 192 x 4 x 399 x 399, for
ncu --set full --import-source on -k regex:f32_gemm_kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import onebit_b200 as ob  # noqa: F401,E402
from onebit_b200._cabi import lib  # noqa: E402
from onebit_b200.matmul import bmm_nt  # noqa: E402

dev = "cuda"
B, H, T, d = 192, 4, 399, 64
W = H * d
g = torch.Generator(device=dev).manual_seed(0)
p = torch.empty(B, H, T, 400, device=dev)[..., :T]
p.copy_(torch.softmax(torch.randn(B, H, T, T, device=dev, generator=g), -1))
v = torch.randn(B, T, W, device=dev, generator=g)
out = torch.empty_like(v)
hv = v.view(B, T, H, d).permute(0, 2, 1, 3)
ho = out.view(B, T, H, d).permute(0, 2, 1, 3)
for atm, passes in ((0, 3), (0, 1), (2, 3)):          # shared-memory operands, plain tf32, A through tensor memory
    lib.ob_debug_set(11, atm)
    bmm_nt(p, hv.transpose(-1, -2), out=ho, passes=passes)
lib.ob_debug_set(11, 1)
torch.cuda.synchronize()
print("done")
