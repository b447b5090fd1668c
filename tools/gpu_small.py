"""A few launches of the streaming kernels around the layer at the bench shape (for ncu captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import onebit_b200 as ob  # noqa: E402
from onebit_b200 import _cabi  # noqa: E402

lib = _cabi.lib
M, K, N = 25536, 256, 1024
dev = "cuda"
st = torch.cuda.current_stream().cuda_stream
h = torch.randn(M, N, device=dev)
keep = torch.rand(M, N, device=dev) > 0.1
q = torch.empty(M, N, device=dev, dtype=torch.int8)
s = torch.empty(M, device=dev)
x = torch.randn(M, K, device=dev)
dy = torch.randn(M, K, device=dev)
w, b = torch.ones(K, device=dev), torch.zeros(K, device=dev)
y, stats = torch.empty(M, K, device=dev), torch.empty(2, M, device=dev)
dx, dp = torch.empty(M, K, device=dev), torch.empty(2, K, device=dev)
ws = torch.empty(lib.ob_layernorm_bwd_workspace_bytes(K), device=dev, dtype=torch.uint8)
for _ in range(4):
    lib.ob_swish_drop_quant(h.data_ptr(), keep.data_ptr(), 1 / 0.9, 0, 0, 0, M, N, q.data_ptr(), s.data_ptr(), st)
    lib.ob_layernorm_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), 1e-5, M, K, y.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(), st)
    lib.ob_layernorm_bwd(dy.data_ptr(), x.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(), w.data_ptr(), M, K, dx.data_ptr(),
                         dp[0].data_ptr(), dp[1].data_ptr(), ws.data_ptr(), st)
    lib.ob_act_quant_i8(x.data_ptr(), 0, M, K, q.data_ptr(), s.data_ptr(), st)
torch.cuda.synchronize()
print("done")
