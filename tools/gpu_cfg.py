"""Tiling sweep of the forward and grad_x GEMMs at the model shapes (CUDA-graph timing)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import onebit_b200 as ob  # noqa: E402
from onebit_b200 import _cabi, quant as obq  # noqa: E402

lib = _cabi.lib


def graph_time(fn, n=20):
    fn()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(n):
                fn()
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 3 / n * 1e3


for (M, K, N) in [(25536, 256, 256), (25536, 256, 1024), (25536, 1024, 256), (102144, 256, 256), (102144, 1024, 256)]:
    torch.manual_seed(0)
    layer = ob.QuantizedLinear(K, N).cuda()
    pk, pkt = layer.packed_weight(2)
    q, s = ob.act_quant_int8(torch.randn(M, K, device="cuda"))
    dys = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    dx = torch.empty(M, K, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    row = f"M={M} K={K} N={N}:"
    for bn in (0, 64, 128, 256, 1128, 1256):
        lib.ob_debug_set(_cabi.DBG_FORCE_BLOCK_N, bn)
        tf = graph_time(lambda: obq.gemm_fwd(q, s, pk, layer.alpha, layer.bias, N, torch.float32))
        tb = graph_time(lambda: lib.ob_bwd_dx(dys.data_ptr(), s.data_ptr(), pkt.data_ptr(), layer.alpha.data_ptr(), 1, M, N, K,
                                              dx.data_ptr(), 0, torch.cuda.current_stream().cuda_stream))
        row += f"  bn={bn}: fwd {tf:5.1f} dx {tb:5.1f}"
    lib.ob_debug_set(_cabi.DBG_FORCE_BLOCK_N, 0)
    print(row, flush=True)
