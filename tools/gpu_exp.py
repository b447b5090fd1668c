"""Timing experiments on the forward GEMM with parts of the kernel disabled (ob_debug_set key 5)."""
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


for (M, K, N, odt) in [(25536, 256, 1024, torch.float32), (76608, 256, 1024, torch.float32), (76608, 1024, 256, torch.float32), (76608, 256, 256, torch.float32)]:
    torch.manual_seed(0)
    layer = ob.QuantizedLinear(K, N).cuda()
    pk, _ = layer.packed_weight(2)
    q, s = ob.act_quant_int8(torch.randn(M, K, device="cuda"))
    for bn in (0,):
        lib.ob_debug_set(_cabi.DBG_FORCE_BLOCK_N, bn)
        for flags, name in [(0, "full"), (1, "no TMA store"), (8, "no fence"), (16, "no wait_read"), (24, "no fence, no wait"), (25, "no fence/wait/store"), (2, "no math"), (3, "no math no store"), (4, "no expansion"), (7, "no expansion no epilogue")]:
            lib.ob_debug_set(5, flags)
            t = graph_time(lambda: obq.gemm_fwd(q, s, pk, layer.alpha, layer.bias, N, odt))
            print(f"M={M} K={K} N={N} {str(odt)[6:]} bn={bn} {name:30s}: {t:7.1f} us", flush=True)
    lib.ob_debug_set(5, 0)
    lib.ob_debug_set(_cabi.DBG_FORCE_BLOCK_N, 0)
