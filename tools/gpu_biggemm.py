"""Where the time of the large forward GEMM (65536 x 2048 x 2048, BASELINE configs[1]) goes: the kernel with parts
disabled (ob_debug_set key 5: 1 no TMA store, 2 no epilogue math, 4 no expansion) next to cuBLASLt's int8 GEMM of the
same shape, as a short burst (one graph of 20 launches) and sustained (~1 s)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import onebit_b200 as ob  # noqa: E402
from onebit_b200 import _cabi, quant as obq  # noqa: E402

lib = _cabi.lib


def graph_time(fn, n=20, reps=3):
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
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / n * 1e3


def main():
    shapes = [(65536, 2048, 2048)]
    if len(sys.argv) > 1 and sys.argv[1] == "all":
        shapes += [(65536, 1024, 1024), (65536, 1024, 2048), (65536, 2048, 1024)]
    for (M, K, N) in shapes:
        torch.manual_seed(0)
        layer = ob.QuantizedLinear(K, N).cuda()
        pk, _ = layer.packed_weight(2)
        q, s = ob.act_quant_int8(torch.randn(M, K, device="cuda").bfloat16())
        ops = 2.0 * M * N * K
        wq = torch.randint(-1, 2, (K, N), device="cuda", dtype=torch.int8)
        t = graph_time(lambda: torch._int_mm(q, wq))
        print(f"M={M} K={K} N={N} cuBLASLt _int_mm (int32 out) burst: {t:7.1f} us  {ops / t * 1e-6:7.1f} TOPS", flush=True)
        t = graph_time(lambda: torch._int_mm(q, wq), n=20, reps=250)
        print(f"M={M} K={K} N={N} cuBLASLt _int_mm (int32 out) ~1 s : {t:7.1f} us  {ops / t * 1e-6:7.1f} TOPS", flush=True)
        for odt in (torch.bfloat16,):
            for flags, name in [(0, "full"), (4, "no expansion"), (3, "no epilogue"), (7, "no expansion, no epilogue")]:
                lib.ob_debug_set(5, flags)
                t = graph_time(lambda: obq.gemm_fwd(q, s, pk, layer.alpha, layer.bias, N, odt))
                t2 = graph_time(lambda: obq.gemm_fwd(q, s, pk, layer.alpha, layer.bias, N, odt), n=20, reps=250)
                print(f"M={M} K={K} N={N} {str(odt)[6:]} {name:28s}: burst {t:7.1f} us {ops / t * 1e-6:7.1f} TOPS | ~1 s {t2:7.1f} us "
                      f"{ops / t2 * 1e-6:7.1f} TOPS", flush=True)
            lib.ob_debug_set(5, 0)


if __name__ == "__main__":
    main()
