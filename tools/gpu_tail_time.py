"""Forward GEMM with the fused module tail vs plain GEMM + residual_dropout kernel, model shapes (quick timing aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import onebit_b200 as ob
from onebit_b200 import _cabi
lib = _cabi.lib
dev = "cuda"
st = torch.cuda.current_stream().cuda_stream


def timeit(f, n=40):
    for i in range(3):
        f(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        f(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for M in (25536, 51072, 76608):
    for K, N in ((1024, 256), (256, 256)):
        torch.manual_seed(0)
        layer = ob.QuantizedLinear(K, N).to(dev)
        pk, _ = layer.packed_weight(2)
        nb = 4
        qs = [ob.act_quant_int8(torch.randn(M, K, device=dev)) for _ in range(nb)]
        ys = [torch.empty(M, N, device=dev) for _ in range(nb)]
        outs = [torch.empty(M, N, device=dev) for _ in range(nb)]
        res = [torch.randn(M, N, device=dev) for _ in range(nb)]
        a = layer.alpha
        for thr in (0, 6554):
            ik = 65536.0 / (65536 - thr)
            t_f = timeit(lambda i: lib.ob_gemm_tern_i8_fwd(qs[i % nb][0].data_ptr(), qs[i % nb][1].data_ptr(), pk.data_ptr(), a.data_ptr(), 1,
                                                          layer.bias.data_ptr(), M, N, K, ys[i % nb].data_ptr(), 0, st))
            t_r = timeit(lambda i: lib.ob_residual_dropout_fwd(res[i % nb].data_ptr(), ys[i % nb].data_ptr(), None, 0.5, ik, 1, 4 * i, thr, M, N,
                                                              outs[i % nb].data_ptr(), st))
            t_t = timeit(lambda i: lib.ob_gemm_tern_i8_fwd_tail(qs[i % nb][0].data_ptr(), qs[i % nb][1].data_ptr(), pk.data_ptr(), a.data_ptr(), 1,
                                                               layer.bias.data_ptr(), M, N, K, res[i % nb].data_ptr(), None, 0.5 * ik, 1, 4 * i,
                                                               thr, 0, outs[i % nb].data_ptr(), st))
            by = M * K + N * K / 4 + 8.0 * M * N + 4 * M
            print(f"M={M} {K}->{N} dropout={'on' if thr else 'off'}: gemm {t_f:.1f} + tail kernel {t_r:.1f} = {t_f + t_r:.1f} us;  fused {t_t:.1f} us "
                  f"({by / t_t / 1e3:.0f} GB/s)")
