"""One launch of every kernel of the quantised layer at the model's widest routed shape (M = 25536 tokens, 256 -> 1024) and of the
large GEMM (65536 x 2048 x 2048), behind a cache flush, for `ncu --set full`:

    ncu --set full --clock-control none --import-source on -k regex:'gemm_expand|dw_|bwd_prep|act_quant|ln_quant|gemv|swish_drop_quant' \
        -o gpurun_out/r02b_layer python tools/gpu_layer_ncu.py
    ncu -i gpurun_out/r02b_layer.ncu-rep --page raw --csv > gpurun_out/r02b_layer_raw.csv
    python tools/ncu_traffic.py gpurun_out/r02b_layer_raw.csv   # per-kernel table + gpurun_out/ncu_kernel_summary.json
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import onebit_b200 as ob  # noqa: E402
from onebit_b200 import _cabi, fused, quant as obq  # noqa: E402

lib = _cabi.lib
dev = "cuda"
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(M, K, N, big=False):
    torch.manual_seed(0)
    layer = ob.QuantizedLinear(K, N).to(dev)
    pk, pkt = layer.packed_weight(2)
    x = torch.randn(M, K, device=dev)
    gy = torch.randn(M, N, device=dev)
    lnw, lnb = torch.ones(K, device=dev), torch.zeros(K, device=dev)
    a = layer.alpha
    for rep in range(2):                                            # 2nd repetition is the one to read (first = warm-up)
        flush.zero_()
        if big:
            xb = x.to(torch.bfloat16)
            flush.zero_()
            q, s = ob.act_quant_int8(xb)
            flush.zero_()
            obq.gemm_fwd(q, s, pk, a, layer.bias, N, torch.bfloat16)
            continue
        q, s = ob.act_quant_int8(x)
        flush.zero_()
        q, s, stats = fused._ln_quant(x, lnw, lnb, 1e-5)
        flush.zero_()
        y = obq.gemm_fwd(q, s, pk, a, layer.bias, N, torch.float32)
        flush.zero_()
        dys = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        colsum = torch.empty(lib.ob_bwd_colsum_blocks(M), N, device=dev)
        lib.ob_bwd_prep(gy.data_ptr(), 0, s.data_ptr(), q.data_ptr(), M, N, K, dys.data_ptr(), None, colsum.data_ptr(), st)
        flush.zero_()
        dx = torch.empty(M, K, device=dev)
        lib.ob_bwd_dx(dys.data_ptr(), s.data_ptr(), pkt.data_ptr(), a.data_ptr(), 1, M, N, K, dx.data_ptr(), 0, st)
        flush.zero_()
        gw, ga, gb = torch.empty(N, K, device=dev), torch.empty((), device=dev), torch.empty(N, device=dev)
        nbytes = lib.ob_bwd_dw_workspace_bytes(M, N, K)
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        lib.ob_bwd_dw_q8(dys.data_ptr(), q.data_ptr(), colsum.data_ptr(), layer.weight.data_ptr(), a.data_ptr(), 1, 2, M, N, K,
                         gw.data_ptr(), ga.data_ptr(), gb.data_ptr(), ws.data_ptr(), nbytes, st)
        flush.zero_()
        # fused neighbours: tail epilogue (1024 -> 256 as lin2), swish-mode prep, swish+dropout quantiser
        layer2 = ob.QuantizedLinear(N, K).to(dev)
        pk2, _ = layer2.packed_weight(2)
        thr = int(0.1 * 65536)
        q2 = torch.empty(M, N, device=dev, dtype=torch.int8)
        s2 = torch.empty(M, device=dev)
        lib.ob_swish_drop_quant(y.data_ptr(), None, 65536.0 / (65536 - thr), 1, 0, thr, M, N, q2.data_ptr(), s2.data_ptr(), st)
        flush.zero_()
        out = torch.empty(M, K, device=dev)
        lib.ob_gemm_tern_i8_fwd_tail(q2.data_ptr(), s2.data_ptr(), pk2.data_ptr(), layer2.alpha.data_ptr(), 1, layer2.bias.data_ptr(),
                                     M, K, N, x.data_ptr(), None, 0.5 * 65536.0 / (65536 - thr), 1, 4, thr, 0, out.data_ptr(), st)
        flush.zero_()
        lib.ob_bwd_prep_fused(gy.data_ptr(), 2, None, y.data_ptr(), 65536.0 / (65536 - thr), 1, 0, thr, 0, s.data_ptr(), q.data_ptr(),
                              M, N, K, dys.data_ptr(), None, colsum.data_ptr(), st)
    torch.cuda.synchronize()


run(25536, 256, 1024)
run(76608, 256, 1024)          # the stacked step's row count (prep, grad_W) for the bench line's roofline.traffic
run(65536, 2048, 2048, big=True)
# small batch (GEMV-like regime): M = 1, 8, 64 at 256 -> 1024
layer = ob.QuantizedLinear(256, 1024).to(dev)
pk, _ = layer.packed_weight(2)
for M in (1, 8, 64):
    q, s = ob.act_quant_int8(torch.randn(M, 256, device=dev))
    for _ in range(2):
        flush.zero_()
        obq.gemm_fwd(q, s, pk, layer.alpha, layer.bias, 1024, torch.float32)
torch.cuda.synchronize()
print("done")
