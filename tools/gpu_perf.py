"""Quick kernel timings on a B200 (CUDA events, L2-exceeding or rotating buffers).  python tools/gpu_perf.py [fwd|bwd|all]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import onebit_b200 as ob  # noqa: E402
from onebit_b200 import _cabi, quant as obq  # noqa: E402

lib = _cabi.lib
PEAK_HBM = 6545.6


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def fwd(shapes, out_dtypes=(torch.bfloat16, torch.float32), bns=(0,)):
    for (M, K, N) in shapes:
        torch.manual_seed(0)
        layer = ob.QuantizedLinear(K, N).cuda()
        pk, pkt = layer.packed_weight(2)
        nbuf = max(1, int(300e6 // (M * K + 2 * M * N)))     # rotate buffers so the working set exceeds L2
        xs = [torch.randn(M, K, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
        qs = [ob.act_quant_int8(x) for x in xs]
        for odt in out_dtypes:
            for bn in bns:
                lib.ob_debug_set(_cabi.DBG_FORCE_BLOCK_N, bn)
                i = [0]

                def f():
                    q, s = qs[i[0] % nbuf]
                    i[0] += 1
                    obq.gemm_fwd(q, s, pk, layer.alpha, layer.bias, N, odt)
                us = timeit(f)
                ob_ = 2 if odt == torch.bfloat16 else 4
                by = M * K + N * K / 4 + ob_ * M * N + 4 * M + 4 * N
                print(f"fwd M={M:6d} K={K:5d} N={N:5d} out={str(odt)[6:]:8s} bn={bn:3d}: {us:8.1f} us  "
                      f"{2.0 * M * N * K / us / 1e6:7.1f} TOPS  {by / us / 1e3:7.1f} GB/s ({by / us / 1e3 / PEAK_HBM * 100:4.1f}% hbm)",
                      flush=True)
        lib.ob_debug_set(_cabi.DBG_FORCE_BLOCK_N, 0)
        i = [0]

        def fa():
            i[0] += 1
            ob.act_quant_int8(xs[i[0] % nbuf])
        us = timeit(fa)
        by = M * K * 3 + 4 * M
        print(f"act M={M:6d} K={K:5d} (bf16 in): {us:8.1f} us  {by / us / 1e3:7.1f} GB/s ({by / us / 1e3 / PEAK_HBM * 100:4.1f}% hbm)")


def bwd(shapes):
    st = lambda: torch.cuda.current_stream().cuda_stream
    for (M, K, N) in shapes:
        torch.manual_seed(0)
        layer = ob.QuantizedLinear(K, N).cuda()
        pk, pkt = layer.packed_weight(2)
        x = torch.randn(M, K, device="cuda")
        q, s = ob.act_quant_int8(x)
        gy = torch.randn(M, N, device="cuda")
        dys = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        qb = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
        colsum = torch.empty(lib.ob_bwd_colsum_blocks(M), N, device="cuda")
        dx = torch.empty(M, K, device="cuda")
        gw = torch.empty(N, K, device="cuda")
        ga = torch.empty((), device="cuda")
        gb = torch.empty(N, device="cuda")
        nbytes = lib.ob_bwd_dw_workspace_bytes(M, N, K)
        ws = torch.empty(nbytes, device="cuda", dtype=torch.uint8)
        f_prep = lambda: lib.ob_bwd_prep(gy.data_ptr(), 0, s.data_ptr(), q.data_ptr(), M, N, K, dys.data_ptr(),
                                         qb.data_ptr(), colsum.data_ptr(), st())
        f_dx = lambda: lib.ob_bwd_dx(dys.data_ptr(), s.data_ptr(), pkt.data_ptr(), layer.alpha.data_ptr(), 1, M, N, K,
                                     dx.data_ptr(), 0, st())
        f_dw = lambda: lib.ob_bwd_dw(dys.data_ptr(), qb.data_ptr(), colsum.data_ptr(), layer.weight.data_ptr(),
                                     layer.alpha.data_ptr(), 1, 2, M, N, K, gw.data_ptr(), ga.data_ptr(), gb.data_ptr(),
                                     ws.data_ptr(), nbytes, st())
        t_prep, t_dx, t_dw = timeit(f_prep), timeit(f_dx), timeit(f_dw)
        by_prep = M * N * 6 + M * K * 3
        by_dx = M * N * 2 + N * K / 4 + M * K * 4
        by_dw = M * N * 2 + M * K * 2 + 8 * N * K
        fl = 2.0 * M * N * K
        print(f"bwd M={M:6d} K={K:5d} N={N:5d}: prep {t_prep:7.1f} us ({by_prep / t_prep / 1e3:6.0f} GB/s) | "
              f"dx {t_dx:7.1f} us ({fl / t_dx / 1e6:6.1f} TF, {by_dx / t_dx / 1e3:6.0f} GB/s) | "
              f"dw+ste {t_dw:7.1f} us ({fl / t_dw / 1e6:6.1f} TF, {by_dw / t_dw / 1e3:6.0f} GB/s)", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    model = [(25536, 256, 256), (25536, 256, 1024), (25536, 1024, 256), (399, 256, 256)]
    big = [(65536, 2048, 2048), (65536, 1024, 1024), (65536, 256, 256), (4096, 2048, 2048)]
    if what in ("fwd", "all"):
        fwd(big + model, bns=(0, 256, 1256, 1128) if what == "fwd" else (0,))
    if what in ("bwd", "all"):
        bwd(model + [(65536, 2048, 2048)])


def kern(shapes):
    """Per-kernel device times (CUPTI via torch.profiler) of one layer forward+backward at each shape."""
    from torch.profiler import ProfilerActivity, profile
    for (M, K, N) in shapes:
        torch.manual_seed(0)
        layer = ob.QuantizedLinear(K, N).cuda()
        x = torch.randn(M, K, device="cuda", requires_grad=True)
        gy = torch.randn(M, N, device="cuda")
        for _ in range(3):
            obq._ActQuantCache.clear()
            layer(x, 2).backward(gy)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(10):
                obq._ActQuantCache.clear()
                layer(x, 2).backward(gy)
            torch.cuda.synchronize()
        print(f"--- layer fwd+bwd M={M} K={K} N={N}")
        tot = 0.0
        for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
            if e.device_time_total > 0:
                tot += e.device_time_total / 10
                print(f"   {e.device_time_total / e.count:9.1f} us x{e.count // 10:2d}  {e.key[:110]}")
        print(f"   total {tot:.1f} us per fwd+bwd", flush=True)


if len(sys.argv) > 1 and sys.argv[1] == "kern":
    kern([(25536, 256, 256), (25536, 256, 1024), (25536, 1024, 256), (399, 256, 256)])


def modelshape():
    """A few launches of each layer kernel at the model's widest routed shape (for ncu captures)."""
    M, K, N = 25536, 256, 1024
    torch.manual_seed(0)
    layer = ob.QuantizedLinear(K, N).cuda()
    x = torch.randn(M, K, device="cuda", requires_grad=True)
    gy = torch.randn(M, N, device="cuda")
    for _ in range(4):
        obq._ActQuantCache.clear()
        layer(x, 2).backward(gy)
    torch.cuda.synchronize()
    print("modelshape done")


if len(sys.argv) > 1 and sys.argv[1] == "modelshape":
    modelshape()


def attn():
    """The fused attention chain at the bench shape (B=64, H=4, T=399), forward + backward, a few launches."""
    from onebit_b200.attention import rel_attention_probs
    B, H, T = 64, 4, 399
    ac = torch.randn(B, H, T, T, device="cuda", requires_grad=True)
    bd = torch.randn(B, H, T, T, device="cuda", requires_grad=True)
    mask = torch.ones(B, T, T, device="cuda", dtype=torch.bool)
    gy = torch.randn(B, H, T, T, device="cuda")
    for _ in range(3):
        rel_attention_probs(ac, bd, mask, 0.125, 0.1, True).backward(gy)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    out = rel_attention_probs(ac, bd, mask, 0.125, 0.1, True)
    e[1].record()
    out.backward(gy)
    e[2].record()
    torch.cuda.synchronize()
    nb = B * H * T * T
    print(f"attn chain fwd(+bernoulli) {e[0].elapsed_time(e[1]) * 1e3:.0f} us, bwd {e[1].elapsed_time(e[2]) * 1e3:.0f} us "
          f"({nb * 18 / 1e6:.0f} MB each way)")


if len(sys.argv) > 1 and sys.argv[1] == "attn":
    attn()
