"""grad_W on CTA pairs reading the int8 codes (ob_bwd_dw_q8) against the single-CTA kernel on the bf16 copy (ob_bwd_dw):
agreement through the C ABI on ragged shapes, then device time per launch (graph replay over rotating operand sets)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import onebit_b200 as ob  # noqa: E402
from onebit_b200 import _cabi  # noqa: E402

lib, check = _cabi.lib, _cabi.check
RAW = 1


def run(M, N, K, bw, seed=0, time_it=False):
    torch.manual_seed(seed)
    layer = ob.QuantizedLinear(K, N).cuda()
    x = torch.randn(M, K, device="cuda")
    q, s = ob.act_quant_int8(x)
    g = torch.randn(M, N, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    dys = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    qb = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
    colsum = torch.empty(lib.ob_bwd_colsum_blocks(M), N, device="cuda")
    check(lib.ob_bwd_prep(g.data_ptr(), 0, s.data_ptr(), q.data_ptr(), M, N, K, dys.data_ptr(), qb.data_ptr(), colsum.data_ptr(), st))
    nbytes = lib.ob_bwd_dw_workspace_bytes(M, N, K)
    ws = torch.empty(nbytes, device="cuda", dtype=torch.uint8)
    out = {}
    for name, fn, src in (("bf16", lib.ob_bwd_dw, qb), ("q8", lib.ob_bwd_dw_q8, q)):
        gw = torch.full((N, K), float("nan"), device="cuda")
        ga = torch.empty((), device="cuda")
        gb = torch.empty(N, device="cuda")
        check(fn(dys.data_ptr(), src.data_ptr(), colsum.data_ptr(), layer.weight.data_ptr(), layer.alpha.data_ptr(), RAW, bw, M, N, K,
                 gw.data_ptr(), ga.data_ptr(), gb.data_ptr(), ws.data_ptr(), nbytes, st))
        torch.cuda.synchronize()
        out[name] = (gw, ga, gb)
    (gw0, ga0, gb0), (gw1, ga1, gb1) = out["bf16"], out["q8"]
    ref = (dys.double().t() @ qb.double())
    a_eff = layer.alpha.abs().double() + 1e-8
    mask = ((layer.weight.double() / a_eff).abs() <= 1.0)
    ref = ref * mask
    scale = ref.abs().max().item() + 1e-30
    e01 = (gw0.double() - gw1.double()).abs().max().item() / scale
    e1r = (gw1.double() - ref).abs().max().item() / scale
    e0r = (gw0.double() - ref).abs().max().item() / scale
    ea = abs(ga0.item() - ga1.item()) / (abs(ga0.item()) + 1e-30)
    eb = (gb0 - gb1).abs().max().item()
    ok = e01 < 2e-5 and e1r < 2e-5 and ea < 1e-4 and eb == 0.0 and not torch.isnan(gw1).any().item()
    print(f"M={M:6d} N={N:5d} K={K:5d} bw={bw}: q8 vs bf16 {e01:.2e}  q8 vs fp64 {e1r:.2e}  bf16 vs fp64 {e0r:.2e}  alpha {ea:.1e}  bias {eb:.1e}  "
          f"{'OK' if ok else 'FAIL'}", flush=True)
    return ok


def timing(M, N, K):
    torch.manual_seed(0)
    layer = ob.QuantizedLinear(K, N).cuda()
    nset = max(2, int(300e6 // (2 * M * N + 2 * M * K)) + 1)
    sets = []
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(nset):
        q, s = ob.act_quant_int8(torch.randn(M, K, device="cuda"))
        dys = torch.randn(M, N, device="cuda").bfloat16()
        sets.append((q, q.to(torch.bfloat16), dys))
    nbytes = lib.ob_bwd_dw_workspace_bytes(M, N, K)
    ws = torch.empty(nbytes, device="cuda", dtype=torch.uint8)
    gw, ga = torch.empty(N, K, device="cuda"), torch.empty((), device="cuda")
    res = {}
    for name in ("bf16", "q8"):
        def call(i):
            q, qb, dys = sets[i % nset]
            fn, src = (lib.ob_bwd_dw, qb) if name == "bf16" else (lib.ob_bwd_dw_q8, q)
            check(fn(dys.data_ptr(), src.data_ptr(), None, layer.weight.data_ptr(), layer.alpha.data_ptr(), RAW, 2, M, N, K,
                     gw.data_ptr(), ga.data_ptr(), None, ws.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream))
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            call(0)
            with torch.cuda.graph(g, stream=side):
                for i in range(20):
                    call(i)
        torch.cuda.current_stream().wait_stream(side)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 60 * 1e3
    by = 2.0 * M * N + M * K + 8.0 * N * K
    print(f"M={M} N={N} K={K}: bf16 {res['bf16']:6.1f} us   q8 pair {res['q8']:6.1f} us  ({by / res['q8'] * 1e-3:6.0f} GB/s on 2MN + MK + 8NK)", flush=True)


if __name__ == "__main__":
    good = True
    for (M, N, K) in [(300, 256, 256), (4100, 1024, 256), (4100, 256, 1024), (1000, 320, 512), (25536, 1024, 256), (25536, 256, 1024),
                      (2049, 576, 320), (64, 256, 256), (9000, 2048, 2048)]:
        for bw in (2, 1):
            good &= run(M, N, K, bw)
    print("agreement:", "OK" if good else "FAIL", flush=True)
    for (M, N, K) in [(25536, 1024, 256), (25536, 256, 1024), (25536, 256, 256), (76608, 1024, 256), (76608, 256, 1024), (76608, 256, 256)]:
        timing(M, N, K)
