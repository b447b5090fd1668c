"""Kernel-by-kernel bring-up checks on a B200 (run through gpurun).  Each case runs in its own
subprocess under a timeout so one hung kernel cannot take the whole call down.

    python tools/gpu_check.py            # all cases
    python tools/gpu_check.py fwd        # one case in-process
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = ["quant", "act", "fwd", "dx", "dw", "layer"]


def _imports():
    import numpy as np
    import torch

    import onebit_b200 as ob
    from onebit_b200 import quant as obq
    from oracle import onebit_oracle as orc
    return np, torch, ob, obq, orc


def case_quant():
    np, torch, ob, obq, orc = _imports()
    dev = "cuda"
    ok = True
    for (N, K) in [(256, 256), (256, 1024), (1024, 256), (192, 128), (2048, 2048)]:
        torch.manual_seed(N + K)
        W = (torch.rand(N, K) * 4 - 2) / (K ** 0.5)
        W[0, :8] = torch.tensor([0.0, -0.0, 0.5, -0.5, 1.0, -1.0, 0.25, 2.0]) * W.abs().mean()
        alpha = -W.abs().mean()
        Wd, ad = W.to(dev), alpha.to(dev).reshape(())
        for bw in (1, 2):
            ref = orc.quant_codes(W.numpy(), orc.alpha_eff(alpha.numpy()), bw)
            packed, packed_t = obq.pack_weight(Wd, ad, bw)
            got = obq.unpack_codes(packed, 0).cpu().numpy()
            got_np = orc.unpack_codes(packed.cpu().numpy(), "i8")
            got_t = obq.unpack_codes(packed_t, 1).cpu().numpy()
            got_t_np = orc.unpack_codes(packed_t.cpu().numpy(), "bf16")
            e1, e2 = (got != ref).sum(), (got_np != ref).sum()
            e3, e4 = (got_t != ref.T).sum(), (got_t_np != ref.T).sum()
            print(f"quant N={N} K={K} bw={bw}: mismatches dev-unpack={e1} np-unpack={e2} T-dev={e3} T-np={e4}")
            ok &= (e1 + e2 + e3 + e4) == 0
        am = obq.weight_absmean(Wd).item()
        print(f"  absmean dev={am:.8f} ref={W.abs().double().mean().item():.8f}")
        ok &= abs(am - W.abs().double().mean().item()) < 1e-6 * abs(am) + 1e-9
        # dense quantiser + STE backward
        a_eff = torch.tensor(float(orc.alpha_eff(alpha.numpy())), device=dev, requires_grad=True)
        Wp = Wd.clone().requires_grad_(True)
        g = torch.randn(N, K, generator=torch.Generator().manual_seed(5))
        for bw in (1, 2):
            Wp.grad = a_eff.grad = None
            what = ob.quantize_weight(Wp, a_eff, bw)
            what.backward(g.to(dev))
            ref_what = orc.quantize_weight(W.numpy(), a_eff.item(), bw)
            gw, ga = orc.ste_backward(g.numpy(), W.numpy(), np.float32(a_eff.item()), bw)
            e_w = (what.detach().cpu().numpy() != ref_what).sum()
            e_g = (Wp.grad.cpu().numpy() != gw).sum()
            rel_a = abs(a_eff.grad.item() - float(ga)) / (abs(float(ga)) + 1e-9)
            print(f"  dense bw={bw}: W_hat mismatches={e_w} grad_W mismatches={e_g} grad_alpha rel err={rel_a:.2e}")
            ok &= e_w == 0 and e_g == 0 and rel_a < 1e-4
    return ok


def case_act():
    np, torch, ob, obq, orc = _imports()
    ok = True
    for (M, K) in [(111, 128), (996, 256), (1000, 512), (513, 1024), (300, 2048), (77, 320), (4096, 256)]:
        g = torch.Generator().manual_seed(M * 7 + K)
        x = torch.randn(M, K, generator=g)
        x[0] = 0.0
        x[1, 3] = 55.0
        x[2] = x[2] * 1e-7
        for dt in (torch.float32, torch.bfloat16):
            xd = x.to(dt)
            q, s = ob.act_quant_int8(xd.cuda())
            q_ref, s_ref = orc.act_quant(xd.float().numpy())
            eq = (q.cpu().numpy() != q_ref).sum()
            es = (s.cpu().numpy() != s_ref).sum()
            print(f"act M={M} K={K} {dt}: code mismatches={eq} scale mismatches={es}")
            ok &= eq == 0 and es == 0
    return ok


def _int_ref(torch, q, Q):
    return q.double() @ Q.double().t()


def case_fwd():
    np, torch, ob, obq, orc = _imports()
    from onebit_b200 import _cabi
    dev = "cuda"
    ok = True
    shapes = [(128, 64, 128), (128, 256, 256), (300, 256, 256), (996, 1024, 256), (996, 256, 1024), (2048, 512, 512),
              (4096, 2048, 2048), (777, 192, 320)]
    for (M, N, K) in shapes:
        for bn in (0, 64, 128, 256, 1128, 1256):
            _cabi.lib.ob_debug_set(_cabi.DBG_FORCE_BLOCK_N, bn)
            g = torch.Generator().manual_seed(M + N + K)
            q = torch.randint(-127, 128, (M, K), generator=g, dtype=torch.int8).to(dev)
            scale = (torch.rand(M, generator=g) * 50 + 10).to(dev)
            codes = torch.randint(-1, 2, (N, K), generator=g, dtype=torch.int8)
            packed = torch.from_numpy(orc.pack_codes(codes.numpy(), "i8")).to(dev)
            alpha = torch.tensor(0.0625, device=dev)
            bias = torch.randn(N, generator=g).to(dev)
            ref = _int_ref(torch, q, codes.to(dev)) * (alpha.double() / scale.double())[:, None] + bias.double()
            for odt in (torch.float32, torch.bfloat16):
                y = obq.gemm_fwd(q, scale, packed, alpha, bias, N, odt, _cabi.OB_ALPHA_EFF)
                torch.cuda.synchronize()
                err = (y.double() - ref).abs()
                tol = 1e-5 * ref.abs().max().item() if odt == torch.float32 else 1e-2 * ref.abs().max().item()
                bad = (err > tol)
                nbad = int(bad.sum())
                print(f"fwd M={M} N={N} K={K} bn={bn} {str(odt)[6:]}: max err={err.max().item():.3e} "
                      f"(tol {tol:.1e}) bad={nbad}")
                if nbad:
                    ok = False
                    rows = bad.any(1).nonzero().flatten()[:8].tolist()
                    cols = bad.any(0).nonzero().flatten()[:8].tolist()
                    print(f"   first bad rows {rows} cols {cols}; y[0,:4]={y[0, :4].tolist()} ref={ref[0, :4].tolist()}")
    _cabi.lib.ob_debug_set(_cabi.DBG_FORCE_BLOCK_N, 0)
    return ok


def case_dx():
    np, torch, ob, obq, orc = _imports()
    from onebit_b200 import _cabi
    lib = _cabi.lib
    dev = "cuda"
    ok = True
    for (M, N, K) in [(128, 64, 64), (128, 256, 256), (300, 256, 256), (996, 1024, 256), (996, 256, 1024),
                      (2048, 512, 512), (777, 192, 320)]:
        for bn in (0, 64, 128, 256, 1128, 1256):
            lib.ob_debug_set(_cabi.DBG_FORCE_BLOCK_N, bn)
            g = torch.Generator().manual_seed(M + N + K + 1)
            dys = torch.randn(M, N, generator=g).to(torch.bfloat16).to(dev)
            scale = (torch.rand(M, generator=g) * 50 + 10).to(dev)
            codes = torch.randint(-1, 2, (N, K), generator=g, dtype=torch.int8)
            packed_t = torch.from_numpy(orc.pack_codes(np.ascontiguousarray(codes.numpy().T), "bf16")).to(dev)
            alpha = torch.tensor(0.0625, device=dev)
            ref = (dys.double() @ codes.to(dev).double()) * (alpha.double() * scale.double())[:, None]
            dx = torch.empty(M, K, device=dev)
            _cabi.check(lib.ob_bwd_dx(dys.data_ptr(), scale.data_ptr(), packed_t.data_ptr(), alpha.data_ptr(),
                                      _cabi.OB_ALPHA_EFF, M, N, K, dx.data_ptr(), _cabi.OB_F32,
                                      torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            err = (dx.double() - ref).abs()
            tol = 1e-5 * ref.abs().max().item()
            nbad = int((err > tol).sum())
            print(f"dx M={M} N={N} K={K} bn={bn}: max err={err.max().item():.3e} (tol {tol:.1e}) bad={nbad}")
            ok &= nbad == 0
    lib.ob_debug_set(_cabi.DBG_FORCE_BLOCK_N, 0)
    return ok


def case_dw():
    np, torch, ob, obq, orc = _imports()
    from onebit_b200 import _cabi
    lib = _cabi.lib
    dev = "cuda"
    results = {}
    for swap in (0, 1):
        lib.ob_debug_set(_cabi.DBG_SWAP_LBO_SBO, swap)
        ok = True
        for (M, N, K) in [(64, 128, 64), (128, 128, 256), (300, 256, 256), (996, 1024, 256), (996, 256, 1024),
                          (5000, 512, 512), (777, 192, 320)]:
            for splits in (0, 1, 3):
                lib.ob_debug_set(_cabi.DBG_FORCE_SPLITS, splits)
                g = torch.Generator().manual_seed(M + N + K + 2)
                dys = torch.randn(M, N, generator=g).to(torch.bfloat16).to(dev)
                qb = torch.randint(-127, 128, (M, K), generator=g).to(torch.bfloat16).to(dev)
                W = ((torch.rand(N, K, generator=g) * 4 - 2) / K ** 0.5).to(dev)
                alpha = W.abs().mean().reshape(())
                colsum = torch.randn(lib.ob_bwd_colsum_blocks(M), N, generator=g).to(dev)
                gw = torch.empty(N, K, device=dev)
                ga = torch.empty((), device=dev)
                gb = torch.empty(N, device=dev)
                nbytes = lib.ob_bwd_dw_workspace_bytes(M, N, K)
                ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
                _cabi.check(lib.ob_bwd_dw(dys.data_ptr(), qb.data_ptr(), colsum.data_ptr(), W.data_ptr(),
                                          alpha.data_ptr(), _cabi.OB_ALPHA_RAW, 2, M, N, K, gw.data_ptr(),
                                          ga.data_ptr(), gb.data_ptr(), ws.data_ptr(), nbytes,
                                          torch.cuda.current_stream().cuda_stream))
                torch.cuda.synchronize()
                g_hat = (dys.double().t() @ qb.double()).float().cpu().numpy()
                gw_ref, ga_ref = orc.ste_backward(g_hat, W.cpu().numpy(), orc.alpha_eff(alpha.item()), 2)
                err = np.abs(gw.cpu().numpy() - gw_ref)
                tol = 1e-4 * np.abs(gw_ref).max()
                nbad = int((err > tol).sum())
                rel_a = abs(ga.item() - float(ga_ref)) / (abs(float(ga_ref)) + 1e-9)
                eb = (gb.double() - colsum.double().sum(0)).abs().max().item()
                print(f"dw swap={swap} M={M} N={N} K={K} splits={splits}: max err={err.max():.3e} (tol {tol:.1e}) "
                      f"bad={nbad} galpha rel={rel_a:.2e} gbias err={eb:.2e}")
                ok &= nbad == 0 and rel_a < 1e-3 and eb < 1e-3
        results[swap] = ok
    lib.ob_debug_set(_cabi.DBG_SWAP_LBO_SBO, 0)
    lib.ob_debug_set(_cabi.DBG_FORCE_SPLITS, 0)
    print("dw summary: lbo/sbo as designed ->", results[0], "; swapped ->", results[1])
    return results[0]


def case_layer():
    import __graft_entry__ as ge
    ge.smoke()
    return True


def main():
    if len(sys.argv) > 1:
        ok = globals()["case_" + sys.argv[1]]()
        print(f"CASE {sys.argv[1]}: {'PASS' if ok else 'FAIL'}")
        sys.exit(0 if ok else 1)
    summary = {}
    for c in CASES:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), c], timeout=240)
            summary[c] = "PASS" if r.returncode == 0 else f"FAIL(rc={r.returncode})"
        except subprocess.TimeoutExpired:
            summary[c] = "TIMEOUT"
        print(f"=== {c}: {summary[c]} ({time.time() - t0:.0f}s)", flush=True)
    print("SUMMARY", summary)


if __name__ == "__main__":
    main()
