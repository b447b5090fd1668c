"""Generate the golden fixtures by EXECUTING the unmodified reference on CPU.

Run in the build container only (needs /root/reference, which is not present on the GPU box):

    python tests/golden/make_golden.py

Writes tests/golden/{kat_edges.json,kat_seeded.json,kat_layer.npz,kat_layer_stats.json,kat_decode.npz}.
Nothing from the reference is copied: it is imported from where it lies and only its
numerical outputs are stored.  Recipes follow SURVEY.md section 8(c).
"""
import hashlib
import json
import os
import struct
import sys

import numpy as np
import torch

REF = os.environ.get("ONEBIT_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(REF, "onebit_asr"), REF]
import quant as refq  # noqa: E402  (the reference's onebit_asr/quant.py)

HERE = os.path.dirname(os.path.abspath(__file__))


def f32bits(v) -> str:
    return "0x%08x" % struct.unpack("<I", struct.pack("<f", float(v)))[0]


def sha16(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def ref_codes(W, a_eff, bw):
    what = refq.quantize_weight(W, a_eff, bw)
    return torch.round(what / a_eff).to(torch.int8).numpy()


def kat_edges():
    """KAT-1: ties and window edges, alpha = 1."""
    W = torch.tensor([0.5, -0.5, 0.0, 1.0, -1.0, 1.5, 0.49999997, 0.25, -0.75, -0.0])
    g = torch.arange(1, 11, dtype=torch.float32)
    out = {"W": W.tolist(), "g": g.tolist(), "alpha": 1.0}
    for bw in (1, 2):
        Wp = W.clone().requires_grad_(True)
        a = torch.tensor(1.0, requires_grad=True)
        what = refq.quantize_weight(Wp, a, bw)
        what.backward(g)
        out[f"bw{bw}"] = {"W_hat": what.detach().tolist(), "grad_W": Wp.grad.tolist(),
                          "grad_alpha": float(a.grad)}
    # extra: a non-unit alpha with exact ties (|W/alpha| == 0.5 and == 1.0)
    W2 = torch.tensor([0.125, -0.125, 0.25, -0.25, 0.375, 0.1249999, -0.2500001, 0.0, 3.0, -3.0])
    out["alpha2"] = 0.25
    out["W2"] = W2.tolist()
    for bw in (1, 2):
        Wp = W2.clone().requires_grad_(True)
        a = torch.tensor(0.25, requires_grad=True)
        what = refq.quantize_weight(Wp, a, bw)
        what.backward(g)
        out[f"alpha2_bw{bw}"] = {"W_hat": what.detach().tolist(), "grad_W": Wp.grad.tolist(),
                                 "grad_alpha": float(a.grad)}
    return out


def kat_seeded():
    """KAT-2: seeded layer construction -> alpha bits and code hashes."""
    out = {}
    for (K, N) in [(256, 1024), (256, 256), (1024, 256), (2048, 2048), (512, 512)]:
        torch.manual_seed(0)
        ql = refq.QuantizedLinear(K, N)
        a_eff = (ql.alpha.abs() + 1e-8).detach()
        W = ql.weight.detach()
        q2, q1 = ref_codes(W, a_eff, 2), ref_codes(W, a_eff, 1)
        out[f"{K}x{N}"] = {
            "in": K, "out": N,
            "alpha_bits": f32bits(ql.alpha), "alpha_eff_bits": f32bits(a_eff),
            "sum_W_bits": f32bits(W.double().sum().float()),
            "sha_W": sha16(W.numpy()),
            "sum_q2": int(q2.astype(np.int64).sum()), "nnz_q2": int((q2 != 0).sum()), "sha_q2": sha16(q2),
            "sum_q1": int(q1.astype(np.int64).sum()), "sha_q1": sha16(q1),
            "ste_in_window": int(((W / a_eff).abs() <= 1.0).sum()),
        }
    return out


def act_quant_ref(x):
    """Oracle-B recipe written once more with torch ops."""
    s = 127.0 / x.abs().amax(-1, keepdim=True).clamp(min=1e-5)
    q = (x * s).round().clamp(-128, 127)
    return q, s


def run_ref_layer(ql, x, gy, bw, act8):
    for p in ql.parameters():
        p.grad = None
    xr = x.clone().requires_grad_(True)
    if act8 and bw != 32:
        q, s = act_quant_ref(xr)
        xin = xr + (q / s - xr).detach()
    else:
        xin = xr
    y = ql(xin, bw)
    y.backward(gy)
    return {"y": y.detach().numpy(), "gx": xr.grad.numpy().copy(),
            "gW": ql.weight.grad.numpy().copy(),
            "galpha": np.float32(0.0 if ql.alpha.grad is None else ql.alpha.grad.item()),
            "gb": ql.bias.grad.numpy().copy()}


def kat_layer_small():
    """Full tensors of a small ragged layer problem: in=128, out=192, x [3,37,128]."""
    torch.manual_seed(7)
    ql = refq.QuantizedLinear(128, 192)
    with torch.no_grad():
        ql.bias.copy_(torch.randn(192) * 0.1)
        ql.alpha.mul_(-1.0)            # negative alpha exercises |alpha| and sign(alpha)
    g = torch.Generator().manual_seed(99)
    x = torch.randn(3, 37, 128, generator=g)
    x[0, 0] = 0.0                      # an all-zero token: amax clamp path
    x[1, 5, 7] = 40.0                  # an outlier token
    gy = torch.randn(3, 37, 192, generator=g)
    out = {"W": ql.weight.detach().numpy().copy(), "alpha": np.float32(ql.alpha.item()),
           "bias": ql.bias.detach().numpy().copy(), "x": x.numpy(), "gy": gy.numpy()}
    q, s = act_quant_ref(x)
    out["act_q"] = q.to(torch.int8).numpy()
    out["act_s"] = s.squeeze(-1).numpy()
    a_eff = (ql.alpha.abs() + 1e-8).detach()
    for bw in (1, 2):
        out[f"codes_bw{bw}"] = ref_codes(ql.weight.detach(), a_eff, bw)
    for bw in (1, 2, 32):
        for act8 in (False, True):
            if bw == 32 and act8:
                continue
            r = run_ref_layer(ql, x, gy, bw, act8)
            tag = f"bw{bw}_{'B' if act8 else 'A'}"
            for k, v in r.items():
                out[f"{tag}_{k}"] = v
    return out


def kat_layer_stats():
    """KAT-3: summary statistics of the seed-0 (256->1024) layer at x [4,249,256]."""
    torch.manual_seed(0)
    ql = refq.QuantizedLinear(256, 1024)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(4, 249, 256, generator=g)
    gy = torch.randn(4, 249, 1024, generator=g)
    out = {}
    q, s = act_quant_ref(x)
    out["act"] = {"sha_q": sha16(q.to(torch.int8).numpy()), "sha_s": sha16(s.squeeze(-1).numpy()),
                  "q_min": int(q.min()), "q_max": int(q.max()), "s00_bits": f32bits(s[0, 0, 0])}
    for bw in (1, 2):
        for act8 in (False, True):
            r = run_ref_layer(ql, x, gy, bw, act8)
            out[f"bw{bw}_{'B' if act8 else 'A'}"] = {
                "sum_y": float(r["y"].astype(np.float64).sum()),
                "mean_abs_y": float(np.abs(r["y"]).mean()),
                "mean_abs_gx": float(np.abs(r["gx"]).mean()),
                "mean_abs_gW": float(np.abs(r["gW"]).mean()),
                "nnz_frac_gW": float((r["gW"] != 0).mean()),
                "galpha": float(r["galpha"]),
                "sum_gb": float(r["gb"].astype(np.float64).sum()),
            }
    return out


def kat_decode():
    """Greedy CTC decode of the reference (onebit_asr/metrics.py:51-60) on logits with ties, blanks and repeats."""
    import metrics as refm          # reference onebit_asr/metrics.py
    g = torch.Generator().manual_seed(77)
    B, T, V = 4, 61, 40
    logits = torch.randn(B, T, V, generator=g)
    steer = torch.randint(0, 8, (B, T), generator=g)                 # small alphabet -> many repeats and blanks (id 3)
    logits.scatter_(2, steer.unsqueeze(-1), 6.0)
    logits[0, 5, 9] = logits[0, 5, 2] = 9.0                           # exact tie -> first maximal index
    logits[1, 0, 3] = 12.0                                            # starts with a blank
    lens = torch.tensor([61, 40, 1, 17])
    hyps = [refm.ctc_greedy_decode(logits[b, : int(lens[b])], blank_id=3) for b in range(B)]
    pad = np.full((B, T), -1, dtype=np.int32)
    for b, h in enumerate(hyps):
        pad[b, : len(h)] = h
    return {"logits": logits.numpy(), "lens": lens.numpy().astype(np.int32), "tokens": pad,
            "out_lens": np.array([len(h) for h in hyps], dtype=np.int32)}


def main():
    torch.set_num_threads(1)           # fixed summation order for the stored float sums
    with open(os.path.join(HERE, "kat_edges.json"), "w") as f:
        json.dump(kat_edges(), f, indent=1)
    with open(os.path.join(HERE, "kat_seeded.json"), "w") as f:
        json.dump(kat_seeded(), f, indent=1)
    np.savez_compressed(os.path.join(HERE, "kat_layer.npz"), **kat_layer_small())
    with open(os.path.join(HERE, "kat_layer_stats.json"), "w") as f:
        json.dump(kat_layer_stats(), f, indent=1)
    np.savez_compressed(os.path.join(HERE, "kat_decode.npz"), **kat_decode())
    print("golden fixtures written to", HERE, "torch", torch.__version__)


if __name__ == "__main__":
    main()
