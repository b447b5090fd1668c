"""CPU oracle for the quantised-linear hot path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (numpy for the code/integer work, torch-CPU for
the differentiable module) of the algorithm in the reference's
``onebit_asr/quant.py``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package never does (it fails loudly when ``libonebit.so`` is missing).

Parity status
-------------
* Oracle-A (pure reference semantics: learnable-scale binary/ternary weight
  quantiser, clip-window STE, custom d/d-alpha, fp32 activations) is PINNED: the
  reference ships no golden vectors, so
  ``tests/golden/make_golden.py`` executes the unmodified reference on CPU in
  the build container and commits its outputs; ``tests/test_oracle.py`` checks
  this restatement against those fixtures bit-exactly (codes, masks) or to
  fp32 round-off (sums).
* Oracle-B (Oracle-A preceded by the BitNet-b1.58 per-token absmax int8
  activation quantiser) has NO counterpart in the reference (SURVEY.md section 0,
  row D4): for the int8 activation codes/scales the parity is "unpinned by the
  reference"; it is our published spec, restated here, and its fixtures are
  produced by wrapping the real reference layer (see make_golden.py).

Reference lines followed (relative to /root/reference):
  quantiser forward   onebit_asr/quant.py:45-70
  STE / d-alpha       onebit_asr/quant.py:72-92
  layer forward       onebit_asr/quant.py:120-127
  layer init          onebit_asr/quant.py:100-118
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32

# --------------------------------------------------------------------------
# weight quantiser (Oracle-A)
# --------------------------------------------------------------------------


def alpha_eff(alpha) -> np.float32:
    """Effective scale used by the layer: |alpha| + 1e-8 in fp32 (quant.py:124)."""
    return F32(np.abs(F32(alpha)) + F32(1e-8))


def scaled_weight(W: np.ndarray, a_eff) -> np.ndarray:
    """Wa = W / alpha, IEEE fp32 division (quant.py:49)."""
    return (np.asarray(W, dtype=F32) / F32(a_eff)).astype(F32)


def quant_codes(W: np.ndarray, a_eff, bitwidth: int) -> np.ndarray:
    """Integer codes Q in {-1,0,+1} (bitwidth 2) or {-1,+1} (bitwidth 1), int8.

    bitwidth 1: sign of the clipped Wa with zeros mapped to +1 (quant.py:52-55).
    bitwidth 2: 0 where |clip(Wa)| < 0.5 (strict), else the sign (quant.py:56-60).
    A NaN Wa keeps no code in the reference (sign(NaN)=NaN); we do not model NaNs.
    """
    if bitwidth not in (1, 2):
        raise ValueError("bitwidth must be one of {1,2,32}")
    wa = np.clip(scaled_weight(W, a_eff), F32(-1.0), F32(1.0))
    sgn = np.sign(wa).astype(np.int8)
    if bitwidth == 1:
        return np.where(sgn == 0, np.int8(1), sgn).astype(np.int8)
    return np.where(np.abs(wa) < F32(0.5), np.int8(0), sgn).astype(np.int8)


def quantize_weight(W: np.ndarray, a_eff, bitwidth: int) -> np.ndarray:
    """W_hat = alpha * Q in fp32 (quant.py:68); bitwidth 32 is a passthrough (quant.py:61-64)."""
    if bitwidth == 32:
        return np.asarray(W, dtype=F32)
    return (F32(a_eff) * quant_codes(W, a_eff, bitwidth).astype(F32)).astype(F32)


def ste_mask(W: np.ndarray, a_eff) -> np.ndarray:
    """Clip-window indicator 1[|Wa| <= 1] (quant.py:81)."""
    return np.abs(scaled_weight(W, a_eff)) <= F32(1.0)


def alpha_term(W: np.ndarray, a_eff, bitwidth: int) -> np.ndarray:
    """d W_hat / d alpha per element (quant.py:86-90).

    inside the window (|Wa| < 1, strict):  -Wa + Pi(Wa)
         Pi = sign(Wa)*1[|Wa| >= 0.5] for 2 bits, sign(Wa) for 1 bit (sign(0)=0 here,
         unlike the forward's 0 -> +1 convention);
    outside:                               sign(Wa).
    """
    wa = scaled_weight(W, a_eff)
    sgn = np.sign(wa).astype(F32)
    if bitwidth == 2:
        proj = np.where(np.abs(wa) >= F32(0.5), sgn, F32(0.0))
    else:
        proj = sgn
    inner = (-wa + proj).astype(F32)
    return np.where(np.abs(wa) < F32(1.0), inner, sgn).astype(F32)


def ste_backward(g_what: np.ndarray, W: np.ndarray, a_eff, bitwidth: int):
    """(grad_W, grad_alpha_eff) from the gradient w.r.t. W_hat (quant.py:72-92).

    The alpha reduction is done in float64 and rounded once, so callers compare
    with a tolerance (torch sums in fp32 with its own order)."""
    g = np.asarray(g_what, dtype=F32)
    if bitwidth == 32:
        return g.copy(), F32(0.0)
    gw = (g * ste_mask(W, a_eff).astype(F32)).astype(F32)
    ga = np.sum(g.astype(np.float64) * alpha_term(W, a_eff, bitwidth).astype(np.float64))
    return gw, F32(ga)


# --------------------------------------------------------------------------
# activation quantiser (Oracle-B; BitNet b1.58 per-token absmax int8)
# --------------------------------------------------------------------------


def act_quant(x: np.ndarray):
    """Per-row absmax int8 codes and fp32 scale  s = (1 / max(amax, 1e-5)) * 127.

    The spec expression is torch's ``127.0 / amax.clamp(min=1e-5)``; torch evaluates
    ``scalar / tensor`` as ``tensor.reciprocal() * scalar`` - two fp32 roundings, not one
    division - and the int8 codes/scales are pinned bit-exactly to that (the fixtures are
    produced by that torch expression), so it is restated literally here.
    q = clamp(round_half_even(x * s), -128, 127);  x_tilde = q / s.
    Returns (q int8 [..., K], s fp32 [...])."""
    x = np.asarray(x, dtype=F32)
    amax = np.maximum(np.max(np.abs(x), axis=-1), F32(1e-5)).astype(F32)
    s = ((F32(1.0) / amax).astype(F32) * F32(127.0)).astype(F32)
    q = np.clip(np.rint((x * s[..., None]).astype(F32)), -128, 127).astype(np.int8)
    return q, s


def act_dequant(q: np.ndarray, s: np.ndarray) -> np.ndarray:
    return (q.astype(F32) / s[..., None].astype(F32)).astype(F32)


# --------------------------------------------------------------------------
# layer forward / backward (numpy; float64 accumulation as the "exact" answer)
# --------------------------------------------------------------------------


def linear_forward(x, W, alpha, bias, bitwidth: int, act_bits: int = 8) -> np.ndarray:
    """y = x_used @ W_hat^T + b  (quant.py:120-127); act_bits=8 inserts Oracle-B, 32 keeps x."""
    x = np.asarray(x, dtype=F32)
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    if bitwidth == 32:
        w_used, xu = np.asarray(W, dtype=F32), x2
    else:
        w_used = quantize_weight(W, alpha_eff(alpha), bitwidth)
        xu = act_dequant(*act_quant(x2)) if act_bits == 8 else x2
    y = xu.astype(np.float64) @ w_used.astype(np.float64).T
    if bias is not None:
        y = y + np.asarray(bias, dtype=np.float64)
    return y.astype(F32).reshape(*lead, W.shape[0])


def linear_backward(gy, x, W, alpha, bias, bitwidth: int, act_bits: int = 8):
    """Gradients of sum(y*gy): dict(x, weight, alpha, bias).

    grad_x = gy @ W_hat (identity STE through the activation quantiser),
    grad_What = gy^T @ x_used, then the weight STE; grad_alpha carries
    d|alpha|/d alpha = sign(alpha) (autograd of quant.py:124)."""
    gy = np.asarray(gy, dtype=F32)
    x = np.asarray(x, dtype=F32)
    g2 = gy.reshape(-1, gy.shape[-1]).astype(np.float64)
    x2 = x.reshape(-1, x.shape[-1])
    out = {"bias": None if bias is None else g2.sum(0).astype(F32)}
    if bitwidth == 32:
        out["x"] = (g2 @ np.asarray(W, np.float64)).astype(F32).reshape(x.shape)
        out["weight"] = (g2.T @ x2.astype(np.float64)).astype(F32)
        out["alpha"] = None
        return out
    a = alpha_eff(alpha)
    what = quantize_weight(W, a, bitwidth).astype(np.float64)
    xu = act_dequant(*act_quant(x2)) if act_bits == 8 else x2
    out["x"] = (g2 @ what).astype(F32).reshape(x.shape)
    g_what = (g2.T @ xu.astype(np.float64)).astype(F32)
    gw, ga = ste_backward(g_what, W, a, bitwidth)
    out["weight"] = gw
    out["alpha"] = F32(np.sign(F32(alpha)) * ga)
    return out


# --------------------------------------------------------------------------
# packed 2-bit storage formats (OUR format; the reference has none) restated on CPU
# --------------------------------------------------------------------------
# A field is 2 bits: 0b00 -> 0, 0b10 -> -1, 0b11 -> +1 (0b01 never written, decodes to 0).
# 16 consecutive codes along the contraction axis share one little-endian 32-bit word;
# code t (0..15) of the group sits at bit FIELD_POS_*[t].  Two orders exist so that each
# GEMM can expand a word with one byte-permute per output word:
#   "i8"   (forward GEMM, int8 operand):  pos = 4*(t&3) + 2*((t>>2)&1) + 16*(t>>3)
#   "bf16" (grad_x GEMM, bf16 operand):   pos = 8*(t&1) + 2*((t>>1)&3) + 16*(t>>3)

FIELD_POS_I8 = np.array([4 * (t & 3) + 2 * ((t >> 2) & 1) + 16 * (t >> 3) for t in range(16)])
FIELD_POS_BF16 = np.array([8 * (t & 1) + 2 * ((t >> 1) & 3) + 16 * (t >> 3) for t in range(16)])
_ENC = np.array([2, 0, 3], dtype=np.uint32)          # code+1 -> field
_DEC = np.array([0, 0, -1, 1], dtype=np.int8)         # field -> code


def pack_codes(Q: np.ndarray, order: str = "i8") -> np.ndarray:
    """Pack int8 codes [R, C] (C % 16 == 0) to uint8 [R, C/4] along the last axis."""
    pos = FIELD_POS_I8 if order == "i8" else FIELD_POS_BF16
    R, C = Q.shape
    assert C % 16 == 0
    f = _ENC[(Q.astype(np.int64) + 1)].reshape(R, C // 16, 16)
    words = np.zeros((R, C // 16), dtype=np.uint32)
    for t in range(16):
        words |= f[:, :, t].astype(np.uint32) << np.uint32(pos[t])
    return words.view(np.uint8).reshape(R, C // 4)


def unpack_codes(P: np.ndarray, order: str = "i8") -> np.ndarray:
    """Inverse of pack_codes: uint8 [R, C/4] -> int8 [R, C]."""
    pos = FIELD_POS_I8 if order == "i8" else FIELD_POS_BF16
    R, B = P.shape
    words = np.ascontiguousarray(P).view(np.uint32).reshape(R, B // 4)
    out = np.empty((R, B // 4, 16), dtype=np.int8)
    for t in range(16):
        out[:, :, t] = _DEC[(words >> np.uint32(pos[t])) & np.uint32(3)]
    return out.reshape(R, B * 4)


# --------------------------------------------------------------------------
# greedy CTC decode (onebit_asr/metrics.py:51-60)
# --------------------------------------------------------------------------


def ctc_greedy_decode(logits: np.ndarray, blank_id: int = 3):
    """logits [T, V] -> token ids: argmax per frame (first maximal index), drop blanks, collapse repeats."""
    pred = np.argmax(np.asarray(logits), axis=-1).tolist()
    out, prev = [], None
    for t in pred:
        if t != blank_id and t != prev:
            out.append(int(t))
        prev = t
    return out


def ctc_greedy_decode_batch(logits: np.ndarray, lens, blank_id: int = 3):
    return [ctc_greedy_decode(logits[b, : int(lens[b])], blank_id) for b in range(logits.shape[0])]


# --------------------------------------------------------------------------
# layer init (quant.py:100-118) — needs torch for the RNG stream; imported lazily
# --------------------------------------------------------------------------


def init_layer_params(in_features: int, out_features: int, bias: bool = True):
    """Seeded construction identical to the reference: kaiming-uniform(a=sqrt 5) x2,
    alpha = mean|W| (0-dim), zero bias.  Consumes the global torch RNG exactly as the
    reference constructor does."""
    import torch

    w = torch.empty(out_features, in_features)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    w.mul_(2.0)
    a = w.abs().mean()
    b = torch.zeros(out_features) if bias else None
    return w, a, b


# ---------------------------------------------------------------------------------------------
# Dropout streams of the fused kernels (include/onebit.h: ob_swish_drop_quant, ob_relattn_softmax_fwd).
# The reference draws nn.Dropout masks from torch's generator (conformer.py:38, 128); any Bernoulli(1-p) stream is
# equivalent, so the kernels' stream is OUR spec: Philox4x32-10 (Salmon et al., SC'11), restated here from the
# published algorithm and pinned to the Random123 known-answer vectors in tests/test_oracle.py.
# ---------------------------------------------------------------------------------------------
_PHILOX_M0, _PHILOX_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PHILOX_W0, _PHILOX_W1 = 0x9E3779B9, 0xBB67AE85
_U32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Ten rounds of Philox-4x32 on arrays of counter words (uint32 values held in uint64); returns four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _U32 for c in (c0, c1, c2, c3))
    for r in range(10):
        ka = np.uint64((k0 + r * _PHILOX_W0) & 0xFFFFFFFF)
        kb = np.uint64((k1 + r * _PHILOX_W1) & 0xFFFFFFFF)
        p0, p1 = _PHILOX_M0 * c0, _PHILOX_M1 * c2
        c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ ka, p1 & _U32, (p0 >> np.uint64(32)) ^ c3 ^ kb, p0 & _U32)
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def _philox_words(group: np.ndarray, offset: int, seed: int) -> np.ndarray:
    """[..., 4] words of the block with counter (group, offset) and key seed, as the kernels lay them out."""
    group = np.asarray(group, dtype=np.uint64)
    zero = np.zeros_like(group)
    w = philox4x32_10(group & _U32, group >> np.uint64(32), zero + np.uint64(offset & 0xFFFFFFFF),
                      zero + np.uint64((offset >> 32) & 0xFFFFFFFF), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(w, axis=-1)


def _philox_lanes16(group: np.ndarray, offset: int, seed: int) -> np.ndarray:
    """[..., 8] 16-bit lanes of the block: lane k = half k % 2 (low first) of word k // 2."""
    w = _philox_words(group, offset, seed)
    return np.stack([w & np.uint32(0xFFFF), w >> np.uint32(16)], axis=-1).reshape(*w.shape[:-1], 8)


def dropout_keep_flat(n: int, seed: int, offset: int, threshold: int) -> np.ndarray:
    """Keep mask of the fused swish/dropout kernels over a flat tensor of n (multiple of 256) elements: with f = e // 4,
    element e is lane 4 * ((f >> 5) & 1) + e % 4 of the block with counter (f & ~32, offset); kept iff lane >= threshold
    (16-bit)."""
    assert n % 256 == 0 and 0 <= threshold < 65536
    e = np.arange(n, dtype=np.uint64)
    f = e >> np.uint64(2)
    lanes = _philox_lanes16(np.arange(n // 4, dtype=np.uint64), offset, seed)          # block of every f (half are unused)
    pick = (np.uint64(4) * ((f >> np.uint64(5)) & np.uint64(1)) + (e & np.uint64(3))).astype(np.int64)
    return lanes[(f & ~np.uint64(32)).astype(np.int64), pick] >= threshold


def dropout_keep_groups8(n: int, seed: int, offset: int, threshold: int) -> np.ndarray:
    """Keep mask of the module-tail kernel (ob_residual_dropout_fwd): element e is lane e % 8 of the block with counter
    (e // 8, offset)."""
    assert n % 8 == 0 and 0 <= threshold < 65536
    return (_philox_lanes16(np.arange(n // 8, dtype=np.uint64), offset, seed) >= threshold).reshape(n)


def dropout_keep_relattn(B: int, H: int, T: int, seed: int, offset: int, threshold: int) -> np.ndarray:
    """Keep mask [B,H,T,T] of the fused attention chain: column j = lane + 32 u of row r = (b*H + h)*T + i is 16-bit lane
    u % 8 of the block with counter ((r * 32 + lane) * 8 + u // 8, offset)."""
    assert 0 <= threshold < 65536
    rows = B * H * T
    nu = (T + 31) // 32
    r = np.arange(rows, dtype=np.uint64)[:, None, None]
    lane = np.arange(32, dtype=np.uint64)[None, :, None]
    g = np.arange((nu + 7) // 8, dtype=np.uint64)[None, None, :]
    lanes = _philox_lanes16((r * np.uint64(32) + lane) * np.uint64(8) + g, offset, seed)    # [rows, 32, G, 8]
    keep_lu = (lanes >= threshold).reshape(rows, 32, -1)                                   # [rows, lane, u]
    keep = keep_lu.transpose(0, 2, 1).reshape(rows, -1)[:, :T]                             # column = u * 32 + lane
    return keep.reshape(B, H, T, T)
