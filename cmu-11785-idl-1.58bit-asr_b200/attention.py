"""Relative-position self-attention core of the reference (conformer.py:96-129) on B200 kernels.

``rel_attention_probs(ac, bd_raw, mask, scale, p, training)`` returns what the reference computes as
``dropout(nan_to_num(softmax(masked_fill((ac + rel_shift(bd_raw)) / sqrt(d), mask == 0, -inf))))`` - one kernel forward,
one backward, fp32, same formulas - instead of nine / twelve torch kernels over ``[B, H, T, T]`` tensors.

``rel_attention(q, k, v, pos, u, w, mask, ...)`` is the whole core between the projections: the three matrix products
of the forward and the six of the backward run on the tensor cores with fp32-level accuracy (``matmul.bmm_nt``, 3 x tf32
split), read the ``[B, T, H*d]`` projection outputs in place (no head transposes) and write the mixed values back in that
layout.  The score matrices are kept with a row pitch that is a multiple of 4 floats (TMA alignment).
"""
from __future__ import annotations

import torch

from ._cabi import check, lib
from .matmul import bmm_nt
from .quant import _NO_RNG, _stream, draw_dropout_stream

MAX_T = 2048


def _pitch(t: torch.Tensor) -> int:
    """Row pitch of a [B,H,T,T] score tensor laid out as dense rows with a common pitch, or 0 if it is not."""
    B, H, T, _ = t.shape
    ld = t.stride(2) if T > 1 else T
    ok = t.stride(3) == 1 and ld >= T and t.stride(1) == T * ld and (B == 1 or t.stride(0) == H * T * ld)
    return ld if ok else 0


def _scores_like(B, H, T, device) -> torch.Tensor:
    ld = (T + 3) // 4 * 4
    return torch.empty(B, H, T, ld, device=device, dtype=torch.float32)[..., :T]


def _softmax_fwd(ac, bd_raw, mask, keep, inv_keep, scale, rng):
    B, H, T, _ = ac.shape
    ld = _pitch(ac)
    if ld == 0 or _pitch(bd_raw) != ld:
        ac, bd_raw = ac.contiguous(), bd_raw.contiguous()
        ld = T
    y = torch.empty_strided(ac.shape, ac.stride(), device=ac.device, dtype=ac.dtype)
    drop = keep is not None or rng[2] != 0
    attn_d = torch.empty_strided(ac.shape, ac.stride(), device=ac.device, dtype=ac.dtype) if drop else None
    check(lib.ob_relattn_softmax_fwd(ac.data_ptr(), bd_raw.data_ptr(), mask.data_ptr(),
                                     None if keep is None else keep.data_ptr(), inv_keep, *rng, scale, B, H, T, ld,
                                     y.data_ptr(), None if attn_d is None else attn_d.data_ptr(), _stream()))
    return y, attn_d


def _softmax_bwd(g, y, keep, inv_keep, scale, rng):
    B, H, T, _ = y.shape
    ld = _pitch(y)
    if _pitch(g) != ld:
        gg = torch.empty_strided(y.shape, y.stride(), device=y.device, dtype=y.dtype)
        gg.copy_(g)
        g = gg
    d_ac = torch.empty_strided(y.shape, y.stride(), device=y.device, dtype=y.dtype)
    d_bd = torch.empty_strided(y.shape, y.stride(), device=y.device, dtype=y.dtype)
    check(lib.ob_relattn_softmax_bwd(g.data_ptr(), y.data_ptr(), None if keep is None else keep.data_ptr(), inv_keep, *rng,
                                     scale, B, H, T, ld, d_ac.data_ptr(), d_bd.data_ptr(), _stream()))
    return d_ac, d_bd


class _RelAttnSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ac, bd_raw, mask, keep, inv_keep, scale, rng):
        y, attn_d = _softmax_fwd(ac, bd_raw, mask, keep, inv_keep, scale, rng)
        if keep is None:
            ctx.save_for_backward(y)
        else:
            ctx.save_for_backward(y, keep)
        ctx.inv_keep, ctx.scale, ctx.rng = inv_keep, scale, rng
        return y if attn_d is None else attn_d

    @staticmethod
    def backward(ctx, g):
        saved = ctx.saved_tensors
        keep = saved[1] if len(saved) > 1 else None
        d_ac, d_bd = _softmax_bwd(g, saved[0], keep, ctx.inv_keep, ctx.scale, ctx.rng)
        return d_ac, d_bd, None, None, None, None, None


def usable(ac: torch.Tensor, mask) -> bool:
    return (ac.is_cuda and ac.dtype == torch.float32 and mask is not None and ac.dim() == 4
            and ac.shape[-1] == ac.shape[-2] and ac.shape[-1] <= MAX_T)


def _dropout_args(device, p, training, keep):
    if keep is not None:
        return keep.contiguous(), 1.0 / (1.0 - p), _NO_RNG
    if training and p > 0.0:
        inv_keep, rng = draw_dropout_stream(device, p)                              # mask generated inside the kernels
        return None, inv_keep, rng
    return None, 1.0, _NO_RNG


def rel_attention_probs(ac, bd_raw, mask, scale: float, p: float = 0.0, training: bool = False, keep=None):
    """ac, bd_raw: [B, H, T, T] fp32 (bd_raw BEFORE the relative shift); mask: [B, T, T] bool (False = masked).
    ``keep`` (bool [B,H,T,T]) overrides the sampled dropout mask (tests)."""
    keep, inv_keep, rng = _dropout_args(ac.device, p, training, keep)
    return _RelAttnSoftmaxFn.apply(ac, bd_raw, mask.contiguous(), keep, inv_keep, scale, rng)


def _heads(t: torch.Tensor, H: int) -> torch.Tensor:
    """[B, T, H*d] -> [B, H, T, d] view (no copy)."""
    B, T, W = t.shape
    return t.view(B, T, H, W // H).permute(0, 2, 1, 3)


class _RelAttentionFn(torch.autograd.Function):
    """out[b, t, h*d:(h+1)*d] = dropout(softmax(mask((q+u) k^T + shift((q+w) pos^T)) * scale)) v   per head.

    q, k, v: [B, T, H*d] (projection outputs, read in place), pos: [1, T, H*d] (or [B, T, H*d]), u, w: [H, d].
    ``pos_b`` / ``split``: a second positional table [1, T, H*d] used by the utterances ``[split, B)`` - the stacked co-training
    passes project the table once per bitwidth; the positional product runs once per group with the table broadcast, instead
    of materialising a per-utterance copy of it."""

    @staticmethod
    def forward(ctx, q, k, v, pos, u, w, mask, keep, inv_keep, scale, rng, n_heads, pos_b=None, split=0):
        B, T, W = q.shape
        H = n_heads
        q, k, v, pos = q.contiguous(), k.contiguous(), v.contiguous(), pos.contiguous()
        if pos_b is not None:
            pos_b = pos_b.contiguous()
        if W % 64 == 0:
            qu, qw = torch.empty_like(q), torch.empty_like(q)
            check(lib.ob_add_bias2(q.data_ptr(), u.contiguous().data_ptr(), w.contiguous().data_ptr(), B * T, W, qu.data_ptr(),
                                   qw.data_ptr(), _stream()))
        else:
            qu, qw = q + u.reshape(1, 1, W), q + w.reshape(1, 1, W)
        ac = bmm_nt(_heads(qu, H), _heads(k, H), out=_scores_like(B, H, T, q.device))
        bd = _scores_like(B, H, T, q.device)
        if pos_b is None:
            bmm_nt(_heads(qw, H), _heads(pos, H), out=bd)
        else:
            bmm_nt(_heads(qw[:split], H), _heads(pos, H), out=bd[:split])
            bmm_nt(_heads(qw[split:], H), _heads(pos_b, H), out=bd[split:])
        y, attn_d = _softmax_fwd(ac, bd, mask, keep, inv_keep, scale, rng)
        probs = y if attn_d is None else attn_d
        out = torch.empty_like(q)
        bmm_nt(probs, _heads(v, H).transpose(-1, -2), out=_heads(out, H))
        ctx.save_for_backward(qu, qw, k, v, pos, y, probs, pos if pos_b is None else pos_b, *(() if keep is None else (keep,)))
        ctx.inv_keep, ctx.scale, ctx.rng, ctx.H = inv_keep, scale, rng, H
        ctx.two_tables, ctx.split = pos_b is not None, split
        return out

    @staticmethod
    def backward(ctx, g):
        qu, qw, k, v, pos, y, probs, pos_second = ctx.saved_tensors[:8]
        keep = ctx.saved_tensors[8] if len(ctx.saved_tensors) > 8 else None
        H = ctx.H
        B, T, W = qu.shape
        g = g.contiguous()
        gh = _heads(g, H)
        g_probs = bmm_nt(gh, _heads(v, H), out=_scores_like(B, H, T, g.device))               # dO . v^T
        g_v = torch.empty_like(v)
        bmm_nt(probs.transpose(-1, -2), gh.transpose(-1, -2), out=_heads(g_v, H))              # P^T . dO
        d_ac, d_bd = _softmax_bwd(g_probs, y, keep, ctx.inv_keep, ctx.scale, ctx.rng)
        g_qu, g_qw, g_k = torch.empty_like(qu), torch.empty_like(qu), torch.empty_like(k)
        bmm_nt(d_ac, _heads(k, H).transpose(-1, -2), out=_heads(g_qu, H))                      # dS_ac . k
        bmm_nt(d_ac.transpose(-1, -2), _heads(qu, H).transpose(-1, -2), out=_heads(g_k, H))    # dS_ac^T . (q+u)
        pos_b, split = (pos_second if ctx.two_tables else None), ctx.split
        if pos_b is None:
            bmm_nt(d_bd, _heads(pos, H).transpose(-1, -2), out=_heads(g_qw, H))                # dS_bd . pos
        else:
            bmm_nt(d_bd[:split], _heads(pos, H).transpose(-1, -2), out=_heads(g_qw[:split], H))
            bmm_nt(d_bd[split:], _heads(pos_b, H).transpose(-1, -2), out=_heads(g_qw[split:], H))
        g_pos_b = torch.empty_like(qw)                                                         # per utterance, then summed
        bmm_nt(d_bd.transpose(-1, -2), _heads(qw, H).transpose(-1, -2), out=_heads(g_pos_b, H))
        g_pos2 = None
        if pos_b is not None:                                                                  # one table per group
            g_pos, g_pos2 = g_pos_b[:split].sum(dim=0, keepdim=True), g_pos_b[split:].sum(dim=0, keepdim=True)
        else:
            g_pos = g_pos_b.sum(dim=0, keepdim=True) if pos.shape[0] == 1 else g_pos_b    # shared table: sum over the batch
        if W % 64 == 0:
            g_q, sums = torch.empty_like(qu), torch.empty(2, W, device=g.device, dtype=torch.float32)
            ws = torch.empty(lib.ob_add_colsum2_workspace_bytes(B * T, W), device=g.device, dtype=torch.uint8)
            check(lib.ob_add_colsum2(g_qu.data_ptr(), g_qw.data_ptr(), B * T, W, g_q.data_ptr(), sums.data_ptr(), ws.data_ptr(),
                                     _stream()))
            g_u, g_w = sums[0].view(H, W // H), sums[1].view(H, W // H)
        else:
            g_q = g_qu + g_qw
            g_u, g_w = g_qu.sum(dim=(0, 1)).view(H, W // H), g_qw.sum(dim=(0, 1)).view(H, W // H)
        return g_q, g_k, g_v, g_pos, g_u, g_w, None, None, None, None, None, None, g_pos2, None


def rel_attention_usable(q: torch.Tensor, mask, n_heads: int) -> bool:
    from .matmul import DISABLED
    return (q.is_cuda and q.dtype == torch.float32 and mask is not None and q.dim() == 3 and q.shape[1] <= MAX_T
            and q.shape[2] % (4 * n_heads) == 0 and "attn" not in DISABLED)


def rel_attention(q, k, v, pos, u, w, mask, n_heads: int, p: float = 0.0, training: bool = False, keep=None, pos_b=None,
                  split: int = 0):
    """The attention core of ``MHSA.forward`` (conformer.py:113-129) without the projections.

    q, k, v: ``[B, T, H*d]``; pos: ``[1, T, H*d]`` (projected positional encoding) or ``[B, T, H*d]``; u, w: ``[H, d]`` (``pos_bias_u``,
    ``pos_bias_v``); mask: ``[B, T, T]`` bool.  ``pos_b`` ``[1, T, H*d]`` + ``split``: utterances ``[split, B)`` use this second table
    (stacked passes at two bitwidths).  Returns ``[B, T, H*d]``."""
    d = q.shape[2] // n_heads
    keep, inv_keep, rng = _dropout_args(q.device, p, training, keep)
    if pos_b is not None and not (0 < split < q.shape[0] and pos.shape[0] == 1 and pos_b.shape[0] == 1):
        raise ValueError("rel_attention: pos_b needs 0 < split < batch and two [1, T, W] tables")
    return _RelAttentionFn.apply(q, k, v, pos, u, w, mask.contiguous(), keep, inv_keep, 1.0 / (d ** 0.5), rng, n_heads, pos_b,
                                 split)
