"""Fused element-wise chain of the reference's relative-position attention (conformer.py:118-128).

``rel_attention_probs(ac, bd_raw, mask, scale, p, training)`` returns what the reference computes as
``dropout(nan_to_num(softmax(masked_fill((ac + rel_shift(bd_raw)) / sqrt(d), mask == 0, -inf))))`` - one kernel forward,
one backward, fp32, same formulas - instead of nine / twelve torch kernels over ``[B, H, T, T]`` tensors.  The matmuls
around it stay the reference's fp32 ``torch.matmul``.
"""
from __future__ import annotations

import torch

from ._cabi import check, lib
from .quant import _NO_RNG, _stream, draw_dropout_stream

MAX_T = 2048


class _RelAttnSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ac, bd_raw, mask, keep, inv_keep, scale, rng):
        B, H, T, _ = ac.shape
        ac, bd_raw = ac.contiguous(), bd_raw.contiguous()
        y = torch.empty_like(ac)
        attn_d = torch.empty_like(ac) if (keep is not None or rng[2] != 0) else None
        check(lib.ob_relattn_softmax_fwd(ac.data_ptr(), bd_raw.data_ptr(), mask.data_ptr(),
                                         None if keep is None else keep.data_ptr(), inv_keep, *rng, scale, B, H, T,
                                         y.data_ptr(), None if attn_d is None else attn_d.data_ptr(), _stream()))
        if keep is None:
            ctx.save_for_backward(y)
        else:
            ctx.save_for_backward(y, keep)
        ctx.inv_keep, ctx.scale, ctx.rng = inv_keep, scale, rng
        return y if attn_d is None else attn_d

    @staticmethod
    def backward(ctx, g):
        saved = ctx.saved_tensors
        y = saved[0]
        keep = saved[1] if len(saved) > 1 else None
        B, H, T, _ = y.shape
        g = g.contiguous()
        d_ac = torch.empty_like(y)
        d_bd = torch.empty_like(y)
        check(lib.ob_relattn_softmax_bwd(g.data_ptr(), y.data_ptr(), None if keep is None else keep.data_ptr(),
                                         ctx.inv_keep, *ctx.rng, ctx.scale, B, H, T, d_ac.data_ptr(), d_bd.data_ptr(),
                                         _stream()))
        return d_ac, d_bd, None, None, None, None, None


def usable(ac: torch.Tensor, mask) -> bool:
    return (ac.is_cuda and ac.dtype == torch.float32 and mask is not None and ac.dim() == 4
            and ac.shape[-1] == ac.shape[-2] and ac.shape[-1] <= MAX_T)


def rel_attention_probs(ac, bd_raw, mask, scale: float, p: float = 0.0, training: bool = False, keep=None):
    """ac, bd_raw: [B, H, T, T] fp32 (bd_raw BEFORE the relative shift); mask: [B, T, T] bool (False = masked).
    ``keep`` (bool [B,H,T,T]) overrides the sampled dropout mask (tests)."""
    mask = mask.contiguous()
    inv_keep, rng = 1.0, _NO_RNG
    if keep is not None:
        keep, inv_keep = keep.contiguous(), 1.0 / (1.0 - p)
    elif training and p > 0.0:
        inv_keep, rng = draw_dropout_stream(ac.device, p)                      # mask generated inside the kernels
    return _RelAttnSoftmaxFn.apply(ac, bd_raw, mask, keep, inv_keep, scale, rng)
