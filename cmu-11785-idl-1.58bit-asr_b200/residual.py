"""Tail of every encoder module of the reference (conformer.py:45, 133-138, 163-167) in one kernel each way:
``x + scale * dropout(y) * frame_mask`` instead of torch's dropout, mask multiply, scale and add (and their four backward
kernels).  The dropout bits come from the library's Philox stream and are regenerated in the backward.
"""
from __future__ import annotations

import torch

from ._cabi import check, lib
from .quant import _NO_RNG, _stream, draw_dropout_stream


class _ResidualDropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, rowmask, scale, inv_keep, rng):
        C = x.shape[-1]
        x2, y2 = x.contiguous(), y.contiguous()
        M = x2.numel() // C
        out = torch.empty_like(x2)
        check(lib.ob_residual_dropout_fwd(x2.data_ptr(), y2.data_ptr(), None if rowmask is None else rowmask.data_ptr(), scale,
                                          inv_keep, *rng, M, C, out.data_ptr(), _stream()))
        ctx.rowmask, ctx.scale, ctx.inv_keep, ctx.rng, ctx.dims = rowmask, scale, inv_keep, rng, (M, C)
        return out

    @staticmethod
    def backward(ctx, g):
        M, C = ctx.dims
        g = g.contiguous()
        gy = None
        if ctx.needs_input_grad[1]:
            gy = torch.empty_like(g)
            check(lib.ob_residual_dropout_bwd(g.data_ptr(), None if ctx.rowmask is None else ctx.rowmask.data_ptr(), ctx.scale,
                                              ctx.inv_keep, *ctx.rng, M, C, gy.data_ptr(), _stream()))
        return (g if ctx.needs_input_grad[0] else None), gy, None, None, None, None


def usable(x: torch.Tensor) -> bool:
    from .matmul import DISABLED
    return x.is_cuda and x.dtype == torch.float32 and x.shape[-1] % 8 == 0 and x.numel() > 0 and "tail" not in DISABLED


class _RowMaskCache:
    """The frame-validity mask is the same tensor for every module of a pass: convert it to a float row vector once."""
    src = None
    key = None
    rows = None

    @classmethod
    def get(cls, frame_mask: torch.Tensor) -> torch.Tensor:
        key = (frame_mask.data_ptr(), frame_mask._version, tuple(frame_mask.shape), tuple(frame_mask.stride()), frame_mask.dtype)
        if cls.key != key:
            cls.src, cls.key = frame_mask, key            # keeps the storage alive, so the key cannot be recycled
            cls.rows = frame_mask.reshape(-1).to(torch.float32).contiguous()
        return cls.rows


def residual_dropout(x, y, frame_mask=None, scale: float = 1.0, p: float = 0.0, training: bool = False):
    """``x + scale * dropout(y, p) * frame_mask``; frame_mask: ``[..., 1]`` (or ``[...]``) of 0/1 per row, or None."""
    rowmask = None if frame_mask is None else _RowMaskCache.get(frame_mask)
    inv_keep, rng = draw_dropout_stream(x.device, p) if (training and p > 0.0) else (1.0, _NO_RNG)
    return _ResidualDropoutFn.apply(x, y, rowmask, float(scale), inv_keep, rng)
