"""LayerNorm in front of the routed projections on the B200 library.

Same math as ``torch.nn.LayerNorm`` over the last axis (conformer.py:19-24), fp32; the backward is one streaming
kernel for dx plus a fixed-order two-stage reduction for the parameter gradients (PyTorch's gamma/beta backward
kernel takes ~130 us on a [25536, 256] input; this one is bandwidth-bound).
"""
from __future__ import annotations

import functools

import torch

from . import routes
from ._cabi import check, lib
from .quant import _stream

SUPPORTED_WIDTHS = (128, 256, 512, 1024)


@functools.lru_cache(maxsize=None)
def _bwd_ws_bytes(C: int) -> int:
    return lib.ob_layernorm_bwd_workspace_bytes(C)


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        C = x.shape[-1]
        x2 = x.reshape(-1, C)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        M = x2.shape[0]
        y = torch.empty_like(x2)
        stats = torch.empty((2, M), device=x.device, dtype=torch.float32)
        sp = stats.data_ptr()                                  # mean row, rstd row (pointer arithmetic: no view objects)
        check(lib.ob_layernorm_fwd(x2.data_ptr(), weight.data_ptr(), bias.data_ptr(), eps, M, C, y.data_ptr(), sp, sp + 4 * M,
                                   _stream()))
        ctx.save_for_backward(x2, stats, weight)
        ctx.x_shape = x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, gy):
        x2, stats, weight = ctx.saved_tensors
        M, C = x2.shape
        g2 = gy.reshape(M, C)
        if not g2.is_contiguous():
            g2 = g2.contiguous()
        dx = torch.empty_like(x2)
        dparams = torch.empty((2, C), device=x2.device, dtype=torch.float32)
        ws = torch.empty(_bwd_ws_bytes(C), device=x2.device, dtype=torch.uint8)
        sp, dp = stats.data_ptr(), dparams.data_ptr()
        check(lib.ob_layernorm_bwd(g2.data_ptr(), x2.data_ptr(), sp, sp + 4 * M, weight.data_ptr(), M, C, dx.data_ptr(), dp,
                                   dp + 4 * C, ws.data_ptr(), _stream()))
        return dx.view(ctx.x_shape), dparams[0], dparams[1], None


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """LayerNorm over the last axis; uses the library for CUDA fp32 inputs of a supported width, torch otherwise."""
    on_library = (x.is_cuda and x.dtype == torch.float32 and x.shape[-1] in SUPPORTED_WIDTHS and x.numel() > 0
                  and weight is not None and bias is not None and weight.dtype == torch.float32)
    if routes.taken("layer_norm", on_library, x):
        return _LayerNormFn.apply(x, weight, bias, eps)
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), weight, bias, eps)
