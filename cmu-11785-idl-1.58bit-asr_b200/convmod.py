"""Middle of the reference's convolution module (conformer.py:141-167) on B200 kernels, in the ``[B, T, C]`` layout.

``glu_dwconv_bn_swish(a, dw_weight, dw_bias, bn_weight, bn_bias, eps)`` = ``swish(BatchNorm(depthwise_conv1d(GLU(a))))``
with batch statistics (``nn.BatchNorm1d(track_running_stats=False)``), two kernels forward and five backward instead
of cuDNN's depthwise convolutions plus GLU / BatchNorm / swish / transposes; the 1x1 convolutions on either side are
``matmul.linear`` over the channel axis, so the module never leaves the channel-last layout.
"""
from __future__ import annotations

import functools

import torch

from ._cabi import check, lib
from .quant import _stream

MAX_TAPS = 31


def usable(x: torch.Tensor, channels: int, taps: int) -> bool:
    from .matmul import DISABLED
    return (x.is_cuda and x.dtype == torch.float32 and channels % 64 == 0 and taps <= MAX_TAPS and taps % 2 == 1
            and "conv" not in DISABLED)


@functools.lru_cache(maxsize=None)
def _ws_bytes(B: int, T: int, C: int) -> int:
    return lib.ob_convmod_workspace_bytes(B, T, C)


def _workspace(B, T, C, device):
    return torch.empty(_ws_bytes(B, T, C), device=device, dtype=torch.uint8)


class _GluDwBnSwishFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, dw_weight, dw_bias, gamma, beta, eps, groups):
        B, T, C2 = a.shape
        C = C2 // 2
        ks = dw_weight.shape[-1]
        a = a.contiguous()
        w = dw_weight.reshape(C, ks).contiguous()
        gamma, beta = gamma.contiguous(), beta.contiguous()
        d = torch.empty(B, T, C, device=a.device, dtype=a.dtype)
        stats = torch.empty(2, groups, C, device=a.device, dtype=a.dtype)      # mean rows, rstd rows (per group)
        ws = _workspace(B, T, C, a.device)
        st = _stream()
        sp = stats.data_ptr()
        check(lib.ob_glu_dwconv_bn_fwd(a.data_ptr(), w.data_ptr(), None if dw_bias is None else dw_bias.data_ptr(), B, T, C, ks,
                                       float(eps), groups, d.data_ptr(), sp, sp + 4 * groups * C, ws.data_ptr(), st))
        s = torch.empty_like(d)
        check(lib.ob_bn_swish_fwd(d.data_ptr(), sp, sp + 4 * groups * C, gamma.data_ptr(), beta.data_ptr(), B * T, C, groups,
                                  s.data_ptr(), st))
        ctx.save_for_backward(a, w, d, stats, gamma, beta)
        ctx.has_bias, ctx.w_shape, ctx.groups = dw_bias is not None, dw_weight.shape, groups
        return s

    @staticmethod
    def backward(ctx, gs):
        a, w, d, stats, gamma, beta = ctx.saved_tensors
        B, T, C = d.shape
        ks = w.shape[1]
        groups = ctx.groups
        gs = gs.contiguous()
        ws = _workspace(B, T, C, a.device)
        st = _stream()
        gd = torch.empty_like(d)
        ggb = torch.empty(groups, 2, C, device=a.device, dtype=a.dtype)        # per group: (g_beta, g_gamma)
        sp = stats.data_ptr()
        check(lib.ob_bn_swish_bwd(gs.data_ptr(), d.data_ptr(), sp, sp + 4 * groups * C, gamma.data_ptr(), beta.data_ptr(), B * T, C,
                                  groups, gd.data_ptr(), ggb.data_ptr(), ws.data_ptr(), st))
        ga = torch.empty_like(a)
        gw = torch.empty_like(w)
        gb = torch.empty(C, device=a.device, dtype=a.dtype) if ctx.has_bias else None
        check(lib.ob_glu_dwconv_bwd(gd.data_ptr(), a.data_ptr(), w.data_ptr(), B, T, C, ks, ga.data_ptr(), gw.data_ptr(),
                                    None if gb is None else gb.data_ptr(), ws.data_ptr(), st))
        ggb = ggb[0] if groups == 1 else ggb.sum(0)
        return ga, gw.view(ctx.w_shape), gb, ggb[1], ggb[0], None, None


def glu_dwconv_bn_swish(a, dw_weight, dw_bias, bn_weight, bn_bias, eps: float = 1e-5, groups: int = 1):
    """a: ``[B, T, 2C]`` (output of the first 1x1 convolution); dw_weight: ``[C, 1, ks]`` (``nn.Conv1d(groups=C)``);
    bn_weight / bn_bias: ``[C]``.  ``groups`` > 1: the batch stacks that many independent passes (``B / groups`` utterances
    each), each normalised with its own batch statistics.  Returns ``[B, T, C]``."""
    return _GluDwBnSwishFn.apply(a, dw_weight, dw_bias, bn_weight, bn_bias, eps, groups)
