"""First layer of the reference's subsampling front-end (conformer.py:177-181) on a B200 kernel pair.

``conv1_relu(feats, weight, bias)`` = ``relu(conv2d(feats[:, None], weight, bias, stride=2))`` for the one-input-channel
3 x 3 convolution: a write-bound stencil (2 GB of activations for the training batch) done in one pass each way, emitted
channels-last so that cuDNN's tensor-core kernels of the second convolution need no layout conversion.
"""
from __future__ import annotations

import torch

from ._cabi import check, lib
from .quant import _stream


class _Conv1ReluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, weight, bias):
        B, T, F = feats.shape
        C = weight.shape[0]
        x = feats.contiguous()
        w = weight.reshape(C, 9).contiguous()
        y = torch.empty(B, (T - 3) // 2 + 1, (F - 3) // 2 + 1, C, device=feats.device, dtype=feats.dtype)
        check(lib.ob_conv1_relu_fwd(x.data_ptr(), w.data_ptr(), None if bias is None else bias.data_ptr(), B, T, F, C,
                                    y.data_ptr(), _stream()))
        ctx.save_for_backward(x, w, bias)
        ctx.w_shape = weight.shape
        return y.permute(0, 3, 1, 2)                      # [B, C, T1, F1] view with channels-last strides

    @staticmethod
    def backward(ctx, g):
        x, w, bias = ctx.saved_tensors
        B, T, F = x.shape
        C = w.shape[0]
        g = g.permute(0, 2, 3, 1).contiguous()            # no copy when the gradient arrives channels-last
        gw = torch.empty_like(w)
        gb = None if bias is None else torch.empty_like(bias)
        ws = torch.empty(lib.ob_conv1_relu_workspace_bytes(), device=x.device, dtype=torch.uint8)
        check(lib.ob_conv1_relu_bwd(g.data_ptr(), x.data_ptr(), w.data_ptr(), None if bias is None else bias.data_ptr(), B, T, F, C,
                                    gw.data_ptr(), None if gb is None else gb.data_ptr(), ws.data_ptr(), _stream()))
        return None, gw.view(ctx.w_shape), gb


def usable(feats: torch.Tensor, conv: torch.nn.Conv2d) -> bool:
    from .matmul import DISABLED
    return (feats.is_cuda and feats.dtype == torch.float32 and feats.dim() == 3 and not feats.requires_grad
            and conv.in_channels == 1 and conv.out_channels == 256 and conv.kernel_size == (3, 3) and conv.stride == (2, 2)
            and conv.padding == (0, 0) and conv.dilation == (1, 1) and feats.shape[1] >= 3 and feats.shape[2] >= 3
            and "frontend" not in DISABLED)


def conv1_relu(feats: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """feats ``[B, T, F]`` -> ``[B, C, T1, F1]`` (channels-last memory format)."""
    return _Conv1ReluFn.apply(feats, weight, bias)
