"""fp32 matrix products on the tensor cores (``ob_gemm_f32``): the non-routed matmuls of the model.

The reference computes its attention products (conformer.py:113-129), vocabulary projections and 1x1 convolutions
with fp32 ``torch.matmul`` / ``nn.Linear`` / ``nn.Conv1d``.  ``bmm_nt`` is the same contraction on tcgen05 with each
operand split into tf32 hi + lo parts (three products, fp32 accumulation), which keeps fp32-level accuracy; strided
views (``[B, T, H, d]`` projections, transposed operands, broadcast batches) are consumed in place through TMA.
"""
from __future__ import annotations

import os

import torch

from ._cabi import check, lib
from .quant import _stream


def _as4(t: torch.Tensor) -> torch.Tensor:
    while t.dim() < 4:
        t = t.unsqueeze(0)
    if t.dim() != 4:
        raise ValueError(f"expected at most 4 dimensions, got {tuple(t.shape)}")
    return t


def _operand(t: torch.Tensor, what: str):
    """(tensor, mn_major, ld, bs0, bs1) of a [nb0, nb1, rows, K] view; copies only if no axis is contiguous."""
    if t.dtype != torch.float32 or not t.is_cuda:
        raise RuntimeError(f"bmm_nt: {what} must be a CUDA float32 tensor (got {t.dtype} on {t.device}); there is no fallback")
    for attempt in range(2):
        s = t.stride()
        rows, K = t.shape[2], t.shape[3]
        if (s[3] == 1 or K == 1) and (rows == 1 or s[2] >= K):
            mn, ld = 0, (s[2] if rows > 1 else max(K, 4))
        elif (s[2] == 1 or rows == 1) and (K == 1 or s[3] >= rows):
            mn, ld = 1, (s[3] if K > 1 else max(rows, 4))
        else:
            mn, ld = -1, 0
        bs0 = s[0] if t.shape[0] > 1 else 0
        bs1 = s[1] if t.shape[1] > 1 else 0
        ok = mn >= 0 and ld % 4 == 0 and bs0 % 4 == 0 and bs1 % 4 == 0 and t.data_ptr() % 16 == 0 and bs0 >= 0 and bs1 >= 0
        if ok:
            return t, mn, ld, bs0, bs1
        if attempt == 0:                                   # unaligned / fully strided view: one packed copy, rows padded to 4
            rows_, K_ = t.shape[2], t.shape[3]
            buf = torch.empty(t.shape[0], t.shape[1], rows_, (K_ + 3) // 4 * 4, device=t.device, dtype=t.dtype)
            buf[..., :K_].copy_(t)
            t = buf[..., :K_]
    raise RuntimeError(f"bmm_nt: cannot describe {what} with strides {t.stride()}")


def bmm_nt(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor = None, bias: torch.Tensor = None, scale: float = 1.0,
           accumulate: bool = False, passes: int = 3) -> torch.Tensor:
    """``out[.., m, n] (+)= scale * sum_k a[.., m, k] * b[.., n, k] + bias[n]``.

    a: ``[.., M, K]`` and b: ``[.., N, K]`` views with up to two leading batch axes (size-1 axes of ``b`` or ``a``
    broadcast); either of the last two axes may be the contiguous one, so ``x @ y`` is ``bmm_nt(x, y.transpose(-1, -2))``
    without a copy.  out: ``[.., M, N]`` with a contiguous last axis whose pitch is a multiple of 4 (allocated if None)."""
    a4, b4 = _as4(a), _as4(b)
    nb0, nb1 = max(a4.shape[0], b4.shape[0]), max(a4.shape[1], b4.shape[1])
    M, K, N = a4.shape[2], a4.shape[3], b4.shape[2]
    if b4.shape[3] != K:
        raise ValueError(f"bmm_nt: contraction sizes differ ({K} vs {b4.shape[3]})")
    for t, name in ((a4, "a"), (b4, "b")):
        if t.shape[0] not in (1, nb0) or t.shape[1] not in (1, nb1):
            raise ValueError(f"bmm_nt: batch axes of {name} {tuple(t.shape[:2])} do not broadcast to {(nb0, nb1)}")
    a4, a_mn, lda, a_bs0, a_bs1 = _operand(a4, "a")
    b4, b_mn, ldb, b_bs0, b_bs1 = _operand(b4, "b")
    if out is None:
        if accumulate:
            raise ValueError("bmm_nt: accumulate needs an output tensor")
        pitch = (N + 3) // 4 * 4
        out4 = torch.empty(nb0, nb1, M, pitch, device=a.device, dtype=torch.float32)[..., :N]
        lead = len(torch.broadcast_shapes(a.shape[:-2], b.shape[:-2]))
        result = out4 if lead == 2 else (out4[0] if lead == 1 else out4[0, 0])
    else:
        out4 = _as4(out)
        result = out
        if out4.shape != (nb0, nb1, M, N) or out4.dtype != torch.float32 or not out4.is_cuda:
            raise ValueError(f"bmm_nt: out must be float32 CUDA of shape {(nb0, nb1, M, N)}, got {tuple(out4.shape)}")
    so = out4.stride()
    if (so[3] != 1 and N > 1) or (M > 1 and so[2] % 4 != 0) or out4.data_ptr() % 16 != 0:
        raise ValueError("bmm_nt: out needs a contiguous last axis, a row pitch that is a multiple of 4 and 16-byte alignment")
    ldd = so[2] if M > 1 else (N + 3) // 4 * 4
    d_bs0 = so[0] if nb0 > 1 else 0
    d_bs1 = so[1] if nb1 > 1 else 0
    if bias is not None and (bias.dtype != torch.float32 or not bias.is_cuda or bias.numel() != N or not bias.is_contiguous()):
        raise ValueError("bmm_nt: bias must be a contiguous float32 CUDA vector of N elements")
    ws_bytes = lib.ob_gemm_f32_workspace_bytes(M, N, K, nb0, nb1) if K >= 2048 and nb0 * nb1 == 1 else 0
    ws = torch.empty(ws_bytes, device=a.device, dtype=torch.uint8) if ws_bytes else None      # split-K partial products
    check(lib.ob_gemm_f32(a4.data_ptr(), a_mn, lda, a_bs0, a_bs1, b4.data_ptr(), b_mn, ldb, b_bs0, b_bs1, out4.data_ptr(), ldd,
                          d_bs0, d_bs1, None if bias is None else bias.data_ptr(), float(scale), int(accumulate), M, N, K, nb0,
                          nb1, passes, None if ws is None else ws.data_ptr(), ws_bytes, _stream()))
    return result


class _LinearFn(torch.autograd.Function):
    """``F.linear`` of the non-routed fp32 layers (vocabulary projections, front-end output) on ``bmm_nt``."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x2 = x.reshape(-1, x.shape[-1])
        y = bmm_nt(x2, weight, bias=bias)
        if not y.is_contiguous():                          # out_features not a multiple of 4: drop the padded pitch
            y = y.contiguous()
        ctx.save_for_backward(x2, weight)
        ctx.has_bias, ctx.x_shape = bias is not None, x.shape
        return y.reshape(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, weight = ctx.saved_tensors
        g2 = gy.reshape(-1, gy.shape[-1])
        if not g2.is_contiguous():
            g2 = g2.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = bmm_nt(g2, weight.t()).contiguous().reshape(ctx.x_shape)                    # dY . W      (W read as an MN-major operand)
        if ctx.needs_input_grad[1]:
            gw = bmm_nt(g2.t(), x2.t())                                         # dY^T . X    (both MN-major, split over the rows)
            if not gw.is_contiguous():
                gw = gw.contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = g2.sum(0)
        return gx, gw, gb


# A/B switches for measurements (bench.py --torch-nonrouted): "attn", "conv", "linear" in OB_TORCH_NONROUTED route that part of
# the non-routed stack back to torch's own fp32 kernels
DISABLED = set(filter(None, os.environ.get("OB_TORCH_NONROUTED", "").split(",")))


def linear_usable(x: torch.Tensor, weight: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and x.numel() > 0
            and "linear" not in DISABLED)


def linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor = None) -> torch.Tensor:
    """``x @ weight.T + bias`` in fp32 on the tensor cores (3 x tf32 split); ``torch.nn.functional.linear`` elsewhere."""
    if not linear_usable(x, weight):
        return torch.nn.functional.linear(x, weight, bias)
    return _LinearFn.apply(x, weight, bias)
