"""fp32 matrix products on the tensor cores (``ob_gemm_f32``): the non-routed matmuls of the model.

The reference computes its attention products (conformer.py:113-129), vocabulary projections and 1x1 convolutions
with fp32 ``torch.matmul`` / ``nn.Linear`` / ``nn.Conv1d``.  ``bmm_nt`` is the same contraction on tcgen05 with each
operand split into tf32 hi + lo parts (three products, fp32 accumulation), which keeps fp32-level accuracy; strided
views (``[B, T, H, d]`` projections, transposed operands, broadcast batches) are consumed in place through TMA.
"""
from __future__ import annotations

import os

import torch

from . import routes
from ._cabi import check, lib
from .quant import _stream


def _describe(shape, stride):
    """Pure shape/stride arithmetic (no tensors touched): a ``[.., rows, K]`` view with up to two leading batch axes ->
    ``(nb0, nb1, rows, K, mn_major, ld, bs0, bs1)`` as ``ob_gemm_f32`` wants them, or None if the view has no contiguous
    matrix axis / an unaligned pitch (the caller then packs a copy).  K-major: the contraction axis is contiguous."""
    nd = len(shape)
    if nd < 2 or nd > 4:
        raise ValueError(f"bmm_nt: expected 2 to 4 dimensions, got shape {tuple(shape)}")
    rows, K = shape[-2], shape[-1]
    sr, sk = stride[-2], stride[-1]
    nb1, bs1 = (shape[-3], stride[-3]) if nd >= 3 else (1, 0)
    nb0, bs0 = (shape[-4], stride[-4]) if nd == 4 else (1, 0)
    if nb1 == 1:
        bs1 = 0
    if nb0 == 1:
        bs0 = 0
    if (sk == 1 or K == 1) and (rows == 1 or sr >= K):
        mn, ld = 0, (sr if rows > 1 else max((K + 3) // 4 * 4, 4))
    elif (sr == 1 or rows == 1) and (K == 1 or sk >= rows):
        mn, ld = 1, (sk if K > 1 else max((rows + 3) // 4 * 4, 4))
    else:
        return None
    if ld % 4 or bs0 % 4 or bs1 % 4 or bs0 < 0 or bs1 < 0:
        return None
    return nb0, nb1, rows, K, mn, ld, bs0, bs1


def _input(t: torch.Tensor, what: str):
    """(tensor kept alive, data_ptr, descriptor) of an input operand; packs one copy (rows padded to 4 floats) only when the
    view cannot be described or is not 16-byte aligned."""
    if t.dtype != torch.float32 or not t.is_cuda:
        raise RuntimeError(f"bmm_nt: {what} must be a CUDA float32 tensor (got {t.dtype} on {t.device}); there is no fallback")
    d = _describe(t.shape, t.stride())
    ptr = t.data_ptr()
    if d is None or ptr % 16:
        K = t.shape[-1]
        buf = torch.empty(*t.shape[:-1], (K + 3) // 4 * 4, device=t.device, dtype=t.dtype)
        packed = buf[..., :K]
        packed.copy_(t)
        t, ptr, d = packed, packed.data_ptr(), _describe(packed.shape, packed.stride())
        if d is None:
            raise RuntimeError(f"bmm_nt: cannot describe {what} of shape {tuple(t.shape)}")
    return t, ptr, d


def bmm_nt(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor = None, bias: torch.Tensor = None, scale: float = 1.0,
           accumulate: bool = False, passes: int = 3) -> torch.Tensor:
    """``out[.., m, n] (+)= scale * sum_k a[.., m, k] * b[.., n, k] + bias[n]``.

    a: ``[.., M, K]`` and b: ``[.., N, K]`` views with up to two leading batch axes (size-1 axes of ``b`` or ``a``
    broadcast); either of the last two axes may be the contiguous one, so ``x @ y`` is ``bmm_nt(x, y.transpose(-1, -2))``
    without a copy.  out: ``[.., M, N]`` with a contiguous last axis whose pitch is a multiple of 4 (allocated if None)."""
    a, a_ptr, (a_nb0, a_nb1, M, K, a_mn, lda, a_bs0, a_bs1) = _input(a, "a")
    b, b_ptr, (b_nb0, b_nb1, N, Kb, b_mn, ldb, b_bs0, b_bs1) = _input(b, "b")
    if Kb != K:
        raise ValueError(f"bmm_nt: contraction sizes differ ({K} vs {Kb})")
    nb0, nb1 = max(a_nb0, b_nb0), max(a_nb1, b_nb1)
    if a_nb0 not in (1, nb0) or a_nb1 not in (1, nb1) or b_nb0 not in (1, nb0) or b_nb1 not in (1, nb1):
        raise ValueError(f"bmm_nt: batch axes {(a_nb0, a_nb1)} and {(b_nb0, b_nb1)} do not broadcast")
    if out is None:
        if accumulate:
            raise ValueError("bmm_nt: accumulate needs an output tensor")
        pitch = (N + 3) // 4 * 4
        lead = max(a.dim(), b.dim()) - 2
        full = torch.empty((nb0, nb1, M, pitch)[2 - lead:], device=a.device, dtype=torch.float32)
        result = full if pitch == N else full[..., :N]
        d_ptr, ldd, d_bs1, d_bs0 = full.data_ptr(), pitch, M * pitch, nb1 * M * pitch
    else:
        if out.dtype != torch.float32 or not out.is_cuda:
            raise ValueError("bmm_nt: out must be a CUDA float32 tensor")
        d = _describe(out.shape, out.stride())
        d_ptr = out.data_ptr()
        if d is None or d[4] != 0 or d_ptr % 16:
            raise ValueError("bmm_nt: out needs a contiguous last axis, a row pitch that is a multiple of 4 and 16-byte alignment")
        o_nb0, o_nb1, o_m, o_n, _, ldd, d_bs0, d_bs1 = d
        if (o_nb0, o_nb1, o_m, o_n) != (nb0, nb1, M, N):
            raise ValueError(f"bmm_nt: out must have shape {(nb0, nb1, M, N)} (leading 1s optional), got {tuple(out.shape)}")
        result = out
    if nb0 == 1:
        d_bs0 = 0
    if nb1 == 1:
        d_bs1 = 0
    bias_ptr = None
    if bias is not None:
        if bias.dtype != torch.float32 or not bias.is_cuda or bias.numel() != N or not bias.is_contiguous():
            raise ValueError("bmm_nt: bias must be a contiguous float32 CUDA vector of N elements")
        bias_ptr = bias.data_ptr()
    ws_bytes = lib.ob_gemm_f32_workspace_bytes(M, N, K, nb0, nb1) if K >= 2048 and nb0 * nb1 == 1 else 0
    ws = torch.empty(ws_bytes, device=a.device, dtype=torch.uint8) if ws_bytes else None      # split-K partial products
    check(lib.ob_gemm_f32(a_ptr, a_mn, lda, a_bs0, a_bs1, b_ptr, b_mn, ldb, b_bs0, b_bs1, d_ptr, ldd, d_bs0, d_bs1, bias_ptr,
                          scale, accumulate, M, N, K, nb0, nb1, passes, None if ws is None else ws.data_ptr(), ws_bytes,
                          _stream()))
    return result


def column_sums(g2: torch.Tensor) -> torch.Tensor:
    """``g2.sum(0)`` of a contiguous fp32 ``[M, N]`` CUDA matrix on the library's streaming kernel (fixed summation order)."""
    M, N = g2.shape
    if not (g2.is_cuda and g2.dtype == torch.float32 and g2.is_contiguous() and N % 4 == 0 and M > 0 and g2.data_ptr() % 16 == 0):
        return g2.sum(0)
    out = torch.empty((N,), device=g2.device, dtype=torch.float32)
    ws = torch.empty(lib.ob_colsum_workspace_bytes(M, N), device=g2.device, dtype=torch.uint8)
    check(lib.ob_colsum(g2.data_ptr(), M, N, out.data_ptr(), ws.data_ptr(), _stream()))
    return out


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
            gb = column_sums(g2)
        return gx, gw, gb


# A/B switches for measurements (bench.py --torch-nonrouted): "attn", "conv", "linear" in OB_TORCH_NONROUTED route that part of
# the non-routed stack back to torch's own fp32 kernels
DISABLED = set(filter(None, os.environ.get("OB_TORCH_NONROUTED", "").split(",")))


def linear_usable(x: torch.Tensor, weight: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and x.numel() > 0
            and "linear" not in DISABLED)


def linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor = None) -> torch.Tensor:
    """``x @ weight.T + bias`` in fp32 on the tensor cores (3 x tf32 split); ``torch.nn.functional.linear`` elsewhere."""
    if not routes.taken("linear", linear_usable(x, weight), x, "linear" in DISABLED):
        return torch.nn.functional.linear(x, weight, bias)
    return _LinearFn.apply(x, weight, bias)
