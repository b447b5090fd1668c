"""Host side of the B200 quantised linear layer: the reference's interface over libonebit.so.

Mirrors ``onebit_asr/quant.py`` of the reference (citations relative to /root/reference):

* ``QuantizedLinear(in_features, out_features, bias=True)`` with parameters ``weight [out,in]``,
  ``alpha []`` and ``bias [out]`` (same names -> the reference's checkpoints load), the same seeded
  initialisation (quant.py:100-118) and ``forward(x, bitwidth)`` with bitwidth in {1, 2, 32}
  (quant.py:120-127); any other bitwidth raises ``ValueError("bitwidth must be one of {1,2,32}")``.
* ``quantize_weight(W, alpha, bitwidth)`` - the free function (quant.py:95-96) with the clip-window STE and
  the custom d/d-alpha of ``_QuantizeSTE.backward`` (quant.py:72-92).

What differs, by the north_star's spec: activations are quantised per token to int8 (absmax) in front of the
GEMM, the GEMM runs int8 x ternary on tcgen05, weights are kept packed at 2 bits, the backward GEMMs run in
bf16.  The device never syncs with the host (the reference's ``.item()`` per layer, quant.py:75, is gone).
Everything executes in libonebit.so; tensors on the CPU are rejected - there is no fallback.
"""
from __future__ import annotations

import functools
import math
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi
from ._cabi import OB_ALPHA_EFF, OB_ALPHA_RAW, OB_BF16, OB_F32, check, lib

_DTYPE_TAG = {torch.float32: OB_F32, torch.bfloat16: OB_BF16}
_BITWIDTH_ERROR = "bitwidth must be one of {1,2,32}"

# ----------------------------------------------------------------------------------------------
# validity of the cached 2-bit codes
# ----------------------------------------------------------------------------------------------
# A layer keeps the packed codes of its latent weights between forwards (three passes per step use them).  A tensor's
# ``_version`` alone does not tell when they go stale: ``torch.optim.AdamW(fused=True)`` (and any ``.data`` edit) writes the
# parameters without bumping it.  The cache key therefore also carries a process-wide *weight epoch* that moves on
#   * after every ``optimizer.step()`` of any torch optimiser (global post-step hook, registered below),
#   * whenever a backward of the layer has run (gradients exist -> the next forward most likely follows an update),
#   * on ``invalidate_packed_weights()`` - the explicit call for raw ``.data`` / pointer writes outside an optimiser.
# The cost of a spurious bump is one re-quantisation per layer and bitwidth (what the reference does on every forward).
_weight_epoch = [0]


def invalidate_packed_weights() -> None:
    """Declare every cached packed code stale (call after writing latent weights or alpha behind autograd's back)."""
    _weight_epoch[0] += 1
    _ActQuantCache.clear()


def _optimizer_step_hook(optimizer, args, kwargs) -> None:
    _weight_epoch[0] += 1
    _ActQuantCache.clear()


from torch.optim.optimizer import register_optimizer_step_post_hook as _register_step_hook  # noqa: E402

_register_step_hook(_optimizer_step_hook)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device.  Called once per kernel launch, so it goes through
    torch's raw C accessors (0.2 us) instead of building a ``torch.cuda.Stream`` object (10 us) when they exist."""
    if _raw_stream is not None and _raw_device is not None:
        return _raw_stream(_raw_device())
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"onebit_b200: {what} is on '{t.device}'. The quantised path runs only in the sm_100a CUDA library "
            "(no CPU fallback); move the module and its inputs to a B200 device.")


def _tag(t: torch.Tensor) -> int:
    try:
        return _DTYPE_TAG[t.dtype]
    except KeyError:
        raise ValueError(f"onebit_b200: unsupported dtype {t.dtype} (float32 and bfloat16 are supported)") from None


# ----------------------------------------------------------------------------------------------
# thin functional wrappers (one per C entry point) - also what the tests and bench call
# ----------------------------------------------------------------------------------------------
def weight_absmean(W: torch.Tensor) -> torch.Tensor:
    _require_cuda(W, "weight")
    W = W.detach().contiguous().float()
    out = torch.empty((), device=W.device, dtype=torch.float32)
    ws = torch.empty(lib.ob_absmean_workspace_bytes(), device=W.device, dtype=torch.uint8)
    check(lib.ob_weight_absmean(W.data_ptr(), W.numel(), out.data_ptr(), ws.data_ptr(), _stream()))
    return out


def pack_weight(W: torch.Tensor, alpha: torch.Tensor, bitwidth: int, alpha_mode: int = OB_ALPHA_RAW,
                transposed: bool = True):
    """Quantise ``W [N,K]`` and pack to 2 bits: returns (packed [N,K/4] uint8, packed_t [K,N/4] uint8 | None)."""
    _require_cuda(W, "weight")
    if bitwidth not in (1, 2):
        raise ValueError(_BITWIDTH_ERROR)
    N, K = W.shape
    Wc = W.detach().contiguous()
    packed = torch.empty((N, K // 4), device=W.device, dtype=torch.uint8)
    packed_t = torch.empty((K, N // 4), device=W.device, dtype=torch.uint8) if transposed else None
    check(lib.ob_weight_quant_pack(Wc.data_ptr(), alpha.data_ptr(), alpha_mode, N, K, bitwidth, packed.data_ptr(),
                                   packed_t.data_ptr() if transposed else None, _stream()))
    return packed, packed_t


def unpack_codes(packed: torch.Tensor, order: int) -> torch.Tensor:
    R, B = packed.shape
    codes = torch.empty((R, B * 4), device=packed.device, dtype=torch.int8)
    check(lib.ob_unpack_codes(packed.data_ptr(), R, B * 4, order, codes.data_ptr(), _stream()))
    return codes


def act_quant_int8(x: torch.Tensor):
    """Per-token absmax int8 quantiser: returns (q int8 [..., K], scale fp32 [...])."""
    _require_cuda(x, "input")
    K = x.shape[-1]
    x2 = x.detach().reshape(-1, K).contiguous()
    M = x2.shape[0]
    q = torch.empty((M, K), device=x.device, dtype=torch.int8)
    s = torch.empty((M,), device=x.device, dtype=torch.float32)
    check(lib.ob_act_quant_i8(x2.data_ptr(), _tag(x2), M, K, q.data_ptr(), s.data_ptr(), _stream()))
    return q.view(*x.shape[:-1], K), s.view(x.shape[:-1])


def gemm_fwd(q, scale, packed, alpha, bias, N, out_dtype=torch.float32, alpha_mode=OB_ALPHA_RAW):
    M, K = q.shape
    y = torch.empty((M, N), device=q.device, dtype=out_dtype)
    check(lib.ob_gemm_tern_i8_fwd(q.data_ptr(), scale.data_ptr(), packed.data_ptr(), alpha.data_ptr(), alpha_mode,
                                  None if bias is None else bias.data_ptr(), M, N, K, y.data_ptr(),
                                  _DTYPE_TAG[out_dtype], _stream()))
    return y


# ----------------------------------------------------------------------------------------------
# autograd: the whole layer (act quant -> GEMM -> epilogue) is one Function
# ----------------------------------------------------------------------------------------------
class _ActQuantCache:
    """q/k/v projections of the reference's own ``MHSA.forward`` receive the same tensor three times in a row
    (conformer.py:110-112): quantise it once.  (This repo's model quantises inside the fused LayerNorm kernel and hands the
    codes to the three GEMMs explicitly - ``fused.ln_projections`` - and never comes here.)

    One entry, scoped as tightly as the call pattern allows: the key carries the tensor's storage pointer, version, shape,
    dtype AND the stream it was quantised on; the entry is dropped after its third use (q, k, v), at every optimiser step /
    ``invalidate_packed_weights()`` and by ``clear()``, so an activation is not kept alive past the module that made it.  A
    write that bypasses autograd's version counter (``.data`` edits, raw-pointer writes) between two of the three calls is
    not detectable - the same caveat as for the packed weights above."""
    x = None          # keeps the storage alive so (data_ptr, version) cannot be recycled while the entry exists
    key = None
    q = None
    s = None
    uses = 0

    @classmethod
    def get(cls, x2: torch.Tensor):
        key = (x2.data_ptr(), x2._version, tuple(x2.shape), x2.dtype, _stream())
        if cls.key == key:
            q, s = cls.q, cls.s
            cls.uses += 1
            if cls.uses >= 3:
                cls.clear()
            return q, s
        M, K = x2.shape
        q = torch.empty((M, K), device=x2.device, dtype=torch.int8)
        s = torch.empty((M,), device=x2.device, dtype=torch.float32)
        check(lib.ob_act_quant_i8(x2.data_ptr(), _tag(x2), M, K, q.data_ptr(), s.data_ptr(), _stream()))
        cls.x, cls.key, cls.q, cls.s, cls.uses = x2, key, q, s, 1
        return q, s

    @classmethod
    def clear(cls):
        cls.x = cls.key = cls.q = cls.s = None
        cls.uses = 0


class _QuantLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, alpha, bias, bitwidth, packed, packed_t):
        K = x.shape[-1]
        N = weight.shape[0]
        x2 = x.reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        q, s = _ActQuantCache.get(x2)
        M = x2.shape[0]
        y = torch.empty((M, N), device=x.device, dtype=x.dtype)
        check(lib.ob_gemm_tern_i8_fwd(q.data_ptr(), s.data_ptr(), packed.data_ptr(), alpha.data_ptr(), OB_ALPHA_RAW,
                                      None if bias is None else bias.data_ptr(), M, N, K, y.data_ptr(), _tag(y),
                                      _stream()))
        ctx.save_for_backward(q, s, weight, alpha, packed_t)
        ctx.bitwidth = bitwidth
        ctx.has_bias = bias is not None
        ctx.x_shape = x.shape
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, gy):
        q, s, weight, alpha, packed_t = ctx.saved_tensors
        _weight_epoch[0] += 1
        need_x, need_w, need_a, need_b = ctx.needs_input_grad[:4]
        gx, gw, ga, gb = _linear_backward(gy, q, s, weight, alpha, packed_t, ctx.bitwidth, need_x, need_w or need_a,
                                          need_b and ctx.has_bias)
        if gx is not None:
            gx = gx.view(ctx.x_shape)
        return gx, gw, ga, gb, None, None, None


@functools.lru_cache(maxsize=None)
def _colsum_blocks(M: int) -> int:
    return lib.ob_bwd_colsum_blocks(M)


@functools.lru_cache(maxsize=None)
def _dw_ws_bytes(M: int, N: int, K: int) -> int:
    return lib.ob_bwd_dw_workspace_bytes(M, N, K)


# grad_W reads the saved int8 codes directly (CTA-pair kernel, codes converted to bf16 in shared memory) when the layer is at
# least one 256-column tile wide; OB_DW_Q8=0 restores the bf16 copy of q made by the prep kernel (A/B measurements).
DW_Q8 = os.environ.get("OB_DW_Q8", "1") != "0"


def dw_reads_codes(K: int) -> bool:
    return DW_Q8 and K >= 256


def _linear_backward(gy, q, s, weight, alpha, packed_t, bitwidth, need_x, need_w, need_b, gx_out=None):
    """Shared backward of the quantised linear: prep (bf16 casts + column sums), grad_x GEMM, grad_W GEMM + fused
    STE / alpha / bias reductions.  Returns (grad_x [M,K] | None, grad_W | None, grad_alpha | None, grad_bias | None)."""
    M, K = q.shape
    N = weight.shape[0]
    g2 = gy.reshape(M, N)
    if not g2.is_contiguous():
        g2 = g2.contiguous()
    dev, st = g2.device, _stream()
    dys = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
    q8 = dw_reads_codes(K)
    qb = torch.empty((M, K), device=dev, dtype=torch.bfloat16) if need_w and not q8 else None
    colsum = torch.empty((_colsum_blocks(M), N), device=dev, dtype=torch.float32) if need_b else None
    check(lib.ob_bwd_prep(g2.data_ptr(), _tag(g2), s.data_ptr(), q.data_ptr(), M, N, K, dys.data_ptr(),
                          None if qb is None else qb.data_ptr(), None if colsum is None else colsum.data_ptr(), st))
    gx = gw = ga = gb = None
    if need_x:
        gx = gx_out if gx_out is not None else torch.empty((M, K), device=dev, dtype=gy.dtype)
        check(lib.ob_bwd_dx(dys.data_ptr(), s.data_ptr(), packed_t.data_ptr(), alpha.data_ptr(), OB_ALPHA_RAW,
                            M, N, K, gx.data_ptr(), _tag(gx), st))
    if need_w:
        gw = torch.empty((N, K), device=dev, dtype=torch.float32)
        ga = torch.empty((), device=dev, dtype=torch.float32)
        gb = torch.empty((N,), device=dev, dtype=torch.float32) if need_b else None
        nbytes = _dw_ws_bytes(M, N, K)
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        dw = lib.ob_bwd_dw_q8 if q8 else lib.ob_bwd_dw
        check(dw(dys.data_ptr(), (q if q8 else qb).data_ptr(), None if colsum is None else colsum.data_ptr(),
                            weight.data_ptr(), alpha.data_ptr(), OB_ALPHA_RAW, bitwidth, M, N, K,
                            gw.data_ptr(), ga.data_ptr(), None if gb is None else gb.data_ptr(), ws.data_ptr(),
                            nbytes, st))
    elif need_b:
        gb = g2.sum(0, dtype=torch.float32)
    return gx, gw, ga, gb


def _grouped_backward_q8(g2, q, s, weight, alpha, pkt2, pkt1, rows2, gx, need_w, need_b):
    """Layer backward over a stacked batch (rows [0, rows2) at 2 bits, the rest at 1 bit): one prep pass over all rows, grad_x
    per bitwidth group into ``gx`` (or skipped when None), one grad_W launch + finaliser over both groups."""
    M, K = q.shape
    N = weight.shape[0]
    dev, st = g2.device, _stream()
    dys = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
    colsum = torch.empty((_colsum_blocks(M), N), device=dev, dtype=torch.float32) if need_b else None
    check(lib.ob_bwd_prep(g2.data_ptr(), _tag(g2), s.data_ptr(), q.data_ptr(), M, N, K, dys.data_ptr(), None,
                          None if colsum is None else colsum.data_ptr(), st))
    if gx is not None:
        for r0, r1, pkt in ((0, rows2, pkt2), (rows2, M, pkt1)):
            if r1 > r0:
                check(lib.ob_bwd_dx(dys.data_ptr() + 2 * r0 * N, s.data_ptr() + 4 * r0, pkt.data_ptr(), alpha.data_ptr(), OB_ALPHA_RAW,
                                    r1 - r0, N, K, gx.data_ptr() + r0 * K * gx.element_size(), _tag(gx), st))
    gw = ga = gb = None
    if need_w:
        gw = torch.empty((N, K), device=dev, dtype=torch.float32)
        ga = torch.empty((), device=dev, dtype=torch.float32)
        gb = torch.empty((N,), device=dev, dtype=torch.float32) if need_b else None
        nbytes = _dw_ws_bytes(M, N, K)
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        check(lib.ob_bwd_dw_q8_groups(dys.data_ptr(), q.data_ptr(), None if colsum is None else colsum.data_ptr(),
                                      weight.data_ptr(), alpha.data_ptr(), OB_ALPHA_RAW, min(max(rows2, 0), M), M, N, K,
                                      gw.data_ptr(), ga.data_ptr(), None if gb is None else gb.data_ptr(), ws.data_ptr(),
                                      nbytes, st))
    elif need_b:
        gb = g2.sum(0, dtype=torch.float32)
    return gw, ga, gb


_FUSED_SWISH_K = (256, 512, 1024, 2048)
_NO_RNG = (0, 0, 0)
_DROP_DOMAIN = 0x6F6E656269743135   # keeps these Philox streams apart from torch's own use of the same seed


def draw_dropout_stream(device: torch.device, p: float):
    """``inv_keep, (seed, offset, threshold)`` of a fresh Philox stream for one fused dropout (``include/onebit.h``), taken
    from the device's default generator: reproducible under ``torch.manual_seed`` and never repeated, because the
    generator's offset moves on.  An element is dropped iff its 16-bit lane is below ``threshold = round(p * 2^16)``, so
    the realised rate is ``threshold / 65536`` (within 2^-17 of p) and ``inv_keep`` is exact for it.  Nothing here touches
    the device; the backward regenerates the mask from the same triple instead of reading a stored one."""
    index = device.index if device.index is not None else torch.cuda.current_device()
    gen = torch.cuda.default_generators[index]
    seed = (gen.initial_seed() ^ _DROP_DOMAIN) & 0xFFFFFFFFFFFFFFFF
    offset = gen.get_offset()
    gen.set_offset(offset + 4)
    threshold = min(max(int(round(p * 65536.0)), 1), 65535)
    return 65536.0 / (65536 - threshold), (seed, offset, threshold)


class _SwishDropQuantLinearFn(torch.autograd.Function):
    """lin(dropout(swish(h))) with the element-wise chain fused into the activation quantiser (forward) and into one
    kernel behind the grad_x GEMM (backward): the FFN mid-section of conformer.py:36-39."""

    @staticmethod
    def forward(ctx, h, weight, alpha, bias, bitwidth, packed, packed_t, keep, inv_keep, rng):
        K = h.shape[-1]
        N = weight.shape[0]
        h2 = h.reshape(-1, K)
        if not h2.is_contiguous():
            h2 = h2.contiguous()
        M = h2.shape[0]
        q = torch.empty((M, K), device=h.device, dtype=torch.int8)
        s = torch.empty((M,), device=h.device, dtype=torch.float32)
        st = _stream()
        check(lib.ob_swish_drop_quant(h2.data_ptr(), None if keep is None else keep.data_ptr(), inv_keep, *rng, M, K,
                                      q.data_ptr(), s.data_ptr(), st))
        y = torch.empty((M, N), device=h.device, dtype=h.dtype)
        check(lib.ob_gemm_tern_i8_fwd(q.data_ptr(), s.data_ptr(), packed.data_ptr(), alpha.data_ptr(), OB_ALPHA_RAW,
                                      None if bias is None else bias.data_ptr(), M, N, K, y.data_ptr(), _tag(y), st))
        if keep is None:
            ctx.save_for_backward(h2, q, s, weight, alpha, packed_t)
        else:
            ctx.save_for_backward(h2, q, s, weight, alpha, packed_t, keep)
        ctx.bitwidth, ctx.has_bias, ctx.h_shape, ctx.inv_keep, ctx.rng = bitwidth, bias is not None, h.shape, inv_keep, rng
        return y.view(*h.shape[:-1], N)

    @staticmethod
    def backward(ctx, gy):
        saved = ctx.saved_tensors
        _weight_epoch[0] += 1
        h2, q, s, weight, alpha, packed_t = saved[:6]
        keep = saved[6] if len(saved) > 6 else None
        need_h, need_w, need_a, need_b = ctx.needs_input_grad[:4]
        gz, gw, ga, gb = _linear_backward(gy, q, s, weight, alpha, packed_t, ctx.bitwidth, need_h, need_w or need_a,
                                          need_b and ctx.has_bias)
        gh = None
        if need_h:
            gh = torch.empty_like(h2)
            check(lib.ob_swish_drop_bwd(gz.data_ptr(), h2.data_ptr(), None if keep is None else keep.data_ptr(),
                                        ctx.inv_keep, *ctx.rng, h2.numel(), gh.data_ptr(), _stream()))
            gh = gh.view(ctx.h_shape)
        return gh, gw, ga, gb, None, None, None, None, None, None


class _GroupedQuantLinearFn(torch.autograd.Function):
    """One layer applied to a stack of passes that differ in bitwidth: token rows ``[0, rows2)`` use the 2-bit codes,
    rows ``[rows2, M)`` the 1-bit codes (the co-training passes of train.py:83-103 evaluated side by side).  One
    activation quantiser over all rows, one GEMM per group; the backward runs the layer backward per group and adds the
    parameter gradients here instead of through three separate autograd accumulations.  ``pre`` = None, or
    ``(inv_keep, rng)`` to apply swish + dropout in front of the quantiser (FFN mid-section)."""

    @staticmethod
    def forward(ctx, x, weight, alpha, bias, rows2, pk2, pkt2, pk1, pkt1, pre):
        K = x.shape[-1]
        N = weight.shape[0]
        x2 = x.reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        M = x2.shape[0]
        st = _stream()
        if pre is None:
            q, s = _ActQuantCache.get(x2)
        else:
            q = torch.empty((M, K), device=x.device, dtype=torch.int8)
            s = torch.empty((M,), device=x.device, dtype=torch.float32)
            check(lib.ob_swish_drop_quant(x2.data_ptr(), None, pre[0], *pre[1], M, K, q.data_ptr(), s.data_ptr(), st))
        y = torch.empty((M, N), device=x.device, dtype=x.dtype)
        bias_ptr = None if bias is None else bias.data_ptr()
        esz = y.element_size()
        for r0, r1, pk in ((0, rows2, pk2), (rows2, M, pk1)):
            if r1 > r0:
                check(lib.ob_gemm_tern_i8_fwd(q.data_ptr() + r0 * K, s.data_ptr() + 4 * r0, pk.data_ptr(), alpha.data_ptr(),
                                              OB_ALPHA_RAW, bias_ptr, r1 - r0, N, K, y.data_ptr() + r0 * N * esz, _tag(y), st))
        ctx.save_for_backward(q, s, weight, alpha, pkt2, pkt1, *(() if pre is None else (x2,)))
        ctx.rows2, ctx.has_bias, ctx.x_shape, ctx.pre = rows2, bias is not None, x.shape, pre
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, gy):
        q, s, weight, alpha, pkt2, pkt1 = ctx.saved_tensors[:6]
        _weight_epoch[0] += 1
        M, K = q.shape
        N = weight.shape[0]
        need_x, need_w, need_a, need_b = ctx.needs_input_grad[:4]
        g2 = gy.reshape(M, N)
        if not g2.is_contiguous():
            g2 = g2.contiguous()
        gx = torch.empty((M, K), device=g2.device, dtype=gy.dtype) if need_x else None
        gw = ga = gb = None
        if dw_reads_codes(K):                              # one prep pass, one grad_W launch + finaliser over both groups
            gw, ga, gb = _grouped_backward_q8(g2, q, s, weight, alpha, pkt2, pkt1, ctx.rows2, gx, need_w or need_a,
                                              need_b and ctx.has_bias)
        else:                                              # narrow layers: per group, parameter gradients added here
            for r0, r1, pkt, bw in ((0, ctx.rows2, pkt2, 2), (ctx.rows2, M, pkt1, 1)):
                if r1 <= r0:
                    continue
                _, gw_g, ga_g, gb_g = _linear_backward(g2[r0:r1], q[r0:r1], s[r0:r1], weight, alpha, pkt, bw, need_x,
                                                       need_w or need_a, need_b and ctx.has_bias,
                                                       gx_out=None if gx is None else gx[r0:r1])
                gw = gw_g if gw is None or gw_g is None else gw.add_(gw_g)
                ga = ga_g if ga is None or ga_g is None else ga.add_(ga_g)
                gb = gb_g if gb is None or gb_g is None else gb.add_(gb_g)
        if need_x and ctx.pre is not None:                # back through dropout and swish
            h2 = ctx.saved_tensors[6]
            gh = torch.empty_like(h2)
            check(lib.ob_swish_drop_bwd(gx.data_ptr(), h2.data_ptr(), None, ctx.pre[0], *ctx.pre[1], h2.numel(), gh.data_ptr(),
                                        _stream()))
            gx = gh
        if gx is not None:
            gx = gx.view(ctx.x_shape)
        return gx, gw, ga, gb, None, None, None, None, None, None


class _QuantizeWeightFn(torch.autograd.Function):
    """Dense W_hat = alpha * Q with the reference's STE (the free function ``quantize_weight``)."""

    @staticmethod
    def forward(ctx, W, alpha, bitwidth):
        Wc = W.contiguous()
        out = torch.empty_like(Wc)
        check(lib.ob_weight_quant_dense(Wc.data_ptr(), alpha.data_ptr(), OB_ALPHA_EFF, Wc.numel(), bitwidth,
                                        out.data_ptr(), _stream()))
        ctx.save_for_backward(Wc, alpha)
        ctx.bitwidth = bitwidth
        return out

    @staticmethod
    def backward(ctx, g):
        W, alpha = ctx.saved_tensors
        g = g.contiguous()
        gw = torch.empty_like(W)
        ga = torch.empty((), device=W.device, dtype=torch.float32)
        ws = torch.empty(lib.ob_ste_workspace_bytes(W.numel()), device=W.device, dtype=torch.uint8)
        check(lib.ob_weight_ste_backward(g.data_ptr(), W.data_ptr(), alpha.data_ptr(), OB_ALPHA_EFF, W.numel(),
                                         ctx.bitwidth, gw.data_ptr(), ga.data_ptr(), ws.data_ptr(), _stream()))
        return gw, ga.view(alpha.shape), None


def quantize_weight(W: torch.Tensor, alpha: torch.Tensor, bitwidth: int) -> torch.Tensor:
    """``alpha * Pi(clip(W/alpha))`` with the clip-window STE; ``alpha`` is used as given (quant.py:95-96)."""
    if bitwidth == 32:
        return W
    if bitwidth not in (1, 2):
        raise ValueError(_BITWIDTH_ERROR)
    _require_cuda(W, "weight")
    if W.dtype != torch.float32 or alpha.dtype != torch.float32:
        raise ValueError("onebit_b200.quantize_weight: W and alpha must be float32")
    return _QuantizeWeightFn.apply(W, alpha, bitwidth)


class PackedCodeArena:
    """Persistent 2-bit code buffers of MANY layers, refreshed by one kernel launch.

    The co-training step (train.py:83-103) needs the 2-bit and the 1-bit codes of every routed layer once per optimiser step.
    Layer by layer that is 2 launches x 108 layers; with an arena the first ``packed_weight()`` call after the weights changed
    (same staleness rule as the per-layer cache: weight epoch + tensor versions) re-quantises ALL registered layers with
    ``ob_weight_quant_pack_multi`` - each W read once, both bitwidths, both layouts - into buffers that stay allocated.  The
    buffers are rewritten in place, in stream order: a backward that still holds them sees the same codes unless the weights
    changed in between, which no training loop does between a forward and its backward."""

    def __init__(self, layers):
        import struct
        self.layers = [l for l in layers if l.in_features % 64 == 0 and l.out_features % 64 == 0]
        if not self.layers:
            raise ValueError("PackedCodeArena: no layer with feature counts that are multiples of 64")
        dev = self.layers[0].weight.device
        if dev.type != "cuda":
            raise RuntimeError("PackedCodeArena: move the model to a B200 device first")
        self.buffers, blob, tile0 = [], b"", 0
        for l in self.layers:
            N, K = l.out_features, l.in_features
            bufs = tuple(torch.empty(shape, device=dev, dtype=torch.uint8) for shape in ((N, K // 4), (K, N // 4), (N, K // 4), (K, N // 4)))
            self.buffers.append(bufs)
            blob += struct.pack("<6Q4i", l.weight.data_ptr(), l.alpha.data_ptr(), *(b.data_ptr() for b in bufs), N, K, tile0, 0)
            tile0 += (N // 64) * (K // 64)
        self.total_tiles = tile0
        self.descs = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
        self.ptrs = [(l.weight.data_ptr(), l.alpha.data_ptr()) for l in self.layers]
        self.stamps = [None] * len(self.layers)           # (weight epoch, weight version, alpha version) at the last repack
        self.repacks = 0
        for i, l in enumerate(self.layers):
            l._arena, l._arena_index = self, i

    def codes(self, index: int, bitwidth: int):
        layer = self.layers[index]
        w, a = layer.weight, layer.alpha
        if (w.data_ptr(), a.data_ptr()) != self.ptrs[index]:
            raise RuntimeError("PackedCodeArena: a parameter was re-allocated (module moved or re-created); build a new arena")
        if self.stamps[index] != (_weight_epoch[0], w._version, a._version):      # this layer's codes are stale: redo them all
            check(lib.ob_weight_quant_pack_multi(self.descs.data_ptr(), len(self.layers), self.total_tiles, OB_ALPHA_RAW, _stream()))
            epoch = _weight_epoch[0]
            self.stamps = [(epoch, l.weight._version, l.alpha._version) for l in self.layers]
            self.repacks += 1
        b = self.buffers[index]
        return (b[0], b[1]) if bitwidth == 2 else (b[2], b[3])


class QuantizedLinear(nn.Module):
    """Linear layer with 1-bit / 2-bit (ternary) weights chosen per call; drop-in for quant.py:99-127."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        w = torch.empty(out_features, in_features)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))      # same RNG draw as the reference (quant.py:104)
        w.mul_(2.0)                                       # quant.py:107-108
        self.weight = nn.Parameter(w)
        self.alpha = nn.Parameter(w.abs().mean())         # 0-dim, learnable (quant.py:111-113)
        self.bias = nn.Parameter(torch.zeros(out_features)) if bias else None
        self._packed = {}                                 # bitwidth -> (key, packed, packed_t)
        self._arena, self._arena_index = None, -1         # set by PackedCodeArena: codes of many layers from one launch

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}"

    def reset_alpha_absmean(self) -> None:
        """alpha <- mean|W| computed on the device (BitNet-style absmean re-scale)."""
        with torch.no_grad():
            self.alpha.copy_(weight_absmean(self.weight))

    def packed_weight(self, bitwidth: int):
        """2-bit packed codes for ``bitwidth``, cached while the latent weights are known to be unchanged (tensor versions
        and the weight epoch above): one quantiser launch per optimiser step and bitwidth, instead of one per forward as
        in quant.py:124."""
        if self._arena is not None:
            return self._arena.codes(self._arena_index, bitwidth)
        w, a = self.weight, self.alpha
        key = (_weight_epoch[0], w._version, a._version, w.data_ptr(), a.data_ptr())
        hit = self._packed.get(bitwidth)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        packed, packed_t = pack_weight(w, a, bitwidth, OB_ALPHA_RAW, transposed=True)
        self._packed[bitwidth] = (key, packed, packed_t)
        return packed, packed_t

    def forward(self, x: torch.Tensor, bitwidth: int) -> torch.Tensor:
        if bitwidth == 32:                                # full precision bypass (quant.py:121-122)
            return F.linear(x, self.weight, self.bias)
        if bitwidth not in (1, 2):
            raise ValueError(_BITWIDTH_ERROR)
        _require_cuda(x, "input")
        _require_cuda(self.weight, "weight")
        if x.dtype not in _DTYPE_TAG:
            raise ValueError(f"onebit_b200: unsupported input dtype {x.dtype}")
        if self.weight.dtype != torch.float32 or self.alpha.dtype != torch.float32:
            raise ValueError("onebit_b200: the latent weight, alpha and bias must stay float32 (the quantiser works on the fp32 "
                             f"latent weights, quant.py:49); got weight {self.weight.dtype}")
        if x.shape[-1] != self.in_features:
            raise ValueError(f"onebit_b200: expected last dimension {self.in_features}, got {x.shape[-1]}")
        if x.numel() == 0:                                # empty batch: nothing to launch (F.linear returns an empty tensor too)
            return x.new_zeros(*x.shape[:-1], self.out_features) + 0.0 * (self.weight.sum() + self.alpha)
        if self.in_features % 64 or self.out_features % 64:
            return self._forward_padded(x, bitwidth)
        packed, packed_t = self.packed_weight(bitwidth)
        return _QuantLinearFn.apply(x, self.weight, self.alpha, self.bias, bitwidth, packed, packed_t)

    def _forward_padded(self, x: torch.Tensor, bitwidth: int) -> torch.Tensor:
        """Feature counts that are not multiples of 64 (the tensor-core kernels' granularity): zero-pad the operands
        to the next multiple, run the same kernels, slice the result.  Zero activations in the padded columns and
        zero upstream gradients in the padded rows make the padding invisible to y and to every gradient (incl.
        alpha); the packed codes of the padded weight are rebuilt per call (this is the rare path)."""
        K, N = self.in_features, self.out_features
        Kp, Np = -(-K // 64) * 64, -(-N // 64) * 64
        w = F.pad(self.weight, (0, Kp - K, 0, Np - N))
        b = None if self.bias is None else F.pad(self.bias, (0, Np - N))
        xp = F.pad(x, (0, Kp - K))
        packed, packed_t = pack_weight(w, self.alpha, bitwidth, OB_ALPHA_RAW, transposed=True)
        y = _QuantLinearFn.apply(xp, w, self.alpha, b, bitwidth, packed, packed_t)
        return y[..., :N]


    def grouped_usable(self, x: torch.Tensor, swish: bool = False) -> bool:
        return (x.is_cuda and x.dtype == torch.float32 and self.weight.dtype == torch.float32 and x.numel() > 0
                and self.in_features % 64 == 0 and self.out_features % 64 == 0
                and (not swish or self.in_features in _FUSED_SWISH_K))

    def forward_grouped(self, x: torch.Tensor, rows2: int, swish_dropout=None) -> torch.Tensor:
        """``forward`` over stacked passes: the first ``rows2`` token rows at 2 bits, the others at 1 bit (see
        ``_GroupedQuantLinearFn``).  ``swish_dropout = (p, training)`` applies the FFN mid-section in front."""
        M = x.numel() // self.in_features
        if not 0 <= rows2 <= M:
            raise ValueError(f"onebit_b200: rows2 = {rows2} outside [0, {M}]")
        pk2, pkt2 = self.packed_weight(2) if rows2 > 0 else (None, None)
        pk1, pkt1 = self.packed_weight(1) if rows2 < M else (None, None)
        pk2, pkt2 = (pk1, pkt1) if pk2 is None else (pk2, pkt2)      # unused group: any valid tensor
        pk1, pkt1 = (pk2, pkt2) if pk1 is None else (pk1, pkt1)
        pre = None
        if swish_dropout is not None:
            p, training = swish_dropout
            pre = draw_dropout_stream(x.device, p) if (training and p > 0.0) else (1.0, _NO_RNG)
        return _GroupedQuantLinearFn.apply(x, self.weight, self.alpha, self.bias, rows2, pk2, pkt2, pk1, pkt1, pre)

    def forward_swish_dropout(self, h: torch.Tensor, bitwidth: int, p: float = 0.0, training: bool = False,
                              keep: torch.Tensor = None) -> torch.Tensor:
        """``self(dropout(swish(h)), bitwidth)`` (conformer.py:37-39) with swish, dropout and the activation quantiser
        fused into one kernel.  ``keep`` (bool [.., K]) overrides the sampled dropout mask (tests)."""
        fused = (bitwidth in (1, 2) and h.is_cuda and h.dtype == torch.float32 and self.in_features in _FUSED_SWISH_K)
        if not fused:
            z = h * torch.sigmoid(h)
            if keep is not None:
                z = z * keep.to(z.dtype) * (1.0 / (1.0 - p))
            else:
                z = F.dropout(z, p, training)
            return self.forward(z, bitwidth)
        _require_cuda(self.weight, "weight")
        inv_keep, rng = 1.0, _NO_RNG
        if keep is not None:
            keep, inv_keep = keep.reshape(-1, self.in_features).contiguous(), 1.0 / (1.0 - p)
        elif training and p > 0.0:
            inv_keep, rng = draw_dropout_stream(h.device, p)                    # mask generated inside the kernels
        packed, packed_t = self.packed_weight(bitwidth)
        return _SwishDropQuantLinearFn.apply(h, self.weight, self.alpha, self.bias, bitwidth, packed, packed_t, keep,
                                             inv_keep, rng)


BitLinear = QuantizedLinear   # the north_star's name for the same layer


def install_as_reference_quant() -> None:
    """Register this module as ``quant`` so the reference's flat ``from quant import QuantizedLinear``
    (conformer.py:12) resolves to the B200 layer.  Call before importing the reference's conformer."""
    sys.modules["quant"] = sys.modules[__name__]
