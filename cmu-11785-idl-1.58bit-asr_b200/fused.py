"""The routed modules of the Conformer block as fused chains around the quantised layer.

The reference's ``FeedForwardModule.forward`` (conformer.py:34-45) and ``MHSA.forward`` (conformer.py:105-138) are LayerNorm ->
routed projection(s) -> ... -> routed projection -> dropout -> frame mask -> residual.  With the layer on tensor cores those
element-wise neighbours are what is left to pay for (the model's shapes are HBM-bound), so here they live inside the layer's
own kernels:

  forward   LayerNorm + int8 quantiser            one kernel, writes codes + scales only (``ob_layernorm_quant_fwd``)
            swish + dropout + int8 quantiser      one kernel (``ob_swish_drop_quant``)
            GEMM + dequant + bias + dropout + frame mask + residual      the GEMM's epilogue (``ob_gemm_tern_i8_fwd_tail``)
  backward  tail / swish-dropout backward         folded into the bf16 cast in front of the backward GEMMs (``ob_bwd_prep_fused``)
            sum of the q/k/v input gradients, LayerNorm backward, + residual gradient      one kernel (``ob_layernorm_bwd3``)

Three autograd Functions cover the two modules: ``_FfnFn`` (the whole half-step feed-forward), ``_LnProjFn`` (LayerNorm + the
q/k/v projections of the attention) and ``_ProjTailFn`` (out_proj + tail).  All of them take ``rows2``: token rows
``[0, rows2)`` use the layer's 2-bit codes, the others its 1-bit codes - one launch per bitwidth group, which is how the three
co-training passes of train.py:83-103 run side by side on one stacked batch (``asr_model.StackedBits``).

Same values as the unfused chain: int8 codes and the forward output are bit-identical (tests/test_gpu_fused.py); gradients
agree to the layer's bf16 tolerance (the fused backward rounds the same products to bf16 at the same place).
"""
from __future__ import annotations

import functools

import numpy as np
import torch

from ._cabi import OB_ALPHA_RAW, OB_F32, OB_PREP_SWISH, OB_PREP_TAIL, check, lib
from .quant import (_FUSED_SWISH_K, _NO_RNG, _colsum_blocks, _dw_ws_bytes, _stream, _weight_epoch, draw_dropout_stream,
                    dw_reads_codes)

LN_WIDTHS = (128, 256, 512, 1024)


@functools.lru_cache(maxsize=None)
def _ln_ws_bytes(C: int) -> int:
    return lib.ob_layernorm_bwd_workspace_bytes(C)


class LayerCodes:
    """What the kernels need from one routed layer for one call: parameters + packed codes per bitwidth group."""
    __slots__ = ("weight", "alpha", "bias", "pk2", "pkt2", "pk1", "pkt1", "N", "K")

    def __init__(self, layer, rows2: int, M: int):
        self.alpha, self.bias = layer.alpha, layer.bias
        self.N, self.K = layer.out_features, layer.in_features
        if hasattr(layer, "packed_weight"):                          # trainable layer: latent weights + cached codes
            self.weight = layer.weight
            pk2, pkt2 = layer.packed_weight(2) if rows2 > 0 else (None, None)
            pk1, pkt1 = layer.packed_weight(1) if rows2 < M else (None, None)
        else:                                                        # inference.PackedQuantizedLinear: one frozen bitwidth
            self.weight = None
            want = 2 if rows2 > 0 else 1
            if (0 < rows2 < M) or layer.bitwidth != want:
                raise ValueError(f"this layer was packed at bitwidth {layer.bitwidth}, called with rows2={rows2} of {M}")
            pk2 = pk1 = layer.packed
            pkt2 = pkt1 = None
        self.pk2, self.pkt2 = (pk1, pkt1) if pk2 is None else (pk2, pkt2)        # unused group: any valid tensor
        self.pk1, self.pkt1 = (pk2, pkt2) if pk1 is None else (pk1, pkt1)


class TailSpec:
    """``x + scale * dropout(y) * frame_mask`` (conformer.py:41-45, 133-138): float row mask (or None), scale, Philox stream."""
    __slots__ = ("rowmask", "factor", "rng")

    def __init__(self, device, rowmask, scale: float, p: float, training: bool):
        inv_keep, self.rng = draw_dropout_stream(device, p) if (training and p > 0.0) else (1.0, _NO_RNG)
        self.factor = float(np.float32(scale) * np.float32(inv_keep))
        self.rowmask = rowmask


def _ptr(t):
    return None if t is None else t.data_ptr()


def _groups(rows2: int, M: int):
    return [g for g in ((0, rows2, 2), (rows2, M, 1)) if g[1] > g[0]]


# ---------------------------------------------------------------------------------------------- kernels, forward
def _ln_quant(x2, ln_w, ln_b, eps):
    M, C = x2.shape
    q = torch.empty((M, C), device=x2.device, dtype=torch.int8)
    s = torch.empty((M,), device=x2.device, dtype=torch.float32)
    stats = torch.empty((2, M), device=x2.device, dtype=torch.float32)
    sp = stats.data_ptr()
    check(lib.ob_layernorm_quant_fwd(x2.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), eps, M, C, q.data_ptr(), s.data_ptr(),
                                     sp, sp + 4 * M, _stream()))
    return q, s, stats


def _act_quant(x2):
    M, K = x2.shape
    q = torch.empty((M, K), device=x2.device, dtype=torch.int8)
    s = torch.empty((M,), device=x2.device, dtype=torch.float32)
    check(lib.ob_act_quant_i8(x2.data_ptr(), OB_F32, M, K, q.data_ptr(), s.data_ptr(), _stream()))
    return q, s


def _gemm(q, s, lc: LayerCodes, rows2: int, tail: TailSpec = None, resid=None):
    """y [M, N] fp32 = layer(q, s), one launch per bitwidth group; with ``tail`` the module tail runs in the epilogue."""
    M, K = q.shape
    N = lc.N
    y = torch.empty((M, N), device=q.device, dtype=torch.float32)
    st, bias = _stream(), _ptr(lc.bias)
    qp, sp, yp, ap = q.data_ptr(), s.data_ptr(), y.data_ptr(), lc.alpha.data_ptr()
    for r0, r1, bw in _groups(rows2, M):
        pk = (lc.pk2 if bw == 2 else lc.pk1).data_ptr()
        if tail is None:
            check(lib.ob_gemm_tern_i8_fwd(qp + r0 * K, sp + 4 * r0, pk, ap, OB_ALPHA_RAW, bias, r1 - r0, N, K, yp + 4 * r0 * N,
                                          OB_F32, st))
        else:
            check(lib.ob_gemm_tern_i8_fwd_tail(qp + r0 * K, sp + 4 * r0, pk, ap, OB_ALPHA_RAW, bias, r1 - r0, N, K,
                                               resid.data_ptr() + 4 * r0 * N, _ptr(tail.rowmask), tail.factor, *tail.rng, r0,
                                               yp + 4 * r0 * N, st))
    return y


# ---------------------------------------------------------------------------------------------- kernels, backward
def _layer_backward(g, prep, q, s, qb, lc: LayerCodes, rows2: int, need_x: bool, need_w: bool, need_b: bool):
    """Backward of one routed layer over both bitwidth groups.

    ``g`` [M, N] fp32 is the gradient that arrives from the op after the layer; ``prep`` says what that op was:
    None (the layer's own output gradient), ("tail", TailSpec) or ("swish", h, inv_keep, rng).  ``qb``: bf16 copy of q if an
    earlier layer with the same input already made it, else None (made here when the weight gradient is needed).
    Returns (grad_x [M, K] | None, grad_W, grad_alpha, grad_bias, qb)."""
    M, K = q.shape
    N = lc.N
    dev, st = g.device, _stream()
    dys = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
    q8 = dw_reads_codes(K)                              # grad_W converts the int8 codes in shared memory: no bf16 copy of q
    gx = torch.empty((M, K), device=dev, dtype=torch.float32) if need_x else None
    gw = ga = gb = None
    gp, qp, sp, dp, ap = g.data_ptr(), q.data_ptr(), s.data_ptr(), dys.data_ptr(), lc.alpha.data_ptr()

    def run_prep(r0, Mg, qb_out, colsum):
        if prep is None:
            check(lib.ob_bwd_prep(gp + 4 * r0 * N, OB_F32, sp + 4 * r0, qp + r0 * K, Mg, N, K, dp + 2 * r0 * N, qb_out,
                                  _ptr(colsum), st))
        elif prep[0] == "tail":
            t = prep[1]
            check(lib.ob_bwd_prep_fused(gp + 4 * r0 * N, OB_PREP_TAIL, _ptr(t.rowmask), None, t.factor, *t.rng, r0, sp + 4 * r0,
                                        qp + r0 * K, Mg, N, K, dp + 2 * r0 * N, qb_out, _ptr(colsum), st))
        else:
            _, h, inv_keep, rng = prep
            check(lib.ob_bwd_prep_fused(gp + 4 * r0 * N, OB_PREP_SWISH, None, h.data_ptr() + 4 * r0 * N, inv_keep, *rng, r0,
                                        sp + 4 * r0, qp + r0 * K, Mg, N, K, dp + 2 * r0 * N, qb_out, _ptr(colsum), st))

    def run_dx(r0, Mg, bw):
        pkt = (lc.pkt2 if bw == 2 else lc.pkt1).data_ptr()
        check(lib.ob_bwd_dx(dp + 2 * r0 * N, sp + 4 * r0, pkt, ap, OB_ALPHA_RAW, Mg, N, K, gx.data_ptr() + 4 * r0 * K, OB_F32, st))

    if q8:
        # one pass over the upstream gradient for all rows (nothing in it depends on the bitwidth), grad_x per bitwidth
        # group, then ONE grad_W launch + finaliser over both groups: no per-group partial results to add up
        colsum = torch.empty((_colsum_blocks(M), N), device=dev, dtype=torch.float32) if need_b else None
        run_prep(0, M, None, colsum)
        if need_x:
            for r0, r1, bw in _groups(rows2, M):
                run_dx(r0, r1 - r0, bw)
        if need_w:
            gw = torch.empty((N, K), device=dev, dtype=torch.float32)
            ga = torch.empty((), device=dev, dtype=torch.float32)
            gb = torch.empty((N,), device=dev, dtype=torch.float32) if need_b else None
            nbytes = _dw_ws_bytes(M, N, K)
            ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            check(lib.ob_bwd_dw_q8_groups(dp, qp, _ptr(colsum), lc.weight.data_ptr(), ap, OB_ALPHA_RAW, min(max(rows2, 0), M), M, N, K,
                                          gw.data_ptr(), ga.data_ptr(), _ptr(gb), ws.data_ptr(), nbytes, st))
        return gx, gw, ga, gb, qb

    make_qb = need_w and qb is None
    if make_qb:
        qb = torch.empty((M, K), device=dev, dtype=torch.bfloat16)
    for r0, r1, bw in _groups(rows2, M):
        Mg = r1 - r0
        colsum = torch.empty((_colsum_blocks(Mg), N), device=dev, dtype=torch.float32) if need_b else None
        run_prep(r0, Mg, qb.data_ptr() + 2 * r0 * K if make_qb else None, colsum)
        if need_x:
            run_dx(r0, Mg, bw)
        if need_w:
            gw_g = torch.empty((N, K), device=dev, dtype=torch.float32)
            ga_g = torch.empty((), device=dev, dtype=torch.float32)
            gb_g = torch.empty((N,), device=dev, dtype=torch.float32) if need_b else None
            nbytes = _dw_ws_bytes(Mg, N, K)
            ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            check(lib.ob_bwd_dw(dp + 2 * r0 * N, qb.data_ptr() + 2 * r0 * K, _ptr(colsum), lc.weight.data_ptr(), ap, OB_ALPHA_RAW,
                                bw, Mg, N, K, gw_g.data_ptr(), ga_g.data_ptr(), _ptr(gb_g), ws.data_ptr(), nbytes, st))
            gw = gw_g if gw is None else gw.add_(gw_g)
            ga = ga_g if ga is None else ga.add_(ga_g)
            gb = gb_g if gb is None or gb_g is None else gb.add_(gb_g)
    return gx, gw, ga, gb, qb


def _ln_backward(dys, x2, stats, ln_w, resid):
    """dx, d-gamma, d-beta of LayerNorm for the summed upstream gradients ``dys`` (1-3 tensors) + the residual gradient."""
    M, C = x2.shape
    dx = torch.empty_like(x2)
    dparams = torch.empty((2, C), device=x2.device, dtype=torch.float32)
    ws = torch.empty(_ln_ws_bytes(C), device=x2.device, dtype=torch.uint8)
    sp, dp = stats.data_ptr(), dparams.data_ptr()
    d = [t.data_ptr() for t in dys] + [None, None]
    check(lib.ob_layernorm_bwd3(d[0], d[1], d[2], x2.data_ptr(), sp, sp + 4 * M, ln_w.data_ptr(), _ptr(resid), M, C,
                                dx.data_ptr(), dp, dp + 4 * C, ws.data_ptr(), _stream()))
    return dx, dparams[0], dparams[1]


def _rows(t, width):
    t2 = t.reshape(-1, width)
    return t2 if t2.is_contiguous() else t2.contiguous()


# ---------------------------------------------------------------------------------------------- the feed-forward module
class FfnSpec:
    __slots__ = ("eps", "rows2", "lc1", "lc2", "mid_inv_keep", "mid_rng", "tail")


class _FfnFn(torch.autograd.Function):
    """x + 0.5 * drop(lin2(drop(swish(lin1(norm(x)))))) * frame_mask   (conformer.py:34-45) in four kernels forward."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, w1, a1, b1, w2, a2, b2, spec: FfnSpec):
        C = x.shape[-1]
        x2 = _rows(x, C)
        q1, s1, stats = _ln_quant(x2, ln_w, ln_b, spec.eps)
        h = _gemm(q1, s1, spec.lc1, spec.rows2)
        M, F = h.shape
        q2 = torch.empty((M, F), device=x.device, dtype=torch.int8)
        s2 = torch.empty((M,), device=x.device, dtype=torch.float32)
        check(lib.ob_swish_drop_quant(h.data_ptr(), None, spec.mid_inv_keep, *spec.mid_rng, M, F, q2.data_ptr(), s2.data_ptr(),
                                      _stream()))
        out = _gemm(q2, s2, spec.lc2, spec.rows2, spec.tail, x2)
        ctx.save_for_backward(x2, stats, ln_w, q1, s1, h, q2, s2)
        ctx.spec, ctx.x_shape = spec, x.shape
        return out.view(x.shape)

    @staticmethod
    def backward(ctx, g):
        x2, stats, ln_w, q1, s1, h, q2, s2 = ctx.saved_tensors
        spec = ctx.spec
        _weight_epoch[0] += 1
        need = ctx.needs_input_grad
        g2 = _rows(g, x2.shape[1])
        gz, gw2, ga2, gb2, _ = _layer_backward(g2, ("tail", spec.tail), q2, s2, None, spec.lc2, spec.rows2, True,
                                               need[6] or need[7], need[8] and spec.lc2.bias is not None)
        g_ln, gw1, ga1, gb1, _ = _layer_backward(gz, ("swish", h, spec.mid_inv_keep, spec.mid_rng), q1, s1, None, spec.lc1,
                                                 spec.rows2, True, need[3] or need[4], need[5] and spec.lc1.bias is not None)
        gx, dgamma, dbeta = _ln_backward([g_ln], x2, stats, ln_w, g2)
        return gx.view(ctx.x_shape), dgamma, dbeta, gw1, ga1, gb1, gw2, ga2, gb2, None


def ffn_usable(x, lin1, lin2) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and x.numel() > 0 and x.shape[-1] in LN_WIDTHS
            and lin1.in_features == x.shape[-1] and lin2.out_features == x.shape[-1] and lin1.out_features == lin2.in_features
            and lin2.in_features in _FUSED_SWISH_K and lin1.out_features % 256 == 0 and x.shape[-1] % 64 == 0
            and (not hasattr(lin1, "weight") or lin1.weight.dtype == torch.float32))


def ffn_forward(x, ln, lin1, lin2, rows2: int, rowmask, p: float, training: bool, scale: float = 0.5):
    """The half-step feed-forward module on ``x [..., C]``; ``ln`` = nn.LayerNorm, ``lin1`` / ``lin2`` routed layers."""
    M = x.numel() // x.shape[-1]
    spec = FfnSpec()
    spec.eps, spec.rows2 = ln.eps, rows2
    spec.lc1, spec.lc2 = LayerCodes(lin1, rows2, M), LayerCodes(lin2, rows2, M)
    spec.mid_inv_keep, spec.mid_rng = draw_dropout_stream(x.device, p) if (training and p > 0.0) else (1.0, _NO_RNG)
    spec.tail = TailSpec(x.device, rowmask, scale, p, training)
    w1 = lin1.weight if hasattr(lin1, "weight") else None
    w2 = lin2.weight if hasattr(lin2, "weight") else None
    return _FfnFn.apply(x, ln.weight, ln.bias, w1, lin1.alpha, lin1.bias, w2, lin2.alpha, lin2.bias, spec)


# ---------------------------------------------------------------------------------------------- LayerNorm + projections
class _LnProjFn(torch.autograd.Function):
    """(x, proj_1(norm(x)), ..., proj_n(norm(x))) for up to three routed projections of the same normalised tensor
    (conformer.py:109-112).  x is handed through so that the module's residual use of it (the tail of out_proj) sends its
    gradient back HERE, where the LayerNorm backward adds it on store instead of autograd running an add kernel."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, eps, rows2, codes, *wab):
        C = x.shape[-1]
        x2 = _rows(x, C)
        q, s, stats = _ln_quant(x2, ln_w, ln_b, eps)
        ys = tuple(_gemm(q, s, lc, rows2).view(*x.shape[:-1], lc.N) for lc in codes)
        ctx.save_for_backward(x2, stats, ln_w, q, s)
        ctx.codes, ctx.rows2, ctx.x_shape = codes, rows2, x.shape
        return (x,) + ys

    @staticmethod
    def backward(ctx, g_x, *g_ys):
        x2, stats, ln_w, q, s = ctx.saved_tensors
        _weight_epoch[0] += 1
        need = ctx.needs_input_grad
        grads, dys, qb = [], [], None
        for i, (lc, g) in enumerate(zip(ctx.codes, g_ys)):
            if g is None:
                grads += [None, None, None]
                continue
            nw, na, nb = need[6 + 3 * i: 9 + 3 * i]
            gx_i, gw, ga, gb, qb = _layer_backward(_rows(g, lc.N), None, q, s, qb, lc, ctx.rows2, True, nw or na,
                                                   nb and lc.bias is not None)
            dys.append(gx_i)
            grads += [gw, ga, gb]
        resid = None if g_x is None else _rows(g_x, x2.shape[1])
        if dys:
            gx, dgamma, dbeta = _ln_backward(dys, x2, stats, ln_w, resid)
            gx = gx.view(ctx.x_shape)
        else:
            gx, dgamma, dbeta = g_x, None, None
        return (gx, dgamma, dbeta, None, None, None) + tuple(grads)


def ln_proj_usable(x, layers) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and x.numel() > 0 and x.shape[-1] in LN_WIDTHS and 1 <= len(layers) <= 3
            and all(l.in_features == x.shape[-1] and l.out_features % 64 == 0 and l.in_features % 64 == 0 for l in layers)
            and all(not hasattr(l, "weight") or l.weight.dtype == torch.float32 for l in layers))


def ln_projections(x, ln, layers, rows2: int):
    """Returns ``(x_through, [proj(norm(x)) for proj in layers])``; use ``x_through`` for the module's residual."""
    M = x.numel() // x.shape[-1]
    codes = tuple(LayerCodes(l, rows2, M) for l in layers)
    wab = []
    for l in layers:
        wab += [l.weight if hasattr(l, "weight") else None, l.alpha, l.bias]
    out = _LnProjFn.apply(x, ln.weight, ln.bias, ln.eps, rows2, codes, *wab)
    return out[0], list(out[1:])


# ---------------------------------------------------------------------------------------------- projection + module tail
class _ProjTailFn(torch.autograd.Function):
    """resid + scale * dropout(proj(act)) * frame_mask: the attention's out_proj with the module tail in its epilogue
    (conformer.py:131-138)."""

    @staticmethod
    def forward(ctx, act, resid, w, a, b, lc, rows2, tail):
        a2 = _rows(act, lc.K)
        q, s = _act_quant(a2)
        r2 = _rows(resid, lc.N)
        out = _gemm(q, s, lc, rows2, tail, r2)
        ctx.save_for_backward(q, s)
        ctx.lc, ctx.rows2, ctx.tail, ctx.act_shape, ctx.out_shape = lc, rows2, tail, act.shape, resid.shape
        return out.view(resid.shape)

    @staticmethod
    def backward(ctx, g):
        q, s = ctx.saved_tensors
        _weight_epoch[0] += 1
        need = ctx.needs_input_grad
        lc = ctx.lc
        g2 = _rows(g, lc.N)
        g_act, gw, ga, gb, _ = _layer_backward(g2, ("tail", ctx.tail), q, s, None, lc, ctx.rows2, need[0], need[2] or need[3],
                                               need[4] and lc.bias is not None)
        return (None if g_act is None else g_act.view(ctx.act_shape)), (g if need[1] else None), gw, ga, gb, None, None, None


def proj_tail_usable(act, resid, layer) -> bool:
    return (act.is_cuda and act.dtype == torch.float32 and resid.dtype == torch.float32 and act.numel() > 0
            and layer.in_features == act.shape[-1] and layer.out_features == resid.shape[-1]
            and layer.in_features % 64 == 0 and layer.out_features % 64 == 0
            and (not hasattr(layer, "weight") or layer.weight.dtype == torch.float32))


def proj_tail(act, resid, layer, rows2: int, rowmask, p: float, training: bool, scale: float = 1.0):
    M = act.numel() // act.shape[-1]
    lc = LayerCodes(layer, rows2, M)
    tail = TailSpec(act.device, rowmask, scale, p, training)
    w = layer.weight if hasattr(layer, "weight") else None
    return _ProjTailFn.apply(act, resid, w, layer.alpha, layer.bias, lc, rows2, tail)
