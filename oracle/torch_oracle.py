"""Differentiable torch-CPU oracle layer.  TEST INFRASTRUCTURE ONLY (see onebit_oracle.py).

``OracleQuantizedLinear`` restates the reference layer (onebit_asr/quant.py:99-127) with
plain torch ops so that whole-model parity (loss / gradients over a short run) and the
CPU baseline can be computed where /root/reference is not mounted (the GPU box).
``act_bits=32`` is Oracle-A (exactly the reference's math); ``act_bits=8`` is Oracle-B
(per-token absmax int8 activations inserted in front, identity STE).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class _WeightSTE(torch.autograd.Function):
    """quant.py:38-92 restated: forward codes, clip-window STE, custom d/d-alpha."""

    @staticmethod
    def forward(ctx, W, a_eff, bitwidth):
        wa = W / a_eff
        wc = wa.clamp(-1.0, 1.0)
        sgn = torch.sign(wc)
        if bitwidth == 1:
            q = torch.where(sgn == 0, torch.ones_like(sgn), sgn)        # quant.py:52-55
        elif bitwidth == 2:
            q = torch.where(wc.abs() < 0.5, torch.zeros_like(sgn), sgn)  # quant.py:56-60
        else:
            raise ValueError("bitwidth must be one of {1,2,32}")
        ctx.bitwidth = bitwidth
        ctx.save_for_backward(wa)
        return a_eff * q                                                 # quant.py:68

    @staticmethod
    def backward(ctx, g):
        (wa,) = ctx.saved_tensors
        mag = wa.abs()
        sgn = torch.sign(wa)
        g_w = g * (mag <= 1.0).to(g.dtype)                               # quant.py:81-82
        proj = sgn * (mag >= 0.5).to(g.dtype) if ctx.bitwidth == 2 else sgn
        term = torch.where(mag < 1.0, proj - wa, sgn)                    # quant.py:86-90
        return g_w, (g * term).sum(), None                               # quant.py:91-92


def oracle_quantize_weight(W, a_eff, bitwidth):
    if bitwidth == 32:
        return W
    return _WeightSTE.apply(W, a_eff, bitwidth)


def oracle_act_quant(x):
    """Oracle-B activation quantiser: returns x_tilde with identity STE."""
    s = 127.0 / x.abs().amax(dim=-1, keepdim=True).clamp(min=1e-5)
    q = (x * s).round().clamp(-128, 127)
    return x + (q / s - x).detach()


class OracleQuantizedLinear(nn.Module):
    """Constructor / parameters / forward signature of the reference layer (quant.py:99-127)."""

    act_bits_default = 8

    def __init__(self, in_features: int, out_features: int, bias: bool = True, act_bits=None):
        super().__init__()
        w = torch.empty(out_features, in_features)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))                     # quant.py:104
        w.mul_(2.0)                                                      # quant.py:107-108
        self.weight = nn.Parameter(w)
        self.alpha = nn.Parameter(w.detach().abs().mean())               # quant.py:111-113
        self.bias = nn.Parameter(torch.zeros(out_features)) if bias else None
        self.act_bits = self.act_bits_default if act_bits is None else act_bits

    def forward(self, x, bitwidth: int):
        if bitwidth == 32:
            return F.linear(x, self.weight, self.bias)
        if bitwidth not in (1, 2):
            raise ValueError("bitwidth must be one of {1,2,32}")
        w_hat = oracle_quantize_weight(self.weight, self.alpha.abs() + 1e-8, bitwidth)
        if self.act_bits == 8:
            x = oracle_act_quant(x)
        return F.linear(x, w_hat, self.bias)
