"""The Conformer ASR model that calls the quantised layer (the "callers" row of SURVEY.md section 8f).

Why this file exists: BASELINE.json's training and inference workloads are defined on the reference's Conformer
(``onebit_asr/conformer.py``), and the reference tree is not mounted where the benchmarks run.  The module tree
below reproduces that model's *interface*: identical attribute names (hence identical ``state_dict`` keys - the
reference's checkpoints load), identical construction order (hence identical parameters from the same seed;
``tests/test_conformer_cpu.py`` compares against the reference's own state_dict and outputs), identical arithmetic.
Only the nine projections per block that the reference routes through its quantised linear (conformer.py:31-32,
87-91) are B200-specific; the rest is stock fp32 ``torch.nn`` exactly as the reference has it.

What is organised differently, without changing the math:

* ``ModelDims`` carries the hyper-parameters (defaults of train.py:194-203);
* three element-wise chains run as single library kernels when the tensors are fp32 on a CUDA device - the FFN
  mid-section (swish, dropout, activation quantiser of lin2), LayerNorm, and the attention's shift / scale / mask /
  softmax / nan_to_num / dropout sequence - and fall back to the plain torch ops otherwise (CPU oracle runs);
* the convolutional front-end can be evaluated once and handed to several encoder passes (``frontend_out``).

``routed=`` selects the class used for the routed projections: the B200 ``QuantizedLinear`` by default, the CPU
oracle layer in the tests and in the CPU baseline.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import attention, convmod, frontend, fused, matmul, residual, routes
from .norm import layer_norm
from .quant import QuantizedLinear


@dataclass
class ModelDims:
    """Hyper-parameters with the reference's defaults (train.py:194-203)."""
    width: int = 256
    blocks: int = 12
    heads: int = 4
    ffn: int = 1024
    conv_taps: int = 31
    p_drop: float = 0.1


@dataclass(frozen=True)
class StackedBits:
    """Bitwidths of several passes evaluated side by side on one stacked batch (the three co-training passes of
    train.py:83-103 share every weight and differ only in the bitwidth): the first ``utt2`` utterances of the stack go
    through a routed layer at 2 bits, the others at 1 bit.  ``groups`` = number of stacked passes (BatchNorm statistics stay
    per pass)."""
    utt2: int
    groups: int


def swish(t: torch.Tensor) -> torch.Tensor:
    return t * torch.sigmoid(t)


def _routed_one(layer, x, bitwidth: int, swish_dropout):
    if swish_dropout is None:
        return layer(x, bitwidth)
    dropout, training = swish_dropout
    fused = getattr(layer, "forward_swish_dropout", None)
    if fused is not None:           # B200 layer: activation, dropout and the layer's quantiser in one kernel, then the GEMM
        return fused(x, bitwidth, dropout.p, training)
    return layer(dropout(swish(x)), bitwidth)


def _routed(layer, x, bits, swish_dropout=None):
    """``layer(x, bits)`` for an int bitwidth or a ``StackedBits``; ``swish_dropout = (nn.Dropout, training)`` puts the FFN
    mid-section ``dropout(swish(.))`` (conformer.py:37-38) in front of the layer."""
    if not isinstance(bits, StackedBits):
        return _routed_one(layer, x, bits, swish_dropout)
    utt2, total = bits.utt2, x.shape[0]
    grouped = getattr(layer, "forward_grouped", None)
    if grouped is not None and layer.grouped_usable(x, swish=swish_dropout is not None):
        sd = None if swish_dropout is None else (swish_dropout[0].p, swish_dropout[1])
        return grouped(x, utt2 * (x.numel() // (total * x.shape[-1])), sd)
    parts = []
    if utt2 > 0:
        parts.append(_routed_one(layer, x[:utt2], 2, swish_dropout))
    if utt2 < total:
        parts.append(_routed_one(layer, x[utt2:], 1, swish_dropout))
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)


def _rows2(bits, x) -> Optional[int]:
    """Number of leading token rows of ``x [B, T, C]`` that use the 2-bit codes (the rest use the 1-bit codes), or None when the
    call is not a 1-/2-bit one (full precision)."""
    rows = x.numel() // x.shape[-1]
    if isinstance(bits, StackedBits):
        return bits.utt2 * (rows // x.shape[0])
    return rows if bits == 2 else (0 if bits == 1 else None)


def _b200_layers(*layers) -> bool:
    return all(hasattr(l, "packed_weight") or hasattr(l, "packed") for l in layers) and "fuse" not in matmul.DISABLED


def _row_mask(mask):
    return None if mask is None else residual._RowMaskCache.get(_frame_mask(mask))


def _frame_mask(mask: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """[B,T,T] attention mask -> [B,T,1] validity of each query frame (conformer.py:42-44, 134-137)."""
    return None if mask is None else mask[:, :, 0].unsqueeze(-1)


def _module_tail(x, y, mask, scale: float, dropout: nn.Dropout, training: bool):
    """``x + scale * dropout(y) * frame_validity`` - the tail of every encoder module (conformer.py:41-45, 133-138, 163-167);
    one B200 kernel each way when the tensors qualify, the reference's op sequence otherwise."""
    if routes.taken("module_tail", residual.usable(x), x, "tail" in matmul.DISABLED):
        return residual.residual_dropout(x, y, _frame_mask(mask), scale, dropout.p, training)
    y = dropout(y)
    keep = _frame_mask(mask)
    if keep is not None:
        y = y * keep
    return x + y if scale == 1.0 else x + scale * y


# ----------------------------------------------------------------------------------------------------------------
# building blocks that keep the reference's parameter paths
# ----------------------------------------------------------------------------------------------------------------
class LayerNorm(nn.Module):
    """Holds ``ln = nn.LayerNorm(width)`` so that parameters are named ``<name>.ln.weight/bias`` as in the
    reference (conformer.py:19-24); the normalisation itself goes through ``norm.layer_norm``."""

    def __init__(self, width: int):
        super().__init__()
        self.ln = nn.LayerNorm(width)

    def forward(self, t):
        return layer_norm(t, self.ln.weight, self.ln.bias, self.ln.eps)


class RelPositionalEncoding(nn.Module):
    """Sinusoid table ``pe`` (buffer) sliced to the sequence length and returned beside the (dropped-out) input;
    the table grows on demand (conformer.py:48-76)."""

    def __init__(self, d_model: int, dropout_rate: float = 0.1, max_len: int = 5000):
        super().__init__()
        self.d_model = d_model
        self.dropout = nn.Dropout(p=dropout_rate)
        self.extend_pe(max_len)

    @staticmethod
    def _table(length: int, width: int) -> torch.Tensor:
        steps = torch.arange(0, length, dtype=torch.float).unsqueeze(1)
        rates = torch.exp(torch.arange(0, width, 2, dtype=torch.float) * -(math.log(10000.0) / width))
        out = torch.zeros(length, width)
        out[:, 0::2], out[:, 1::2] = torch.sin(steps * rates), torch.cos(steps * rates)
        return out.unsqueeze(0)

    def extend_pe(self, length: int) -> None:
        current = getattr(self, "pe", None)
        if current is None:
            self.register_buffer("pe", self._table(length, self.d_model))
        elif current.size(1) < length:
            self.pe = self._table(length, self.d_model).to(current.device)

    def forward(self, x):
        n = x.size(1)
        self.extend_pe(n)
        return self.dropout(x), self.pe[:, :n]


class FeedForwardModule(nn.Module):
    """Half-step feed-forward of the macaron pair (conformer.py:27-45):
    ``x + 0.5 * drop(lin2(drop(swish(lin1(norm(x))))))`` with lin1 / lin2 routed."""

    def __init__(self, d_model: int, d_ff: int, dropout: float, routed=QuantizedLinear):
        super().__init__()
        self.ln = LayerNorm(d_model)
        self.lin1 = routed(d_model, d_ff)
        self.lin2 = routed(d_ff, d_model)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, bitwidth: int, mask=None):
        rows2 = _rows2(bitwidth, x)
        if rows2 is not None and _b200_layers(self.lin1, self.lin2) and fused.ffn_usable(x, self.lin1, self.lin2):
            # the whole module as one fused chain: LayerNorm inside lin1's quantiser, the tail inside lin2's GEMM epilogue
            routes.taken("ffn_module_fused", True, x)
            return fused.ffn_forward(x, self.ln.ln, self.lin1, self.lin2, rows2, _row_mask(mask), self.dropout.p, self.training)
        hidden = _routed(self.lin1, self.ln(x), bitwidth)
        out = _routed(self.lin2, hidden, bitwidth, swish_dropout=(self.dropout, self.training))
        return _module_tail(x, out, mask, 0.5, self.dropout, self.training)


class MHSA(nn.Module):
    """Self-attention with Transformer-XL style relative positions (conformer.py:79-138).  All five projections are
    routed; q, k and v read the same normalised tensor, which the B200 layer therefore quantises once."""

    def __init__(self, d_model: int, n_heads: int, dropout: float, routed=QuantizedLinear):
        super().__init__()
        if d_model % n_heads:
            raise AssertionError("d_model must be divisible by n_heads")
        self.d_model, self.n_heads, self.d_head = d_model, n_heads, d_model // n_heads
        self.ln = LayerNorm(d_model)
        self.q_proj = routed(d_model, d_model)
        self.k_proj = routed(d_model, d_model)
        self.v_proj = routed(d_model, d_model)
        self.pos_proj = routed(d_model, d_model)
        self.out_proj = routed(d_model, d_model)
        self.dropout = nn.Dropout(dropout)
        self.pos_bias_u = nn.Parameter(torch.randn(self.n_heads, self.d_head) * 0.01)
        self.pos_bias_v = nn.Parameter(torch.randn(self.n_heads, self.d_head) * 0.01)

    @staticmethod
    def rel_shift(scores):
        """The reference's relative shift (conformer.py:96-103): prepend a zero column, reinterpret the padded
        [T1, T2+1] block as [T2+1, T1], drop its first row, read the rest as [T1, T2]."""
        b, h, t1, t2 = scores.shape
        padded = F.pad(scores, (1, 0)).view(b, h, t2 + 1, t1)
        return padded[:, :, 1:].reshape(b, h, t1, t2)

    def _split(self, t, batch):
        return t.view(batch, -1, self.n_heads, self.d_head).transpose(1, 2)      # [batch, heads, T, d_head]

    def _probabilities(self, content, position, mask):
        """``content`` = (q+u) k^T, ``position`` = (q+v) p^T before the shift -> attention weights after dropout."""
        inv_sqrt = 1.0 / math.sqrt(self.d_head)
        if routes.taken("attention_chain", attention.usable(content, mask), content):
            return attention.rel_attention_probs(content, position, mask, inv_sqrt, self.dropout.p, self.training)
        logits = (content + self.rel_shift(position)) / math.sqrt(self.d_head)
        if mask is not None:
            logits = logits.masked_fill(mask[:, None, :, :] == 0, float("-inf"))
        weights = torch.nan_to_num(torch.softmax(logits, dim=-1), nan=0.0)        # all-padding rows become zeros
        return self.dropout(weights)

    def forward(self, x, mask, bitwidth: int, pos_emb: torch.Tensor):
        batch, frames, width = x.shape
        if width != self.d_model:
            raise AssertionError(f"Expected {self.d_model}, got {width}")
        rows2 = _rows2(bitwidth, x)
        qkv = (self.q_proj, self.k_proj, self.v_proj)
        on_library = attention.rel_attention_usable(x, mask, self.n_heads) and pos_emb.shape[:2] == (1, frames)
        if (rows2 is not None and on_library and _b200_layers(*qkv, self.out_proj) and fused.ln_proj_usable(x, qkv)
                and fused.proj_tail_usable(x, x, self.out_proj)):
            # LayerNorm inside the shared q/k/v quantiser, the tail inside out_proj's GEMM epilogue
            routes.taken("attention", True, x)
            routes.taken("mhsa_module_fused", True, x)
            x_res, (q_flat, k_flat, v_flat) = fused.ln_projections(x, self.ln.ln, qkv, rows2)
            if isinstance(bitwidth, StackedBits) and 0 < bitwidth.utt2 < batch:
                # two bitwidth groups: the positional table is projected once per bitwidth and used by its group, broadcast
                pos_flat, pos_b, split = self.pos_proj(pos_emb, 2), self.pos_proj(pos_emb, 1), bitwidth.utt2
            else:
                pos_flat, pos_b, split = self._positions(pos_emb, bitwidth, batch), None, 0
            mixed = attention.rel_attention(q_flat, k_flat, v_flat, pos_flat, self.pos_bias_u, self.pos_bias_v, mask,
                                            self.n_heads, self.dropout.p, self.training, pos_b=pos_b, split=split)
            return fused.proj_tail(mixed, x_res, self.out_proj, rows2, _row_mask(mask), self.dropout.p, self.training)
        normed = self.ln(x)
        q_flat, k_flat, v_flat = (_routed(p, normed, bitwidth) for p in qkv)
        pos_flat = self._positions(pos_emb, bitwidth, batch)
        if routes.taken("attention", on_library, normed, "attn" in matmul.DISABLED):
            # tensor-core path: the projections are consumed in their [B, T, H*d] layout, no head transposes
            mixed = attention.rel_attention(q_flat, k_flat, v_flat, pos_flat, self.pos_bias_u, self.pos_bias_v, mask,
                                            self.n_heads, self.dropout.p, self.training)
            return self._finish(x, mixed, mask, bitwidth)
        q, k, v = self._split(q_flat, batch), self._split(k_flat, batch), self._split(v_flat, batch)
        pos = self._split(pos_flat, pos_flat.shape[0])
        u = self.pos_bias_u.view(1, self.n_heads, 1, self.d_head)
        w = self.pos_bias_v.view(1, self.n_heads, 1, self.d_head)
        probs = self._probabilities(torch.matmul(q + u, k.transpose(-2, -1)), torch.matmul(q + w, pos.transpose(-2, -1)), mask)
        mixed = (probs @ v).transpose(1, 2).contiguous().view(batch, frames, width)
        return self._finish(x, mixed, mask, bitwidth)

    def _positions(self, pos_emb, bitwidth, batch):
        """Projected positional table: ``[1, T, W]``, or per utterance ``[batch, T, W]`` when stacked passes use two bitwidths."""
        if not isinstance(bitwidth, StackedBits):
            return self.pos_proj(pos_emb, bitwidth)
        if bitwidth.utt2 >= batch:
            return self.pos_proj(pos_emb, 2)
        if bitwidth.utt2 <= 0:
            return self.pos_proj(pos_emb, 1)
        two, one = self.pos_proj(pos_emb, 2), self.pos_proj(pos_emb, 1)
        return torch.cat([two.expand(bitwidth.utt2, -1, -1), one.expand(batch - bitwidth.utt2, -1, -1)], dim=0)

    def _finish(self, x, mixed, mask, bitwidth):
        return _module_tail(x, _routed(self.out_proj, mixed, bitwidth), mask, 1.0, self.dropout, self.training)


class ConvModule(nn.Module):
    """Convolution module, full precision in the reference (conformer.py:141-167 and the comment at :225):
    norm, 1x1 conv to 2C, GLU, depthwise conv, BatchNorm with batch statistics, swish, 1x1 conv, dropout."""

    def __init__(self, d_model: int, kernel_size: int = 31, dropout: float = 0.1):
        super().__init__()
        self.ln = LayerNorm(d_model)
        self.pw1 = nn.Conv1d(d_model, 2 * d_model, 1)
        self.glu = nn.GLU(dim=1)
        self.dw = nn.Conv1d(d_model, d_model, kernel_size, padding=kernel_size // 2, groups=d_model)
        self.bn = nn.BatchNorm1d(d_model, track_running_stats=False)
        self.pw2 = nn.Conv1d(d_model, d_model, 1)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, mask=None, groups: int = 1):
        """``groups`` > 1: x stacks that many passes; BatchNorm uses each pass's own batch statistics."""
        taps = self.dw.kernel_size[0]
        on_library = convmod.usable(x, x.shape[-1], taps) and self.bn.affine and self.dw.padding[0] == taps // 2
        if routes.taken("conv_module", on_library, x, "conv" in matmul.DISABLED):
            # channel-last path: 1x1 convolutions as matrix products over the channel axis, B200 kernels in between
            a = matmul.linear(self.ln(x), self.pw1.weight.squeeze(-1), self.pw1.bias)
            s = convmod.glu_dwconv_bn_swish(a, self.dw.weight, self.dw.bias, self.bn.weight, self.bn.bias, self.bn.eps, groups)
            t = matmul.linear(s, self.pw2.weight.squeeze(-1), self.pw2.bias)
        else:
            t = self.ln(x).transpose(1, 2)                # channels first for torch's convolutions
            t = self.dw(self.glu(self.pw1(t)))
            t = self.bn(t) if groups == 1 else torch.cat([self.bn(c) for c in t.chunk(groups, dim=0)], dim=0)
            t = self.pw2(swish(t)).transpose(1, 2)
        return _module_tail(x, t, mask, 1.0, self.dropout, self.training)


class Conv2dSubsampling(nn.Module):
    """Front-end: two stride-2 3x3 convolutions with ReLU, then a linear layer over (channels x remaining mel bins);
    time shrinks to ((T-1)//2 - 1)//2 (conformer.py:170-208)."""

    def __init__(self, idim: int, d_model: int):
        super().__init__()
        self.d_model = d_model
        self.conv = nn.Sequential(nn.Conv2d(1, d_model, 3, 2), nn.ReLU(), nn.Conv2d(d_model, d_model, 3, 2), nn.ReLU())
        bins = ((idim - 1) // 2 - 1) // 2
        if bins <= 0:
            raise ValueError(f"Input dim too small for Conv2dSubsampling: idim={idim}")
        self.out = nn.Linear(d_model * bins, d_model)

    def forward(self, feats):
        if routes.taken("frontend_conv1", frontend.usable(feats, self.conv[0]), feats, "frontend" in matmul.DISABLED):
            # first convolution + bias + ReLU in one write-bound pass, channels-last for cuDNN's second convolution
            maps = self.conv[3](self.conv[2](frontend.conv1_relu(feats, self.conv[0].weight, self.conv[0].bias)))
        else:
            maps = self.conv(feats.unsqueeze(1))          # [B, C, T', F']
        b, c, t, f = maps.shape
        return matmul.linear(maps.transpose(1, 2).contiguous().view(b, t, c * f), self.out.weight, self.out.bias)


class ConformerBlock(nn.Module):
    """ff1, self-attention, convolution, ff2, final norm (conformer.py:212-228); the bitwidth reaches ff1, mhsa, ff2."""

    def __init__(self, d_model, d_ff, n_heads, conv_kernel, dropout, block_index, routed=QuantizedLinear):
        super().__init__()
        self.block_index = block_index
        self.ff1 = FeedForwardModule(d_model, d_ff, dropout, routed)
        self.mhsa = MHSA(d_model, n_heads, dropout, routed)
        self.conv = ConvModule(d_model, kernel_size=conv_kernel, dropout=dropout)
        self.ff2 = FeedForwardModule(d_model, d_ff, dropout, routed)
        self.ln = LayerNorm(d_model)

    def forward(self, x, src_mask, bitwidth_linear: int, pos_emb):
        groups = bitwidth_linear.groups if isinstance(bitwidth_linear, StackedBits) else 1
        x = self.mhsa(self.ff1(x, bitwidth_linear), src_mask, bitwidth_linear, pos_emb)
        return self.ln(self.ff2(self.conv(x, groups=groups), bitwidth_linear))


class ConformerEncoder(nn.Module):
    """Front-end, positional table and the stack of blocks (conformer.py:231-272).  Per-block bitwidth: ``precision``
    for all of them, or - under stochastic precision - 1 bit where ``sp_mask[i] == 1`` and 2 bits elsewhere
    (conformer.py:265-269); anything but 1 or 2 means full precision (32)."""

    def __init__(self, input_dim, d_model, n_layers, n_heads, d_ff, conv_kernel, dropout, routed=QuantizedLinear):
        super().__init__()
        self.subsample = Conv2dSubsampling(input_dim, d_model)
        self.pos_enc = RelPositionalEncoding(d_model, dropout)
        self.blocks = nn.ModuleList(ConformerBlock(d_model, d_ff, n_heads, conv_kernel, dropout, i, routed)
                                    for i in range(n_layers))
        self.ln_out = LayerNorm(d_model)

    def frontend(self, feats):
        """The part that does not depend on the bitwidth and has no dropout: shareable between passes."""
        return self.subsample(feats)

    @staticmethod
    def _bitwidth(i: int, precision: int, sp_mask) -> int:
        bits = precision if sp_mask is None else (1 if sp_mask[i] == 1 else 2)
        return bits if bits in (1, 2) else 32

    def forward(self, feats, feat_lens, precision: int, sp_mask: Optional[Sequence[int]] = None, frontend_out=None):
        x = frontend_out if frontend_out is not None else self.frontend(feats)
        frames = x.size(1)
        valid = torch.arange(frames, device=x.device)[None, :] < (feat_lens // 4)[:, None]      # conformer.py:253-260
        pair_mask = valid[:, :, None] & valid[:, None, :]
        x, pos_emb = self.pos_enc(x)
        for i, block in enumerate(self.blocks):
            x = block(x, pair_mask, self._bitwidth(i, precision, sp_mask), pos_emb)
        return self.ln_out(x), valid


    @staticmethod
    def stack_plan(passes):
        """Order in which ``passes`` = [(precision, sp_mask), ...] can be stacked so that the 2-bit utterances of every block are
        a prefix of the stack: all-2-bit passes, then the (single) stochastic-precision pass, then all-1-bit passes.  Returns
        None when the combination cannot be stacked (full precision passes, several stochastic ones)."""
        two = [i for i, (prec, m) in enumerate(passes) if m is None and prec == 2]
        one = [i for i, (prec, m) in enumerate(passes) if m is None and prec == 1]
        sp = [i for i, (_, m) in enumerate(passes) if m is not None]
        if len(two) + len(one) + len(sp) != len(passes) or len(sp) > 1 or len(passes) < 2:
            return None
        return two + sp + one, len(two), (passes[sp[0]][1] if sp else None)

    def forward_stacked(self, feats, feat_lens, passes, frontend_out=None):
        """Several passes over the same batch side by side (see ``StackedBits``): returns [(memory, valid), ...] in the order of
        ``passes``, equal to ``[self(feats, feat_lens, prec, mask) for prec, mask in passes]`` up to summation order."""
        plan = self.stack_plan(passes)
        if plan is None:
            return [self(feats, feat_lens, prec, m, frontend_out) for prec, m in passes]
        order, n_two, sp_mask = plan
        x = frontend_out if frontend_out is not None else self.frontend(feats)
        n_pass, batch, frames = len(passes), x.size(0), x.size(1)
        valid = torch.arange(frames, device=x.device)[None, :] < (feat_lens // 4)[:, None]
        pair_mask = (valid[:, :, None] & valid[:, None, :]).repeat(n_pass, 1, 1)
        x, pos_emb = self.pos_enc(x.repeat(n_pass, 1, 1))
        for i, block in enumerate(self.blocks):
            sp_two = sp_mask is not None and sp_mask[i] != 1
            x = block(x, pair_mask, StackedBits((n_two + int(sp_two)) * batch, n_pass), pos_emb)
        y = self.ln_out(x)
        outs = [None] * n_pass
        for slot, i in enumerate(order):
            outs[i] = (y[slot * batch:(slot + 1) * batch], valid)
        return outs


class TransformerDecoder(nn.Module):
    """The attention decoder is PyTorch's own ``nn.TransformerDecoder`` (conformer.py:275-299)."""

    def __init__(self, vocab_size, d_model, n_layers, n_heads, d_ff, dropout, pad_id):
        super().__init__()
        self.emb = nn.Embedding(vocab_size, d_model, padding_idx=pad_id)
        self.dec = nn.TransformerDecoder(
            nn.TransformerDecoderLayer(d_model=d_model, nhead=n_heads, dim_feedforward=d_ff, dropout=dropout,
                                       batch_first=True), num_layers=n_layers)
        self.ln = LayerNorm(d_model)
        self.out = nn.Linear(d_model, vocab_size)

    def forward(self, tgt_inp, memory, memory_mask, tgt_key_padding_mask):
        steps = tgt_inp.size(1)
        ahead = torch.ones(steps, steps, device=tgt_inp.device).triu(1).bool()
        causal = ahead.float().masked_fill(ahead, float("-inf"))
        on_library = (matmul.linear_usable(memory, self.out.weight) and "decoder" not in matmul.DISABLED
                      and not self.dec.layers[0].norm_first)
        if routes.taken("decoder_layers", on_library, memory, bool({"decoder", "linear"} & matmul.DISABLED)):
            hidden = self._layers_on_library(self.emb(tgt_inp), memory, causal, memory_mask == 0, tgt_key_padding_mask)
        else:
            hidden = self.dec(self.emb(tgt_inp), memory, tgt_mask=causal, memory_key_padding_mask=(memory_mask == 0),
                              tgt_key_padding_mask=tgt_key_padding_mask)
        return matmul.linear(self.ln(hidden), self.out.weight, self.out.bias)

    # The same computation as nn.TransformerDecoder(post-norm layers, batch_first) with its parameters, written out so that
    # every projection goes through ``matmul.linear`` (fp32-accurate tensor-core GEMM) - in particular the key/value
    # projection of the encoder memory, the one large matmul of the decoder.  Attention itself stays torch's fused SDPA.
    @staticmethod
    def _attend(mha: nn.MultiheadAttention, query, source, additive_mask, key_padding, training: bool, self_attention: bool):
        width, heads = mha.embed_dim, mha.num_heads
        w, b = mha.in_proj_weight, mha.in_proj_bias
        if self_attention:
            q, k, v = matmul.linear(query, w, b).split(width, dim=-1)
        else:
            q = matmul.linear(query, w[:width], None if b is None else b[:width])
            k, v = matmul.linear(source, w[width:], None if b is None else b[width:]).split(width, dim=-1)
        batch, q_len, k_len = query.shape[0], query.shape[1], source.shape[1]

        def heads_first(t, length):
            return t.reshape(batch, length, heads, width // heads).transpose(1, 2)

        mask = None
        if key_padding is not None:
            mask = torch.zeros(batch, 1, 1, k_len, device=query.device, dtype=query.dtype)
            mask = mask.masked_fill(key_padding[:, None, None, :], float("-inf"))
        if additive_mask is not None:
            mask = additive_mask[None, None] if mask is None else mask + additive_mask[None, None]
        if mask is not None:
            mask = mask.expand(batch, heads, q_len, k_len)
        out = F.scaled_dot_product_attention(heads_first(q, q_len), heads_first(k, k_len), heads_first(v, k_len), attn_mask=mask,
                                             dropout_p=mha.dropout if training else 0.0)
        out = out.transpose(1, 2).reshape(batch, q_len, width)
        return matmul.linear(out, mha.out_proj.weight, mha.out_proj.bias)

    def _layers_on_library(self, x, memory, causal, memory_padding, tgt_padding):
        for layer in self.dec.layers:
            attn = self._attend(layer.self_attn, x, x, causal, tgt_padding, self.training, True)
            x = layer.norm1(x + layer.dropout1(attn))
            attn = self._attend(layer.multihead_attn, x, memory, None, memory_padding, self.training, False)
            x = layer.norm2(x + layer.dropout2(attn))
            ff = matmul.linear(layer.dropout(layer.activation(matmul.linear(x, layer.linear1.weight, layer.linear1.bias))),
                               layer.linear2.weight, layer.linear2.bias)
            x = layer.norm3(x + layer.dropout3(ff))
        return x if self.dec.norm is None else self.dec.norm(x)


class ConformerASR(nn.Module):
    """Encoder, CTC head and attention decoder with the reference's keyword arguments (conformer.py:302-322)."""

    def __init__(self, input_dim: int, vocab_size: int, enc_d_model=256, enc_layers=12, enc_heads=4, enc_d_ff=1024,
                 enc_conv_kernel=31, enc_dropout=0.1, dec_layers=2, dec_heads=4, dec_d_ff=1024, dec_dropout=0.1,
                 pad_id=0, linear_cls=QuantizedLinear):
        super().__init__()
        self.dims = ModelDims(enc_d_model, enc_layers, enc_heads, enc_d_ff, enc_conv_kernel, enc_dropout)
        self.encoder = ConformerEncoder(input_dim, enc_d_model, enc_layers, enc_heads, enc_d_ff, enc_conv_kernel,
                                        enc_dropout, routed=linear_cls)
        self.decoder = TransformerDecoder(vocab_size, enc_d_model, dec_layers, dec_heads, dec_d_ff, dec_dropout, pad_id)
        self.ctc_head = nn.Linear(enc_d_model, vocab_size)

    def forward(self, batch, precision: int, sp_mask: Optional[List[int]] = None, frontend_out=None):
        """batch: dict with ``feats [B,T,F]`` and ``feat_lens [B]`` -> (encoder output, frame validity, CTC logits)."""
        memory, valid = self.encoder(batch["feats"], batch["feat_lens"], precision, sp_mask, frontend_out)
        return memory, valid, matmul.linear(memory, self.ctc_head.weight, self.ctc_head.bias)

    def forward_passes(self, batch, passes, frontend_out=None):
        """``[self(batch, prec, mask, frontend_out) for prec, mask in passes]`` with the encoder passes stacked side by side."""
        outs = self.encoder.forward_stacked(batch["feats"], batch["feat_lens"], passes, frontend_out)
        return [(mem, valid, matmul.linear(mem, self.ctc_head.weight, self.ctc_head.bias)) for mem, valid in outs]

    def decode_logits(self, enc_out, enc_mask, tgt_inp, tgt_pad_mask):
        return self.decoder(tgt_inp, enc_out, enc_mask, tgt_pad_mask)

    def quantized_layers(self):
        return [m for m in self.modules() if isinstance(m, QuantizedLinear)]

    def use_packed_code_arena(self):
        """Re-quantise all routed layers (both bitwidths) with ONE launch per optimiser step instead of two per layer; call after
        the model is on its device.  Returns the arena (``.repacks`` counts the launches)."""
        from .quant import PackedCodeArena
        return PackedCodeArena(self.quantized_layers())
