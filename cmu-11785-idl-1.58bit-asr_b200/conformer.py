"""Conformer caller of the quantised layer: the module tree of the reference's ``onebit_asr/conformer.py``.

This is the "callers either side of the path" row of SURVEY.md section 8(f): it exists so that the training-step
workload of BASELINE.json (configs[2], configs[3]) can run where /root/reference is not mounted, and so that the
reference's checkpoints load (same attribute names => same state_dict keys, same construction order => the same
parameters from the same seed).  Only the nine projections per block that the reference routes through
``QuantizedLinear`` (conformer.py:31-32, 87-91) use the B200 layer; everything else is stock ``torch.nn`` exactly
as in the reference (fp32 Conv1d/Conv2d/LayerNorm/BatchNorm, explicit rel-pos attention), cited per class below.

``linear_cls`` selects the routed layer: the B200 ``QuantizedLinear`` by default; the tests and the CPU baseline
pass the oracle layer to obtain the reference's numbers from the same module tree.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import attention
from .norm import layer_norm
from .quant import QuantizedLinear


def swish(t: torch.Tensor) -> torch.Tensor:
    return t * torch.sigmoid(t)


class LayerNorm(nn.Module):
    """Wrapper that keeps the reference's parameter path ``<name>.ln.{weight,bias}`` (conformer.py:19-24)."""

    def __init__(self, width: int):
        super().__init__()
        self.ln = nn.LayerNorm(width)

    def forward(self, t):
        return layer_norm(t, self.ln.weight, self.ln.bias, self.ln.eps)


class FeedForwardModule(nn.Module):
    """Macaron half-step FFN: x + 0.5 * drop(lin2(drop(swish(lin1(LN x)))))   (conformer.py:27-45)."""

    def __init__(self, d_model: int, d_ff: int, dropout: float, linear_cls=QuantizedLinear):
        super().__init__()
        self.ln = LayerNorm(d_model)
        self.lin1 = linear_cls(d_model, d_ff)
        self.lin2 = linear_cls(d_ff, d_model)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, bitwidth: int, mask=None):
        h = self.lin1(self.ln(x), bitwidth)
        if hasattr(self.lin2, "forward_swish_dropout"):
            # B200 layer: swish + dropout + activation quantiser fused in front of lin2's GEMM (same math)
            h = self.lin2.forward_swish_dropout(h, bitwidth, self.dropout.p, self.training)
        else:
            h = self.lin2(self.dropout(swish(h)), bitwidth)
        h = self.dropout(h)
        if mask is not None:                                   # zero padded frames (conformer.py:42-44)
            h = h * mask[:, :, 0].unsqueeze(-1)
        return x + 0.5 * h


class RelPositionalEncoding(nn.Module):
    """Sinusoid table handed to the attention as ``pos_emb``; the input only sees dropout (conformer.py:48-76)."""

    def __init__(self, d_model: int, dropout_rate: float = 0.1, max_len: int = 5000):
        super().__init__()
        self.d_model = d_model
        self.dropout = nn.Dropout(p=dropout_rate)
        self.extend_pe(max_len)

    def extend_pe(self, length: int) -> None:
        have = getattr(self, "pe", None)
        if have is not None and have.size(1) >= length:
            return
        pos = torch.arange(0, length, dtype=torch.float).unsqueeze(1)
        inv = torch.exp(torch.arange(0, self.d_model, 2, dtype=torch.float) * -(math.log(10000.0) / self.d_model))
        table = torch.zeros(length, self.d_model)
        table[:, 0::2] = torch.sin(pos * inv)
        table[:, 1::2] = torch.cos(pos * inv)
        table = table.unsqueeze(0)
        if have is not None:
            self.pe = table.to(have.device)
        else:
            self.register_buffer("pe", table)

    def forward(self, x):
        self.extend_pe(x.size(1))
        return self.dropout(x), self.pe[:, : x.size(1)]


class MHSA(nn.Module):
    """Relative-position multi-head self-attention, Transformer-XL style bias terms (conformer.py:79-138).

    q/k/v share one input tensor, so the B200 layer quantises it once (activation-quantiser cache)."""

    def __init__(self, d_model: int, n_heads: int, dropout: float, linear_cls=QuantizedLinear):
        super().__init__()
        assert d_model % n_heads == 0
        self.d_model, self.n_heads, self.d_head = d_model, n_heads, d_model // n_heads
        self.ln = LayerNorm(d_model)
        self.q_proj = linear_cls(d_model, d_model)
        self.k_proj = linear_cls(d_model, d_model)
        self.v_proj = linear_cls(d_model, d_model)
        self.pos_proj = linear_cls(d_model, d_model)
        self.out_proj = linear_cls(d_model, d_model)
        self.dropout = nn.Dropout(dropout)
        self.pos_bias_u = nn.Parameter(torch.randn(self.n_heads, self.d_head) * 0.01)
        self.pos_bias_v = nn.Parameter(torch.randn(self.n_heads, self.d_head) * 0.01)

    @staticmethod
    def rel_shift(s):
        """Pad one column, reinterpret as [T2+1, T1], drop the first row (conformer.py:96-103)."""
        B, H, T1, T2 = s.shape
        s = F.pad(s, (1, 0)).view(B, H, T2 + 1, T1)
        return s[:, :, 1:].reshape(B, H, T1, T2)

    def _heads(self, t, batch):
        return t.view(batch, -1, self.n_heads, self.d_head).transpose(1, 2)

    def forward(self, x, mask, bitwidth: int, pos_emb: torch.Tensor):
        B, T, C = x.shape
        assert C == self.d_model, f"Expected {self.d_model}, got {C}"
        y = self.ln(x)
        q = self._heads(self.q_proj(y, bitwidth), B)
        k = self._heads(self.k_proj(y, bitwidth), B)
        v = self._heads(self.v_proj(y, bitwidth), B)
        p = self._heads(self.pos_proj(pos_emb, bitwidth), 1)
        ac = torch.matmul(q + self.pos_bias_u.view(1, self.n_heads, 1, self.d_head), k.transpose(-2, -1))
        bd_raw = torch.matmul(q + self.pos_bias_v.view(1, self.n_heads, 1, self.d_head), p.transpose(-2, -1))
        if attention.usable(ac, mask):
            # shift + scale + mask + softmax + nan_to_num + dropout in one kernel each way (same formulas)
            attn = attention.rel_attention_probs(ac, bd_raw, mask, 1.0 / math.sqrt(self.d_head), self.dropout.p,
                                                 self.training)
        else:
            scores = (ac + self.rel_shift(bd_raw)) / math.sqrt(self.d_head)
            if mask is not None:
                scores = scores.masked_fill(mask[:, None, :, :] == 0, float("-inf"))
            attn = torch.nan_to_num(torch.softmax(scores, dim=-1), nan=0.0)  # fully padded rows -> 0 (conformer.py:127)
            attn = self.dropout(attn)
        h = (attn @ v).transpose(1, 2).contiguous().view(B, T, C)
        h = self.dropout(self.out_proj(h, bitwidth))
        if mask is not None:
            h = h * mask[:, :, 0].unsqueeze(-1)
        return x + h


class ConvModule(nn.Module):
    """LN -> pointwise(2C) -> GLU -> depthwise k -> BatchNorm(batch stats) -> swish -> pointwise; kept full
    precision by the reference (conformer.py:141-167, "kept full-precision" at :225)."""

    def __init__(self, d_model: int, kernel_size: int = 31, dropout: float = 0.1):
        super().__init__()
        self.ln = LayerNorm(d_model)
        self.pw1 = nn.Conv1d(d_model, 2 * d_model, kernel_size=1)
        self.glu = nn.GLU(dim=1)
        self.dw = nn.Conv1d(d_model, d_model, kernel_size=kernel_size, padding=kernel_size // 2, groups=d_model)
        self.bn = nn.BatchNorm1d(d_model, track_running_stats=False)
        self.pw2 = nn.Conv1d(d_model, d_model, kernel_size=1)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, mask=None):
        h = self.ln(x).transpose(1, 2)
        h = self.dw(self.glu(self.pw1(h)))
        h = self.pw2(swish(self.bn(h)))
        h = self.dropout(h).transpose(1, 2)
        if mask is not None:
            h = h * mask[:, :, 0].unsqueeze(-1)
        return x + h


class Conv2dSubsampling(nn.Module):
    """Two 3x3 stride-2 convolutions + linear: [B,T,F] -> [B, ((T-1)//2-1)//2, d_model] (conformer.py:170-208)."""

    def __init__(self, idim: int, d_model: int):
        super().__init__()
        self.d_model = d_model
        self.conv = nn.Sequential(
            nn.Conv2d(1, d_model, kernel_size=3, stride=2), nn.ReLU(),
            nn.Conv2d(d_model, d_model, kernel_size=3, stride=2), nn.ReLU())
        out_freq = ((idim - 1) // 2 - 1) // 2
        if out_freq <= 0:
            raise ValueError(f"Input dim too small for Conv2dSubsampling: idim={idim}")
        self.out = nn.Linear(d_model * out_freq, d_model)

    def forward(self, x):
        h = self.conv(x.unsqueeze(1))
        B, C, T, Fq = h.shape
        return self.out(h.transpose(1, 2).contiguous().view(B, T, C * Fq))


class ConformerBlock(nn.Module):
    """ff1 -> mhsa -> conv -> ff2 -> LN (conformer.py:212-228); only ff*/mhsa see the bitwidth."""

    def __init__(self, d_model, d_ff, n_heads, conv_kernel, dropout, block_index, linear_cls=QuantizedLinear):
        super().__init__()
        self.block_index = block_index
        self.ff1 = FeedForwardModule(d_model, d_ff, dropout, linear_cls)
        self.mhsa = MHSA(d_model, n_heads, dropout, linear_cls)
        self.conv = ConvModule(d_model, kernel_size=conv_kernel, dropout=dropout)
        self.ff2 = FeedForwardModule(d_model, d_ff, dropout, linear_cls)
        self.ln = LayerNorm(d_model)

    def forward(self, x, src_mask, bitwidth_linear: int, pos_emb):
        x = self.ff1(x, bitwidth_linear)
        x = self.mhsa(x, src_mask, bitwidth_linear, pos_emb)
        x = self.conv(x)
        x = self.ff2(x, bitwidth_linear)
        return self.ln(x)


class ConformerEncoder(nn.Module):
    """Subsample, positional table, N blocks with a per-layer bitwidth (conformer.py:231-272).

    ``precision`` applies to every block unless ``sp_mask`` (stochastic precision) is given, in which case block
    i runs at 1 bit where sp_mask[i] == 1 and at 2 bits elsewhere (conformer.py:265-269)."""

    def __init__(self, input_dim, d_model, n_layers, n_heads, d_ff, conv_kernel, dropout, linear_cls=QuantizedLinear):
        super().__init__()
        self.subsample = Conv2dSubsampling(input_dim, d_model)
        self.pos_enc = RelPositionalEncoding(d_model, dropout)
        self.blocks = nn.ModuleList(
            [ConformerBlock(d_model, d_ff, n_heads, conv_kernel, dropout, i, linear_cls) for i in range(n_layers)])
        self.ln_out = LayerNorm(d_model)

    def frontend(self, feats):
        """Bitwidth-independent prefix (conv subsampling): identical for the three passes of a training step."""
        return self.subsample(feats)

    def forward(self, feats, feat_lens, precision: int, sp_mask: Optional[Sequence[int]] = None, frontend_out=None):
        x = self.frontend(feats) if frontend_out is None else frontend_out
        B, T = x.shape[:2]
        enc_lens = feat_lens // 4                                            # conformer.py:253
        x, pos_emb = self.pos_enc(x)
        key_mask = torch.arange(T, device=x.device)[None, :] < enc_lens[:, None]
        attn_mask = key_mask[:, :, None] & key_mask[:, None, :]
        for i, blk in enumerate(self.blocks):
            bw = precision if sp_mask is None else (1 if sp_mask[i] == 1 else 2)
            x = blk(x, attn_mask, bw if bw in (1, 2) else 32, pos_emb)
        return self.ln_out(x), key_mask


class TransformerDecoder(nn.Module):
    """Stock attention decoder (conformer.py:275-299); full precision."""

    def __init__(self, vocab_size, d_model, n_layers, n_heads, d_ff, dropout, pad_id):
        super().__init__()
        self.emb = nn.Embedding(vocab_size, d_model, padding_idx=pad_id)
        layer = nn.TransformerDecoderLayer(d_model=d_model, nhead=n_heads, dim_feedforward=d_ff, dropout=dropout,
                                           batch_first=True)
        self.dec = nn.TransformerDecoder(layer, num_layers=n_layers)
        self.ln = LayerNorm(d_model)
        self.out = nn.Linear(d_model, vocab_size)

    def forward(self, tgt_inp, memory, memory_mask, tgt_key_padding_mask):
        n = tgt_inp.size(1)
        future = torch.triu(torch.ones(n, n, device=tgt_inp.device), diagonal=1).bool()
        causal = torch.zeros(n, n, device=tgt_inp.device).masked_fill(future, float("-inf"))
        h = self.dec(self.emb(tgt_inp), memory, tgt_mask=causal, memory_key_padding_mask=(memory_mask == 0),
                     tgt_key_padding_mask=tgt_key_padding_mask)
        return self.out(self.ln(h))


class ConformerASR(nn.Module):
    """Encoder + CTC head + attention decoder with the reference's constructor defaults (conformer.py:302-322)."""

    def __init__(self, input_dim: int, vocab_size: int, enc_d_model=256, enc_layers=12, enc_heads=4, enc_d_ff=1024,
                 enc_conv_kernel=31, enc_dropout=0.1, dec_layers=2, dec_heads=4, dec_d_ff=1024, dec_dropout=0.1,
                 pad_id=0, linear_cls=QuantizedLinear):
        super().__init__()
        self.encoder = ConformerEncoder(input_dim, enc_d_model, enc_layers, enc_heads, enc_d_ff, enc_conv_kernel,
                                        enc_dropout, linear_cls)
        self.decoder = TransformerDecoder(vocab_size, enc_d_model, dec_layers, dec_heads, dec_d_ff, dec_dropout, pad_id)
        self.ctc_head = nn.Linear(enc_d_model, vocab_size)

    def forward(self, batch, precision: int, sp_mask: Optional[List[int]] = None, frontend_out=None):
        enc_out, enc_mask = self.encoder(batch["feats"], batch["feat_lens"], precision, sp_mask, frontend_out)
        return enc_out, enc_mask, self.ctc_head(enc_out)

    def decode_logits(self, enc_out, enc_mask, tgt_inp, tgt_pad_mask):
        return self.decoder(tgt_inp, enc_out, enc_mask, tgt_pad_mask)

    def quantized_layers(self):
        return [m for m in self.modules() if isinstance(m, QuantizedLinear)]
