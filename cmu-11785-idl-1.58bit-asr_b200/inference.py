"""Inference side of the quantised path (BASELINE configs[4]).

* ``PackedQuantizedLinear`` - a frozen layer that holds ONLY the 2-bit packed codes, alpha and bias (16x smaller
  than the fp32 latent weight) and runs act-quant + the tcgen05 GEMM; same ``forward(x, bitwidth)`` signature.
* ``pack_model_for_inference(model, bitwidth)`` swaps every routed projection of a Conformer for its packed form;
  ``packed_state_dict`` / ``load_packed_state_dict`` are the deployment format that sits beside the reference's
  ``ckpt_last.pt`` / ``best.pt`` (train.py:307-317) without changing it.
* ``ctc_greedy_decode`` - the reference's greedy CTC decoder (metrics.py:51-60) on the device for a whole batch.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn as nn

from ._cabi import OB_ALPHA_RAW, check, lib
from .quant import _DTYPE_TAG, QuantizedLinear, _ActQuantCache, _require_cuda, _stream, _tag


class PackedQuantizedLinear(nn.Module):
    """Inference-only form of ``QuantizedLinear`` at one fixed bitwidth: buffers ``packed [N, K/4] uint8``
    (OB_ORDER_I8), ``alpha []``, ``bias [N]``.  No gradient, no latent weight."""

    def __init__(self, in_features: int, out_features: int, bitwidth: int, bias: bool = True):
        super().__init__()
        if bitwidth not in (1, 2):
            raise ValueError("bitwidth must be one of {1,2}")
        self.in_features, self.out_features, self.bitwidth = in_features, out_features, bitwidth
        self.register_buffer("packed", torch.zeros(out_features, in_features // 4, dtype=torch.uint8))
        self.register_buffer("alpha", torch.zeros(()))
        self.register_buffer("bias", torch.zeros(out_features) if bias else None)

    @classmethod
    def from_layer(cls, layer: QuantizedLinear, bitwidth: int) -> "PackedQuantizedLinear":
        m = cls(layer.in_features, layer.out_features, bitwidth, layer.bias is not None)
        m = m.to(layer.weight.device)
        packed, _ = layer.packed_weight(bitwidth)
        m.packed.copy_(packed)
        m.alpha.copy_(layer.alpha.detach())
        if layer.bias is not None:
            m.bias.copy_(layer.bias.detach())
        return m

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, bitwidth={self.bitwidth}"

    @torch.no_grad()
    def forward(self, x: torch.Tensor, bitwidth: int) -> torch.Tensor:
        if bitwidth != self.bitwidth:
            raise ValueError(f"this layer was packed at bitwidth {self.bitwidth}, called with {bitwidth}")
        _require_cuda(x, "input")
        K, N = self.in_features, self.out_features
        x2 = x.reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        q, s = _ActQuantCache.get(x2)
        y = torch.empty((x2.shape[0], N), device=x.device, dtype=x.dtype)
        check(lib.ob_gemm_tern_i8_fwd(q.data_ptr(), s.data_ptr(), self.packed.data_ptr(), self.alpha.data_ptr(),
                                      OB_ALPHA_RAW, None if self.bias is None else self.bias.data_ptr(), x2.shape[0], N, K,
                                      y.data_ptr(), _tag(y), _stream()))
        return y.view(*x.shape[:-1], N)


def pack_model_for_inference(model: nn.Module, bitwidth: int) -> nn.Module:
    """Replace every ``QuantizedLinear`` in ``model`` (in place) by its packed form at ``bitwidth``."""
    for parent in list(model.modules()):
        for name, child in list(parent.named_children()):
            if isinstance(child, QuantizedLinear):
                setattr(parent, name, PackedQuantizedLinear.from_layer(child, bitwidth))
    return model.eval()


def packed_state_dict(model: nn.Module, bitwidth: int) -> Dict[str, torch.Tensor]:
    """{<prefix>.packed, <prefix>.alpha, <prefix>.bias} for every routed layer + every other tensor unchanged."""
    out = {}
    routed = {n for n, m in model.named_modules() if isinstance(m, QuantizedLinear)}
    for name, m in model.named_modules():
        if name in routed:
            packed, _ = m.packed_weight(bitwidth)
            out[f"{name}.packed"] = packed.clone()
            out[f"{name}.alpha"] = m.alpha.detach().clone()
            if m.bias is not None:
                out[f"{name}.bias"] = m.bias.detach().clone()
    for k, v in model.state_dict().items():
        if k.rsplit(".", 1)[0] not in routed:
            out[k] = v
    out["__bitwidth__"] = torch.tensor(bitwidth)
    return out


def load_packed_state_dict(model: nn.Module, state: Dict[str, torch.Tensor]) -> nn.Module:
    """Inverse of ``packed_state_dict`` for a freshly constructed model of the same configuration."""
    bitwidth = int(state["__bitwidth__"])
    pack_model_for_inference(model, bitwidth)
    model.load_state_dict({k: v for k, v in state.items() if k != "__bitwidth__"})
    return model


def ctc_greedy_decode(logits: torch.Tensor, lens: torch.Tensor, blank_id: int = 3):
    """Batched greedy CTC decode on the device.  logits [B, T, V] (fp32 or bf16), lens [B] valid frames.
    Returns (tokens int32 [B, T] compacted and padded with -1, out_lens int32 [B])."""
    _require_cuda(logits, "logits")
    if logits.dtype not in _DTYPE_TAG:
        raise ValueError(f"onebit_b200: unsupported logits dtype {logits.dtype}")
    B, T, V = logits.shape
    lg = logits.detach().contiguous()
    ln = lens.to(device=logits.device, dtype=torch.int32).contiguous()
    toks = torch.empty((B, T), device=logits.device, dtype=torch.int32)
    out_lens = torch.empty((B,), device=logits.device, dtype=torch.int32)
    ws = torch.empty(lib.ob_ctc_decode_workspace_bytes(B, T), device=logits.device, dtype=torch.uint8)
    check(lib.ob_ctc_greedy_decode(lg.data_ptr(), _tag(lg), B, T, V, ln.data_ptr(), blank_id, toks.data_ptr(),
                                   out_lens.data_ptr(), ws.data_ptr(), _stream()))
    return toks, out_lens


def ctc_greedy_decode_lists(logits: torch.Tensor, lens: torch.Tensor, blank_id: int = 3) -> List[List[int]]:
    """Same, as Python lists (one device->host copy for the whole batch)."""
    toks, n = ctc_greedy_decode(logits, lens, blank_id)
    toks, n = toks.cpu(), n.cpu().tolist()
    return [toks[b, : n[b]].tolist() for b in range(len(n))]


@torch.no_grad()
def transcribe_greedy(model, batch, precision: int = 2, blank_id: int = 3):
    """Encoder forward at ``precision`` + CTC head + greedy decode (the eval loop of eval.py:118-128 with the
    Python beam search replaced by the device-side greedy decoder)."""
    _, mask, ctc_logits = model(batch, precision)
    return ctc_greedy_decode(ctc_logits, mask.sum(dim=1), blank_id)


class GraphedTranscriber:
    """``transcribe_greedy`` captured in a CUDA graph for one (batch, frames) shape: the ~1300 launches of an inference
    pass are replayed with one driver call, which is what bounds small batches (the pass is launch-latency-bound
    below batch ~64).  Weights must be frozen (packed); feed inputs through ``__call__`` - they are copied into the
    static buffers the graph was captured on."""

    def __init__(self, model, batch_size: int, frames: int, mel: int = 80, precision: int = 2, blank_id: int = 3,
                 device=None):
        device = device or next(model.parameters()).device
        self.model, self.precision, self.blank_id = model.eval(), precision, blank_id
        self.feats = torch.zeros(batch_size, frames, mel, device=device)
        self.feat_lens = torch.full((batch_size,), frames, device=device, dtype=torch.long)
        batch = {"feats": self.feats, "feat_lens": self.feat_lens}
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(2):                                  # warm-up: lazy initialisations outside the capture
                _ActQuantCache.clear()
                transcribe_greedy(model, batch, precision, blank_id)
            _ActQuantCache.clear()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self.tokens, self.out_lens = transcribe_greedy(model, batch, precision, blank_id)
            _ActQuantCache.clear()                              # the cache must not hand captured buffers to eager calls
        torch.cuda.current_stream(device).wait_stream(side)

    @torch.no_grad()
    def __call__(self, feats: torch.Tensor, feat_lens: torch.Tensor):
        self.feats.copy_(feats, non_blocking=True)
        self.feat_lens.copy_(feat_lens, non_blocking=True)
        self.graph.replay()
        return self.tokens, self.out_lens
