"""Audit of which implementation each op *around* the quantised layer took.

The quantised layer itself (quant.py) has a single implementation and rejects anything it cannot run.  The callers on either
side of it (LayerNorm, module tails, attention core, convolution module, front-end, non-routed linears)
keep the reference's torch ops for tensors the kernels do not cover (CPU tensors in the oracle-driven tests, other dtypes,
unsupported widths).  So that such a choice is never silent on a GPU, every dispatch site reports here:

* ``counts()`` -> ``{"library": {op: n}, "torch": {op: n}}`` - calls on CUDA tensors only (bench.py prints it);
* ``OB_STRICT_ROUTES=1`` (or ``strict(True)``): a CUDA tensor that would take the torch route raises instead, unless that op was
  switched off on purpose with ``OB_TORCH_NONROUTED`` (A/B measurements).
"""
from __future__ import annotations

import collections
import os

import torch

_library: collections.Counter = collections.Counter()
_torch: collections.Counter = collections.Counter()
_strict = os.environ.get("OB_STRICT_ROUTES", "") == "1"


def strict(on: bool) -> None:
    global _strict
    _strict = bool(on)


def reset() -> None:
    _library.clear()
    _torch.clear()


def counts() -> dict:
    return {"library": dict(_library), "torch": dict(_torch)}


def taken(op: str, on_library: bool, t: torch.Tensor, switched_off: bool = False) -> bool:
    """Record the route of one call of ``op`` whose deciding tensor is ``t``; returns ``on_library`` so that a dispatch site can
    write ``if routes.taken("attn", usable(...), x): ...``."""
    if on_library:
        _library[op] += 1
    elif t.is_cuda:
        _torch[op] += 1
        if _strict and not switched_off:
            raise RuntimeError(f"{op}: CUDA tensor of dtype {t.dtype}, shape {tuple(t.shape)} is not covered by the B200 kernels and "
                               "OB_STRICT_ROUTES=1 forbids the torch route")
    return on_library
