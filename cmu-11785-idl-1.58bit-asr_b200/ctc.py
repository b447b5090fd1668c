"""CTC loss of the co-training step on the B200 kernels (csrc/ob_ctc.cu).

Same value and gradient as the reference's ``ctc_loss_from_logits`` (onebit_asr/losses.py:41-47: ``log_softmax``, transpose,
``nn.CTCLoss(blank, zero_infinity=True)`` with mean reduction), computed from the ``[B, T, V]`` logits without materialising
log-probabilities or the time-major copy: a row log-sum-exp, the forward/backward recursions on gathered entries, and one
pass that writes the gradient with respect to the logits.
"""
from __future__ import annotations

import os

import torch

from ._cabi import check, lib
from .quant import _stream

MAX_TARGET_LEN = 511
# the training step's CTC loss takes the library route unless OB_CTC=0 or "ctc" is in OB_TORCH_NONROUTED (A/B measurements)
ENABLED = os.environ.get("OB_CTC", "1") != "0"


def usable(logits: torch.Tensor, targets: torch.Tensor) -> bool:
    from .matmul import DISABLED
    return (ENABLED and "ctc" not in DISABLED and logits.is_cuda and logits.dtype == torch.float32 and logits.dim() == 3
            and logits.numel() > 0 and logits.stride(2) == 1 and logits.stride(0) == logits.shape[1] * logits.stride(1)
            and targets.dim() == 2 and targets.shape[1] <= MAX_TARGET_LEN)


def _lens(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.int64, non_blocking=True).contiguous()


class _CtcLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, in_lens, targets, tgt_lens, blank):
        B, T, V = logits.shape
        Lmax = targets.shape[1]
        dev = logits.device
        in_lens, tgt_lens = _lens(in_lens, dev), _lens(tgt_lens, dev)
        targets = _lens(targets, dev)
        if in_lens.numel() != B or tgt_lens.numel() != B or targets.shape[0] != B:
            raise ValueError(f"ctc_loss: batch of {B} utterances but {in_lens.numel()} input lengths, {tgt_lens.numel()} target "
                             f"lengths, {targets.shape[0]} target rows")
        Sp = lib.ob_ctc_state_pitch(Lmax)
        lse = torch.empty(B * T, device=dev, dtype=torch.float32)
        ab = torch.empty(2, B, T, Sp, device=dev, dtype=torch.float32)          # forward and backward variables
        nll = torch.empty(B, device=dev, dtype=torch.float32)
        loss = torch.empty((), device=dev, dtype=torch.float32)
        ld = logits.stride(1)
        check(lib.ob_ctc_loss_fwd(logits.data_ptr(), ld, in_lens.data_ptr(), targets.data_ptr() if Lmax else None,
                                  targets.stride(0) if Lmax else 0, tgt_lens.data_ptr(), B, T, V, Lmax, int(blank), lse.data_ptr(),
                                  ab[0].data_ptr(), ab[1].data_ptr(), nll.data_ptr(), loss.data_ptr(), _stream()))
        ctx.save_for_backward(logits, in_lens, targets, tgt_lens, lse, ab, nll)
        ctx.blank = int(blank)
        return loss

    @staticmethod
    def backward(ctx, g):
        logits, in_lens, targets, tgt_lens, lse, ab, nll = ctx.saved_tensors
        B, T, V = logits.shape
        Lmax = targets.shape[1]
        g = g.to(torch.float32).contiguous()
        dx = torch.empty((B, T, V), device=logits.device, dtype=torch.float32)
        check(lib.ob_ctc_loss_bwd(logits.data_ptr(), logits.stride(1), in_lens.data_ptr(), targets.data_ptr() if Lmax else None,
                                  targets.stride(0) if Lmax else 0, tgt_lens.data_ptr(), B, T, V, Lmax, ctx.blank, lse.data_ptr(),
                                  ab[0].data_ptr(), ab[1].data_ptr(), nll.data_ptr(), g.data_ptr(), dx.data_ptr(), V, _stream()))
        return dx, None, None, None, None


def ctc_loss(logits: torch.Tensor, in_lens: torch.Tensor, targets: torch.Tensor, tgt_lens: torch.Tensor, blank: int) -> torch.Tensor:
    """Mean over utterances of ``nll_b / max(len(target_b), 1)`` with infeasible alignments zeroed; logits ``[B, T, V]`` fp32 on
    the device, targets ``[B, Lmax]`` (padded), lengths ``[B]`` (host or device tensors)."""
    if not logits.is_cuda or logits.dtype != torch.float32:
        raise RuntimeError(f"ctc_loss: logits must be a CUDA float32 tensor (got {logits.dtype} on {logits.device}); there is no fallback")
    return _CtcLossFn.apply(logits, in_lens, targets, tgt_lens, blank)
