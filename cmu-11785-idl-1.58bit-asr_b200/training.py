"""Co-training step of the reference (onebit_asr/train.py:62-120, onebit_asr/losses.py) for the B200 layer.

One optimiser step = three encoder passes over the same batch with shared latent weights - 2-bit teacher, 1-bit
student, stochastic-precision mix - each with attention-CE + CTC, two KL terms to the detached teacher, ONE
backward, grad-norm clip 5.0, AdamW (train.py:83-120).  The math is restated here so that the workload runs where
the reference tree is not mounted and under data parallelism (``dp.GradAllReducer``); tests/test_conformer_cpu.py and
tests/test_gpu_conformer.py check it against fixtures produced by the reference's own ``run_epoch`` arithmetic
(tests/golden/make_golden_conformer.py).

Differences from the reference are host-side only: the CTC lengths are taken from the host copy of ``feat_lens``
(no device->host sync inside the step; same values as ``mask.sum(1)``), and the conv subsampling front-end - which
is bitwidth-independent and dropout-free - may be computed once and shared by the three passes
(``share_frontend=True``; gradients are identical because the three uses sum).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


@dataclass
class StepConfig:
    """Paper constants and ids, defaults of train.py:186-211 and dataloader_stub.py:199-207."""
    gamma_ctc: float = 0.2
    lambda1: float = 0.5
    lambda2: float = 1.0
    label_smoothing: float = 0.1
    max_grad_norm: float = 5.0
    bos_id: int = 1
    eos_id: int = 2
    pad_id: int = 0
    blank_id: int = 3
    share_frontend: bool = False
    stack_passes: bool = False      # evaluate the three encoder passes side by side on one stacked batch (same math)


# ---------------------------------------------------------------------------------------------- losses (losses.py)
def make_att_targets(tokens: torch.Tensor, bos_id: int, eos_id: int, pad_id: int):
    """Decoder input <bos>+y, target y+<eos>, pad mask of the input (losses.py:11-19)."""
    col = tokens.new_empty((tokens.size(0), 1))
    tgt_inp = torch.cat([col.fill_(bos_id).clone(), tokens], dim=1)
    tgt_out = torch.cat([tokens, col.fill_(eos_id).clone()], dim=1)
    return tgt_inp, tgt_out, tgt_inp == pad_id


def att_ce_loss(logits: torch.Tensor, targets: torch.Tensor, pad_id: int, label_smoothing: float = 0.0):
    """Label-smoothed CE.  As in the reference the smoothed branch averages over ALL positions: its pad mask
    multiplies an already-reduced scalar and cancels (losses.py:22-38); kept literally for parity."""
    if label_smoothing <= 0:
        return F.cross_entropy(logits.transpose(1, 2), targets, ignore_index=pad_id)
    logp = F.log_softmax(logits, dim=-1)
    with torch.no_grad():
        dist = torch.full_like(logp, label_smoothing / (logits.size(-1) - 1))
        dist.scatter_(2, targets.unsqueeze(-1), 1.0 - label_smoothing)
    loss = torch.mean(torch.sum(-dist * logp, dim=-1))
    keep = (targets != pad_id).float()
    return (loss * keep).sum() / keep.sum().clamp_min(1.0)


def ctc_loss_from_logits(ctc_logits, feat_lens, tokens, token_lens, blank_id: int):
    """nn.CTCLoss(blank, zero_infinity=True) on log-softmaxed [T,B,V] (losses.py:41-47); on the device the same value and
    gradient come from the library's kernels straight from the [B,T,V] logits (ctc.py)."""
    from . import ctc, matmul, routes
    if routes.taken("ctc_loss", ctc.usable(ctc_logits, tokens), ctc_logits, "ctc" in matmul.DISABLED or not ctc.ENABLED):
        return ctc.ctc_loss(ctc_logits, feat_lens, tokens, token_lens, blank_id)
    logp = F.log_softmax(ctc_logits, dim=-1).transpose(0, 1)
    return F.ctc_loss(logp, tokens, feat_lens, token_lens, blank=blank_id, reduction="mean", zero_infinity=True)


def kl_logits(student_logits, teacher_logits, pad_mask):
    """KL(stop-grad teacher || student) averaged over non-pad decoder positions (losses.py:50-59)."""
    with torch.no_grad():
        p_t = F.softmax(teacher_logits, dim=-1)
    kl = F.kl_div(F.log_softmax(student_logits, dim=-1), p_t, reduction="none").sum(dim=-1)
    keep = (~pad_mask).float()
    return (kl * keep).sum() / keep.sum().clamp_min(1.0)


# ---------------------------------------------------------------------------------------------- schedule helpers
def sample_sp_mask(n_layers: int, low_p: float = 0.2, high_p: float = 0.9) -> List[int]:
    """Per-layer Bernoulli(1-bit) with log-spaced probabilities, drawn from the global CPU RNG (train.py:56-59)."""
    probs = torch.logspace(math.log10(low_p), math.log10(high_p), steps=n_layers)
    return [int(torch.rand(()) < p) for p in probs]


class WarmupCosine:
    """Linear warm-up then cosine decay to min_lr_ratio (train.py:32-53)."""

    def __init__(self, optimizer, warmup_steps: int, total_steps: int, min_lr_ratio: float = 0.1):
        self.optimizer, self.warmup_steps, self.total_steps = optimizer, warmup_steps, total_steps
        self.min_lr_ratio, self.step_num = min_lr_ratio, 0
        for group in optimizer.param_groups:
            group.setdefault("initial_lr", group.get("lr", 1e-3))

    def scale(self) -> float:
        if self.step_num < self.warmup_steps:
            return self.step_num / max(1, self.warmup_steps)
        frac = (self.step_num - self.warmup_steps) / max(1, self.total_steps - self.warmup_steps)
        frac = min(max(frac, 0.0), 1.0)
        return self.min_lr_ratio + 0.5 * (1 - self.min_lr_ratio) * (1 + math.cos(math.pi * frac))

    def step(self) -> None:
        self.step_num += 1
        s = self.scale()
        for group in self.optimizer.param_groups:
            group["lr"] = group["initial_lr"] * s


# ---------------------------------------------------------------------------------------------- the step
def cotraining_loss(model, batch: Dict[str, torch.Tensor], cfg: StepConfig, sp_mask: Optional[List[int]] = None,
                    n_layers: Optional[int] = None):
    """Lint2 + lambda1 (Lint1 + Lint_sp) + lambda2 (KL1 + KL_sp)  (train.py:83-111).  Returns (loss, parts)."""
    tokens, token_lens = batch["tokens"], batch["token_lens"]
    t_inp, t_out, t_pad = make_att_targets(tokens, cfg.bos_id, cfg.eos_id, cfg.pad_id)
    if sp_mask is None:
        sp_mask = sample_sp_mask(n_layers if n_layers is not None else len(model.encoder.blocks))
    # CTC input lengths = number of valid encoder frames = min(T_sub, feat_lens // 4): from the host copy if given
    lens_host = batch.get("feat_lens_cpu")
    tok_lens_host = batch.get("token_lens_cpu", token_lens)
    shared = model.encoder.frontend(batch["feats"]) if cfg.share_frontend else None
    # device copies of the CTC lengths for the library's CTC kernels (which clamp to the number of frames themselves)
    from . import ctc as ctc_kernels
    feat_lens_dev = batch.get("feat_lens")
    lens_dev = feat_lens_dev // 4 if (feat_lens_dev is not None and feat_lens_dev.is_cuda and token_lens.is_cuda) else None

    stacked = None
    if cfg.stack_passes and hasattr(model, "forward_passes"):
        # the teacher, student and stochastic-precision encoder passes share every weight: one stacked batch, each layer
        # applied once per bitwidth group (same values as three separate passes)
        stacked = dict(zip(((2, None), (1, None), (2, "sp")), model.forward_passes(batch, [(2, None), (1, None), (2, sp_mask)],
                                                                                    frontend_out=shared)))

    def one_pass(precision, mask_list=None):
        if stacked is not None:
            enc, mask, ctc = stacked[(precision, None if mask_list is None else "sp")]
        elif shared is not None:
            enc, mask, ctc = model(batch, precision, mask_list, frontend_out=shared)
        else:
            enc, mask, ctc = model(batch, precision, mask_list)
        logits = model.decode_logits(enc, mask, t_inp, t_pad)
        l_att = att_ce_loss(logits, t_out, cfg.pad_id, cfg.label_smoothing)
        if lens_host is not None:
            ctc_lens = torch.clamp(lens_host // 4, max=enc.size(1))
        else:
            ctc_lens = mask.sum(dim=1).long()
        if lens_dev is not None and ctc_kernels.usable(ctc, tokens):
            l_ctc = ctc_loss_from_logits(ctc, lens_dev, tokens, token_lens, cfg.blank_id)
        else:
            l_ctc = ctc_loss_from_logits(ctc, ctc_lens, tokens, tok_lens_host, cfg.blank_id)
        return (1 - cfg.gamma_ctc) * l_att + cfg.gamma_ctc * l_ctc, logits

    l2, logits2 = one_pass(2)                                  # teacher
    l1, logits1 = one_pass(1)                                  # student
    kl1 = kl_logits(logits1, logits2.detach(), t_pad)
    ls, logits_s = one_pass(2, sp_mask)                        # stochastic precision
    kls = kl_logits(logits_s, logits2.detach(), t_pad)
    loss = l2 + cfg.lambda1 * (l1 + ls) + cfg.lambda2 * (kl1 + kls)
    return loss, {"Lint2": l2, "Lint1": l1, "Lint_sp": ls, "KL1": kl1, "KL_sp": kls}


def reserve_allocator_headroom(device, gib: float = 6.0) -> int:
    """Grow torch's caching allocator by one free, splittable block of ``gib`` GiB on ``device``; returns the bytes reserved.

    Why: the stochastic-precision mask (train.py:56-59) changes the row split of every grouped layer from step to step, so
    the sizes the allocator is asked for keep shifting a little.  With the reserved pool only ~2 GiB above the peak of live
    tensors a request can miss the cached blocks several steps into a run, and the ``cudaMalloc`` it falls back to stalls the
    enqueueing thread until the device has drained - measured as one 207 ms step among 122 ms steps
    (profiles/r01_step_times.json).  A cached block with room to split absorbs those requests.  Call it once after the first
    step; a no-op for CPU devices or ``gib <= 0``."""
    device = torch.device(device)
    if device.type != "cuda" or gib <= 0:
        return 0
    n = int(gib * 2 ** 30)
    block = torch.empty(n, dtype=torch.uint8, device=device)      # released at once: stays in the allocator's large pool
    # requests below 1 MiB come from a separate pool of 2 MiB segments: 64 spare segments for those
    small = [torch.empty(512 * 1024, dtype=torch.uint8, device=device) for _ in range(256)]
    del block, small
    return n + 256 * 512 * 1024


def train_step(model, batch, optimizer, cfg: StepConfig, sched=None, sp_mask=None, grad_sync=None):
    """zero_grad -> 3-pass loss -> backward -> (DP all-reduce) -> clip -> AdamW -> schedule (train.py:114-120).

    ``batch``: one batch dict, or a list of micro-batches whose gradients are accumulated before the one optimiser step
    (loss = mean over the micro-batches: the same update as one batch holding all their utterances, except that BatchNorm
    uses each micro-batch's own statistics - as under data parallelism, conformer.py:148).
    ``grad_sync``: a ``dp.GradAllReducer`` whose hooks launch bucketed all-reduces during (the last) backward; ``finish()``
    waits for them before the global-norm clip."""
    optimizer.zero_grad(set_to_none=True)
    micro = batch if isinstance(batch, (list, tuple)) else [batch]
    total, parts = None, None
    for i, mb in enumerate(micro):
        if grad_sync is not None:
            grad_sync.enabled = i == len(micro) - 1          # all-reduce the accumulated gradients once
        loss, parts = cotraining_loss(model, mb, cfg, sp_mask)
        if len(micro) > 1:
            loss = loss / len(micro)
        loss.backward()
        total = loss.detach() if total is None else total + loss.detach()
    if grad_sync is not None:
        grad_sync.finish()
    torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=cfg.max_grad_norm)
    optimizer.step()
    if sched is not None:
        sched.step()
    return total, parts
