"""CPU baseline driver: the UNMODIFIED reference (``baseline/_ref``, see build_ref.py) timed on the host cores.
BASELINE INFRASTRUCTURE - used only by ``bench.py --impl reference`` and by bench.py's ``cpu_baseline`` leg (which runs this in
a child process so that the product's process never imports the reference and this one never imports the product).

Workload = BASELINE.json configs[0] / BASELINE.md section 3 ("config 1"): default dims (train.py:194-203), V = 5004, batch 4 x 1000
frames x 80 mel, 40 tokens per utterance, weights seed 0, inputs seed 1, AdamW(lr 5e-4, betas (0.9, 0.98), wd 1e-2)
(train.py:259), all host threads.  Two figures:

  full_step : the reference's own ``train.py:run_epoch`` (62-169) on a one-batch loader - three passes, five losses, one backward,
              clip 5.0, AdamW, ``loss.item()`` - exactly the step the GPU arm times;
  one_pass  : one precision-2 forward + backward + clip + AdamW with the reference's model and loss functions.
"""
import os
import time
import types

import torch

from . import ref_loader

FRAME_S = 0.010
CONFIG1 = dict(batch=4, frames=1000, mel=80, vocab=5004, tokens=40)


def _batch(cfg):
    g = torch.Generator().manual_seed(1)
    B, T, U = cfg["batch"], cfg["frames"], cfg["tokens"]
    return {"feats": torch.randn(B, T, cfg["mel"], generator=g), "feat_lens": torch.full((B,), T, dtype=torch.long),
            "tokens": torch.randint(4, cfg["vocab"], (B, U), generator=g), "token_lens": torch.full((B,), U, dtype=torch.long)}


class _OneBatchLoader:
    """What run_epoch needs from the reference's data module (train.py:66, 78): a loader and the special ids
    (dataloader_stub.py:199-207)."""

    def __init__(self, batch, steps):
        self.batch, self.steps = batch, steps

    def train_dataloader(self):
        return [self.batch] * self.steps

    def special_ids(self):
        return dict(bos_id=1, eos_id=2, pad_id=0, blank_id=3)


def run(steps: int, warmup: int, dropout: float = 0.1, cfg=None, budget_s: float = 240.0, one_pass: bool = True):
    """Returns a dict with audio-s/s and s/step of the full reference step (and of one precision-2 pass)."""
    cfg = dict(CONFIG1 if cfg is None else cfg)
    ref = ref_loader.load(with_train=True)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = ref.conformer.ConformerASR(cfg["mel"], cfg["vocab"], enc_dropout=dropout, dec_dropout=dropout)
    opt = torch.optim.AdamW(model.parameters(), lr=5e-4, betas=(0.9, 0.98), weight_decay=1e-2)
    batch = _batch(cfg)
    args = types.SimpleNamespace(enc_layers=len(model.encoder.blocks))
    audio_s = cfg["batch"] * cfg["frames"] * FRAME_S

    def full_step(n=1):
        return ref.train.run_epoch(model, _OneBatchLoader(batch, n), opt, None, "cpu", args, True, 0.5, 1.0, 0.2)[0]

    t_start = time.perf_counter()
    for _ in range(max(1, warmup)):
        full_step()
        if time.perf_counter() - t_start > 0.4 * budget_s:
            break
    times, last = [], float("nan")
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        last = full_step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    dt = sum(times) / len(times)
    out = {"audio_s_per_s": audio_s / dt, "s_per_step": dt, "steps_timed": len(times), "last_loss": float(last),
           "threads": torch.get_num_threads(), "cores": threads, "config": cfg, "dropout": dropout}
    if one_pass:
        L = ref.losses
        model.train()

        def pass2():
            enc, mask, ctc = model(batch, precision=2)
            t_inp, t_out, t_pad = L.make_att_targets(batch["tokens"], 1, 2, 0)
            logits = model.decode_logits(enc, mask, t_inp, t_pad)
            loss = 0.8 * L.att_ce_loss(logits, t_out, 0, label_smoothing=0.1) + 0.2 * L.ctc_loss_from_logits(
                ctc, mask.sum(dim=1).long(), batch["tokens"], batch["token_lens"], 3)
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=5.0)
            opt.step()
            return loss.item()
        pass2()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            pass2()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        out["one_pass"] = {"s_per_step": ts[1], "audio_s_per_s": audio_s / ts[1]}
    return out
