"""Golden fixture for the Conformer co-training step, produced by EXECUTING the unmodified reference on CPU.

    python tests/golden/make_golden_conformer.py      # build container only (needs /root/reference)

The reference model (onebit_asr/conformer.py + quant.py) and its own loss functions (onebit_asr/losses.py) are
imported from where they lie; the arithmetic of train.py:83-120 (three passes, two KL terms, one backward, clip,
AdamW) is driven with a FIXED stochastic-precision mask and dropout 0 so that it is reproducible.  Stored: the
batch, per-step losses of a 4-step run at batch 3 (the reference's working batch size is < 8), and gradient norms
of step 0.  "A" = pure reference; "B" = the same model with the Oracle-B activation quantiser in front of every
routed projection (our spec; built from this repo's module tree with the oracle layer, after checking that tree
against the reference's state_dict and outputs).
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("ONEBIT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(REF, "onebit_asr"), REF, ROOT]

import conformer as refc            # noqa: E402  reference conformer.py (flat `from quant import ...` resolves to the reference)
import losses as refl               # noqa: E402  reference losses.py

CFG = dict(input_dim=80, vocab_size=64, enc_layers=3, dec_layers=1, enc_dropout=0.0, dec_dropout=0.0)
B, T, U = 3, 131, 9
SP_MASKS = [[1, 0, 1], [0, 0, 1], [1, 1, 0], [0, 1, 1]]
LR = 5e-4


def make_batch():
    g = torch.Generator().manual_seed(2024)
    feats = torch.randn(B, T, 80, generator=g)
    feat_lens = torch.tensor([131, 100, 64])
    tokens = torch.randint(4, 64, (B, U), generator=g)
    token_lens = torch.tensor([9, 7, 5])
    tokens[1, 7:] = 0
    tokens[2, 5:] = 0
    return {"feats": feats, "feat_lens": feat_lens, "tokens": tokens, "token_lens": token_lens}


def ref_loss(model, batch, sp_mask):
    """train.py:83-111 with the reference's own functions."""
    bos, eos, pad, blank = 1, 2, 0, 3
    gamma, lam1, lam2 = 0.2, 0.5, 1.0
    enc2, mask2, ctc2 = model(batch, precision=2)
    t_inp, t_out, t_pad = refl.make_att_targets(batch["tokens"], bos, eos, pad)
    logits2 = model.decode_logits(enc2, mask2, t_inp, t_pad)
    l2 = (1 - gamma) * refl.att_ce_loss(logits2, t_out, pad, label_smoothing=0.1) + gamma * refl.ctc_loss_from_logits(
        ctc2, mask2.sum(1).long(), batch["tokens"], batch["token_lens"], blank)
    enc1, mask1, ctc1 = model(batch, precision=1)
    logits1 = model.decode_logits(enc1, mask1, t_inp, t_pad)
    l1 = (1 - gamma) * refl.att_ce_loss(logits1, t_out, pad, label_smoothing=0.1) + gamma * refl.ctc_loss_from_logits(
        ctc1, mask1.sum(1).long(), batch["tokens"], batch["token_lens"], blank)
    kl1 = refl.kl_logits(logits1, logits2.detach(), t_pad)
    encs, masks, ctcs = model(batch, precision=2, sp_mask=sp_mask)
    logitss = model.decode_logits(encs, masks, t_inp, t_pad)
    ls = (1 - gamma) * refl.att_ce_loss(logitss, t_out, pad, label_smoothing=0.1) + gamma * refl.ctc_loss_from_logits(
        ctcs, masks.sum(1).long(), batch["tokens"], batch["token_lens"], blank)
    kls = refl.kl_logits(logitss, logits2.detach(), t_pad)
    return l2 + lam1 * (l1 + ls) + lam2 * (kl1 + kls)


def run(model, batch, loss_fn):
    opt = torch.optim.AdamW(model.parameters(), lr=LR, betas=(0.9, 0.98), weight_decay=1e-2)
    losses, norms = [], {}
    for step, spm in enumerate(SP_MASKS):
        opt.zero_grad()
        loss = loss_fn(model, batch, spm)
        loss.backward()
        if step == 0:
            for name, p in model.named_parameters():
                if p.grad is not None and (name.endswith("alpha") or "blocks.0." in name or "ctc_head" in name):
                    norms[name] = float(p.grad.double().norm())
        total = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=5.0)
        if step == 0:
            norms["__total__"] = float(total)
        opt.step()
        losses.append(float(loss))
    return losses, norms


def main():
    torch.set_num_threads(1)
    batch = make_batch()
    torch.manual_seed(11)
    ref_model = refc.ConformerASR(**CFG)
    ref_model.train()
    init_state = {k: v.clone() for k, v in ref_model.state_dict().items()}
    with torch.no_grad():
        enc2, _, ctc2 = ref_model(batch, precision=2)
        enc32, _, _ = ref_model(batch, precision=32)
    losses_a, norms_a = run(ref_model, batch, ref_loss)

    # Oracle-B through this repo's module tree with the oracle layer
    import onebit_b200 as ob
    from onebit_b200.training import StepConfig, cotraining_loss
    from oracle.torch_oracle import OracleQuantizedLinear

    def build(act_bits):
        torch.manual_seed(11)
        OracleQuantizedLinear.act_bits_default = act_bits
        m = ob.ConformerASR(**CFG, linear_cls=OracleQuantizedLinear)
        m.train()
        return m

    ours_a = build(32)
    assert all(torch.equal(v, init_state[k]) for k, v in ours_a.state_dict().items()), "module tree / init order differs"
    cfg = StepConfig()
    la2, _ = run(ours_a, batch, lambda m, b, spm: cotraining_loss(m, b, cfg, spm)[0])
    assert np.allclose(la2, losses_a, rtol=1e-5), (la2, losses_a)
    ours_b = build(8)
    losses_b, norms_b = run(ours_b, batch, lambda m, b, spm: cotraining_loss(m, b, cfg, spm)[0])
    OracleQuantizedLinear.act_bits_default = 8

    out = {k: v.numpy() for k, v in batch.items()}
    out.update(losses_A=np.array(losses_a), losses_B=np.array(losses_b),
               enc2_A_sum=np.float64(enc2.double().sum()), enc2_A_absmean=np.float64(enc2.abs().mean()),
               enc32_A_absmean=np.float64(enc32.abs().mean()), ctc2_A_absmean=np.float64(ctc2.abs().mean()),
               norm_names=np.array(sorted(norms_a)), norms_A=np.array([norms_a[k] for k in sorted(norms_a)]),
               norms_B=np.array([norms_b[k] for k in sorted(norms_a)]),
               sp_masks=np.array(SP_MASKS), seed=np.int64(11), lr=np.float64(LR))
    np.savez_compressed(os.path.join(HERE, "conformer_step.npz"), **out)
    print("losses A", losses_a)
    print("losses B", losses_b)
    print("total grad norm A/B", norms_a["__total__"], norms_b["__total__"])


if __name__ == "__main__":
    main()
