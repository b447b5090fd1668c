"""Golden fixture at BASELINE.json configs[0] ("config 1"): the reference Conformer at its DEFAULT dims (12 blocks, d_model 256,
d_ff 1024, 4 heads, V = 5004, 2 decoder layers), batch 4 x 1000 frames x 80 mel, 40 tokens - produced by EXECUTING the
unmodified reference on CPU.

    python tests/golden/make_golden_config1.py        # build container only (needs /root/reference); ~2 min

One co-training step (train.py:83-116: 2-bit, 1-bit and stochastic-precision passes, CTC + attention + 2 KL, one backward) with a
fixed precision mask and dropout 0.  Stored: the seeds that regenerate the batch (torch's CPU generator), the loss of the pure
reference ("A"), the loss and gradient norms of the same step with the Oracle-B activation quantiser in front of every routed
projection ("B": the reference's own conformer.py with oracle/torch_oracle.py's layer served as ``quant`` - the same seam the
product uses), and the precision-2 / precision-1 encoder outputs' statistics.  The batch itself is not stored (1.3 MB): it is a
function of the seed.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("ONEBIT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(REF, "onebit_asr"), REF, ROOT]

from make_golden_conformer import ref_loss      # noqa: E402  train.py:83-111 driven with the reference's own functions
import conformer as refc                        # noqa: E402  the reference's conformer.py (with the reference's quant.py)

B, T, U, V = 4, 1000, 40, 5004
SEED_MODEL, SEED_BATCH = 0, 1
SP_MASK = [0, 1, 0, 0, 1, 1, 0, 1, 1, 1, 0, 1]


def make_batch():
    g = torch.Generator().manual_seed(SEED_BATCH)
    feats = torch.randn(B, T, 80, generator=g)
    tokens = torch.randint(4, V, (B, U), generator=g)
    return {"feats": feats, "feat_lens": torch.tensor([1000, 1000, 870, 640]), "tokens": tokens,
            "token_lens": torch.full((B,), U, dtype=torch.long)}


def oracle_b_conformer():
    """The reference's conformer.py executed a second time with the Oracle-B layer served as its ``quant`` module."""
    import importlib.util
    from oracle.torch_oracle import OracleQuantizedLinear
    fake = types.ModuleType("quant")
    fake.QuantizedLinear = OracleQuantizedLinear
    saved = sys.modules.get("quant")
    sys.modules["quant"] = fake
    try:
        spec = importlib.util.spec_from_file_location("conformer_oracle_b", os.path.join(REF, "onebit_asr", "conformer.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.modules["quant"] = saved
    return mod


def grads(model):
    out = {n: float(p.grad.double().norm()) for n, p in model.named_parameters()
           if p.grad is not None and (n.endswith("alpha") or "blocks.0." in n or "blocks.11." in n or "ctc_head" in n)}
    out["__total__"] = float(torch.sqrt(sum(p.grad.double().pow(2).sum() for p in model.parameters() if p.grad is not None)))
    return out


def main():
    torch.set_num_threads(os.cpu_count())
    batch = make_batch()
    torch.manual_seed(SEED_MODEL)
    model_a = refc.ConformerASR(80, V, enc_dropout=0.0, dec_dropout=0.0).train()
    state = {k: v.clone() for k, v in model_a.state_dict().items()}
    loss_a = ref_loss(model_a, batch, SP_MASK)
    loss_a.backward()
    norms_a = grads(model_a)

    refb = oracle_b_conformer()
    torch.manual_seed(SEED_MODEL)
    model_b = refb.ConformerASR(80, V, enc_dropout=0.0, dec_dropout=0.0).train()
    assert all(torch.equal(v, state[k]) for k, v in model_b.state_dict().items())
    with torch.no_grad():
        enc2, _, ctc2 = model_b(batch, precision=2)
        enc1, _, _ = model_b(batch, precision=1)
    loss_b = ref_loss(model_b, batch, SP_MASK)
    loss_b.backward()
    norms_b = grads(model_b)
    names = sorted(norms_b)
    np.savez_compressed(os.path.join(HERE, "conformer_config1.npz"),
                        seed_model=np.int64(SEED_MODEL), seed_batch=np.int64(SEED_BATCH), sp_mask=np.array(SP_MASK),
                        feat_lens=batch["feat_lens"].numpy(), shape=np.array([B, T, U, V]),
                        feats_sum=np.float64(batch["feats"].double().sum()), tokens_sum=np.int64(batch["tokens"].sum()),
                        loss_A=np.float64(loss_a), loss_B=np.float64(loss_b),
                        enc2_B_absmean=np.float64(enc2.abs().mean()), enc1_B_absmean=np.float64(enc1.abs().mean()),
                        ctc2_B_absmean=np.float64(ctc2.abs().mean()),
                        norm_names=np.array(names), norms_A=np.array([norms_a[n] for n in names]),
                        norms_B=np.array([norms_b[n] for n in names]))
    print("loss A", float(loss_a), "loss B", float(loss_b), "total norm A/B", norms_a["__total__"], norms_b["__total__"])


if __name__ == "__main__":
    main()
