"""Golden vectors for the CTC loss: outputs of the reference's own ``ctc_loss_from_logits`` (onebit_asr/losses.py:41-47)
executed on CPU in the build container (needs /root/reference; run:  python tests/golden/make_golden_ctc.py).

Cases cover what the domain offers: ragged input and target lengths, repeated labels (no skip transition between them),
an empty target, a target that does not fit its input (``zero_infinity`` zeroes loss and gradient), padding frames beyond
the input length (zero gradient), blank id 3 as in the training step.  Writes tests/golden/kat_ctc.npz.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("ONEBIT_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(REF, "onebit_asr"), REF]
import losses as ref_losses  # noqa: E402  (the reference's onebit_asr/losses.py)

HERE = os.path.dirname(os.path.abspath(__file__))


def run_case(seed, B, T, V, Lmax, in_lens, tgt_lens, blank, targets=None, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    logits = (scale * torch.randn(B, T, V, generator=g)).requires_grad_(True)
    if targets is None:
        targets = torch.randint(0, V - 1, (B, Lmax), generator=g)
        targets = targets + (targets >= blank).long()            # never the blank id
    in_lens, tgt_lens = torch.tensor(in_lens), torch.tensor(tgt_lens)
    loss = ref_losses.ctc_loss_from_logits(logits, in_lens, targets, tgt_lens, blank)
    (grad,) = torch.autograd.grad(3.0 * loss, logits)                 # upstream gradient 3.0
    return dict(logits=logits.detach().numpy(), targets=targets.numpy(), in_lens=in_lens.numpy(), tgt_lens=tgt_lens.numpy(),
                blank=np.int64(blank), loss=loss.detach().numpy(), grad=grad.numpy(), grad_out=np.float32(3.0))


def main():
    cases = {
        "ragged": run_case(1, 4, 12, 9, 5, [12, 9, 7, 12], [5, 3, 1, 4], 3),
        "repeats": run_case(2, 2, 10, 6, 4, [10, 8], [4, 3], 0, targets=torch.tensor([[2, 2, 2, 5], [1, 1, 4, 0]])),
        "empty_and_infeasible": run_case(3, 3, 6, 7, 4, [6, 3, 6], [0, 4, 2], 3,
                                         targets=torch.tensor([[0, 0, 0, 0], [1, 1, 2, 2], [5, 6, 0, 0]])),
        "longer": run_case(4, 3, 60, 40, 12, [60, 41, 55], [12, 7, 10], 3, scale=2.0),
    }
    flat = {f"{name}.{k}": v for name, c in cases.items() for k, v in c.items()}
    np.savez_compressed(os.path.join(HERE, "kat_ctc.npz"), **flat)
    for name, c in cases.items():
        print(name, "loss", float(c["loss"]), "|grad|max", float(np.abs(c["grad"]).max()))


if __name__ == "__main__":
    main()
