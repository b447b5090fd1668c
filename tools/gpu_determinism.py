"""Run the small inference model several times on the same input and report bit-level run-to-run differences of the logits."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import onebit_b200 as ob  # noqa: E402

cfg = dict(input_dim=80, vocab_size=64, enc_layers=2, dec_layers=1, enc_dropout=0.0, dec_dropout=0.0)
torch.manual_seed(4)
model = ob.ConformerASR(**cfg).cuda().eval()
batch = {"feats": torch.randn(3, 200, 80, device="cuda"), "feat_lens": torch.tensor([200, 150, 90], device="cuda")}
outs = []
with torch.no_grad():
    for i in range(6):
        if i == 3:
            junk = [torch.randn(1 << 20, device="cuda") for _ in range(7)]   # perturb the allocator between runs
        mem, valid, logits = model(batch, 2)
        outs.append((mem.clone(), logits.clone()))
torch.cuda.synchronize()
for i in range(1, 6):
    dm = (outs[i][0] - outs[0][0]).abs().max().item()
    dl = (outs[i][1] - outs[0][1]).abs().max().item()
    print(f"run {i}: max|d memory| {dm:.3e}  max|d logits| {dl:.3e}  argmax flips {(outs[i][1].argmax(-1) != outs[0][1].argmax(-1)).sum().item()}")
print("disabled:", os.environ.get("OB_TORCH_NONROUTED", ""))
