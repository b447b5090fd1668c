"""Host-side cost of one co-training step: cProfile of train_step after warm-up (the step is CPU-bound once the kernels are fast)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import onebit_b200 as ob  # noqa: E402
from onebit_b200.training import StepConfig, train_step  # noqa: E402

B, T = 64, 1600
fused = len(sys.argv) > 1 and sys.argv[1] == "fused"
torch.manual_seed(0)
model = ob.ConformerASR(80, 5004).train().cuda()
opt = torch.optim.AdamW(model.parameters(), lr=5e-4, betas=(0.9, 0.98), weight_decay=1e-2, fused=fused)
g = torch.Generator().manual_seed(1)
batch = {"feats": torch.randn(B, T, 80, generator=g).cuda(), "feat_lens": torch.full((B,), T).cuda(),
         "tokens": torch.randint(4, 5004, (B, 64), generator=g).cuda(), "token_lens": torch.full((B,), 64).cuda(),
         "feat_lens_cpu": torch.full((B,), T), "token_lens_cpu": torch.full((B,), 64)}
cfg = StepConfig(share_frontend=True, stack_passes=os.environ.get("OB_STACK", "1") == "1")
for _ in range(3):
    train_step(model, batch, opt, cfg)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    train_step(model, batch, opt, cfg)
torch.cuda.synchronize()
print(f"fused={fused}: {(time.perf_counter() - t0) / 5 * 1e3:.1f} ms/step wall")
# host time with the GPU out of the way: time until the last launch is queued (no sync inside)
t0 = time.perf_counter()
train_step(model, batch, opt, cfg)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0):.1f} ms, drain {1e3 * (t2 - t1):.1f} ms")
pr = cProfile.Profile()
pr.enable()
train_step(model, batch, opt, cfg)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
