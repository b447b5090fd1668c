"""Probe the Conformer co-training step on a B200: step time, peak memory, top kernels.  python tools/gpu_train_probe.py [B] [T]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import onebit_b200 as ob  # noqa: E402
from onebit_b200.training import StepConfig, train_step  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1600
share = len(sys.argv) > 3 and sys.argv[3] == "share"
torch.manual_seed(0)
model = ob.ConformerASR(80, 5004).train().cuda()
opt = torch.optim.AdamW(model.parameters(), lr=5e-4, betas=(0.9, 0.98), weight_decay=1e-2)
g = torch.Generator().manual_seed(1)
batch = {"feats": torch.randn(B, T, 80, generator=g).cuda(), "feat_lens": torch.full((B,), T).cuda(),
         "tokens": torch.randint(4, 5004, (B, 64), generator=g).cuda(), "token_lens": torch.full((B,), 64).cuda(),
         "feat_lens_cpu": torch.full((B,), T), "token_lens_cpu": torch.full((B,), 64)}
cfg = StepConfig(share_frontend=share, stack_passes=share and os.environ.get("OB_STACK", "1") == "1")
for _ in range(2):
    loss, _ = train_step(model, batch, opt, cfg)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 3
for _ in range(n):
    loss, _ = train_step(model, batch, opt, cfg)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
print(f"B={B} T={T} share_frontend={share}: {dt * 1e3:.1f} ms/step, {B * T * 0.01 / dt:.1f} audio-s/s, loss {loss.item():.4f}, "
      f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    train_step(model, batch, opt, cfg)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60))
# where the copies / strided adds / reductions come from: by op and input shape
by_shape = prof.key_averages(group_by_input_shape=True)
rows = [e for e in by_shape if e.key in ("aten::copy_", "aten::add", "aten::add_", "aten::sum", "aten::mul", "aten::clone", "aten::contiguous")]
rows.sort(key=lambda e: -e.device_time_total)
for e in rows[:24]:
    print(f"{e.device_time_total / 1e3:8.2f} ms {e.count:5d} x  {e.key:18s} {str(e.input_shapes)[:110]}")
# device kernels only, by name
kern = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        k = kern.setdefault(ev.name, [0, 0.0])
        k[0] += 1
        k[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in kern.values())
print(f"device kernels: {sum(v[0] for v in kern.values())} launches, {tot / 1e3:.1f} ms")
for name, (cnt, us) in sorted(kern.items(), key=lambda kv: -kv[1][1])[:70]:
    print(f"{us / 1e3:9.2f} ms {100 * us / tot:5.1f}% {cnt:5d} x {us / cnt:8.1f} us  {name[:150]}")
