"""Per-step device and host times of the bench's training step on a B200 (no synchronisation between steps, as in the timed
region of bench.py): shows whether early steps pay for allocator growth.  python tools/gpu_step_times.py [steps]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import onebit_b200 as ob  # noqa: E402
from onebit_b200 import routes  # noqa: E402
from onebit_b200.training import StepConfig, reserve_allocator_headroom, train_step  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 14
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = ob.ConformerASR(bench.TRAIN["mel"], bench.TRAIN["vocab"], enc_dropout=0.1, dec_dropout=0.1).train().to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=5e-4, betas=(0.9, 0.98), weight_decay=1e-2, fused=True)
cfg = StepConfig(share_frontend=True, stack_passes=True)
batch = bench.make_batch(bench.TRAIN["batch"], bench.TRAIN["frames"], 1000, device=dev)
events = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
host, mallocs = [], []
torch.cuda.synchronize()
events[0].record()
for i in range(n):
    t0 = time.perf_counter()
    train_step(model, batch, opt, cfg)
    if i == 0:
        reserve_allocator_headroom(dev, float(os.environ.get("OB_HEADROOM_GIB", "0")))      # as bench.py does (6 GiB there)
    events[i + 1].record()
    host.append(round((time.perf_counter() - t0) * 1e3, 1))
    mallocs.append(torch.cuda.memory_stats()["num_device_alloc"])
torch.cuda.synchronize()
dev_ms = [round(events[i].elapsed_time(events[i + 1]), 1) for i in range(n)]
print(json.dumps({"device_ms": dev_ms, "host_ms": host, "cudaMallocs_cumulative": mallocs, "routes": routes.counts(),
                  "reserved_gib": round(torch.cuda.memory_reserved() / 2**30, 1)}))
