"""CTC loss at the training shape on a B200: library kernels vs torch (log_softmax + transpose + F.ctc_loss), forward + backward."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import onebit_b200  # noqa: E402,F401
from onebit_b200 import ctc  # noqa: E402

B, T, V, L, blank = 64, 399, 5004, 64, 3
g = torch.Generator().manual_seed(0)
x = torch.randn(B, T, V, generator=g).cuda().requires_grad_(True)
targets = torch.randint(4, V, (B, L), generator=g).cuda()
in_lens, tgt_lens = torch.full((B,), T), torch.full((B,), L)
in_dev, tgt_dev = in_lens.cuda(), tgt_lens.cuda()


def ours():
    loss = ctc.ctc_loss(x, in_dev, targets, tgt_dev, blank)
    return loss, torch.autograd.grad(loss, x)[0]


def theirs():
    loss = F.ctc_loss(F.log_softmax(x, dim=-1).transpose(0, 1), targets, in_lens, tgt_lens, blank=blank, reduction="mean",
                      zero_infinity=True)
    return loss, torch.autograd.grad(loss, x)[0]


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


la, ga = ours()
lb, gb = theirs()
out = {"loss_ours": la.item(), "loss_torch": lb.item(), "grad_max_abs_diff": (ga - gb).abs().max().item(),
       "grad_max": gb.abs().max().item(), "ms_ours": round(timeit(ours), 3), "ms_torch": round(timeit(theirs), 3)}
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ours()
    torch.cuda.synchronize()
out["kernels_us"] = {e.key[:40]: round(e.device_time_total, 1) for e in prof.key_averages() if e.device_time_total > 0}
print(json.dumps(out))
