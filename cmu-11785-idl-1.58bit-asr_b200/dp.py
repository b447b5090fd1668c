"""Data-parallel gradient exchange for the co-training step: one process per GPU, NCCL all-reduce over NVLink.

The reference has no distributed code; the only exchange step of the path is the mean of
the gradients.  Every parameter's AccumulateGrad fires exactly once per ``loss.backward()`` even though the latent
weights are used by three passes, so a post-accumulate-grad hook per parameter is enough: gradients are copied into
flat fp32 buckets (~25 MB, filled in reverse-autograd order) and each full bucket is all-reduced asynchronously on
a side stream while the rest of backward runs (one multi-tensor copy per completed bucket).  ``finish()`` joins the side
stream, averages, and makes every ``.grad`` a view of its bucket before ``clip_grad_norm_`` needs the global gradients
(train.py:117); gradients must be reset with ``zero_grad(set_to_none=True)`` between steps (``train_step`` does).

Replica consistency: the constructor broadcasts every parameter (and the given buffers) from rank 0, as DDP does, so a
rank that was seeded or restored differently cannot diverge silently.  Buckets are all-reduced strictly in bucket-index
order on every rank - a bucket that completes early waits for its predecessors, the rest are flushed in order by
``finish()`` - so the sequence of NCCL calls is identical on all ranks even when the set of parameters that received a
gradient differs between them.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


class GradAllReducer:
    def __init__(self, params, bucket_bytes: int = 25 << 20, process_group=None, buffers=(), broadcast: bool = True):
        params = list(params)
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        if broadcast and self.world > 1:
            with torch.no_grad():
                for t in list(params) + list(buffers):
                    dist.broadcast(t.data, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                                   group=process_group)
        self.cuda = bool(self.params) and self.params[0].is_cuda
        self.stream = torch.cuda.Stream() if self.cuda else None
        # buckets in reverse parameter order ~ the order gradients become ready
        self.buckets = []                 # dicts: params, offsets, flat buffer, numel
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * 4
            if cur and cur_bytes + nbytes > bucket_bytes:
                self._close_bucket(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self._close_bucket(cur)
        self._where = {}
        for bi, b in enumerate(self.buckets):
            for p, off in zip(b["params"], b["offsets"]):
                self._where[p] = (bi, off)
        self._handles = []
        self.enabled = True               # False: gradients only accumulate locally (all but the last micro-batch)
        self.timing = False               # True: CUDA events around finish() -> exposed_ms()
        self._events = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self._reset()

    def _close_bucket(self, plist):
        offsets, n = [], 0
        for p in plist:
            offsets.append(n)
            n += p.numel()
        flat = torch.zeros(n, dtype=torch.float32, device=plist[0].device)
        views = [flat[off:off + p.numel()] for p, off in zip(plist, offsets)]
        self.buckets.append({"params": plist, "offsets": offsets, "flat": flat, "numel": n, "views": views})

    def _reset(self):
        self._pending = [len(b["params"]) for b in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._ready = [False] * len(self.buckets)
        self._next = 0                                   # next bucket index to all-reduce (same order on every rank)
        self._handles = []

    def _on_grad(self, p: torch.nn.Parameter):
        if self.world == 1 or not self.enabled:
            return
        bi, _ = self._where[p]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:                       # the bucket is complete: one multi-tensor copy, then the all-reduce
            b = self.buckets[bi]
            torch._foreach_copy_(b["views"], [q.grad.reshape(-1) for q in b["params"]])
            self._ready[bi] = True
            while self._next < len(self.buckets) and self._ready[self._next]:
                self._launch(self._next)
                self._next += 1

    def _launch(self, bi: int):
        b = self.buckets[bi]
        self._launched[bi] = True
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                h = dist.all_reduce(b["flat"], group=self.group, async_op=True)
        else:
            h = dist.all_reduce(b["flat"], group=self.group, async_op=True)
        self._handles.append(h)

    def finish(self):
        """Wait for every bucket, divide by the world size and write the averaged gradients back."""
        if self.world == 1:
            return
        ev = None
        if self.timing and self.cuda:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        for bi in range(self._next, len(self.buckets)):  # the rest, in index order
            b = self.buckets[bi]
            if not self._ready[bi]:                      # holds parameters without a gradient this step (e.g. alpha at 32 bit)
                for p, v in zip(b["params"], b["views"]):
                    if p.grad is None:
                        v.zero_()
                    else:
                        v.copy_(p.grad.reshape(-1))
            self._launch(bi)
        self._next = len(self.buckets)
        for h in self._handles:
            h.wait()
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.stream)
        inv = 1.0 / self.world
        for b in self.buckets:
            b["flat"].mul_(inv)
            for p, v in zip(b["params"], b["views"]):
                # the averaged gradient lives in the bucket: .grad becomes a view of it (no copy back).  The next backward
                # starts from zero_grad(set_to_none=True) - as train_step does - and refills the bucket from fresh gradients.
                p.grad = v.view_as(p)
        if ev is not None:
            ev[1].record()
            self._events.append(ev)
        self._reset()

    def exposed_ms(self):
        """Mean device time between the end of backward and the averaged gradients being ready (the part of the exchange
        that did NOT overlap with backward), over the ``finish()`` calls made while ``timing`` was set; None if none."""
        if not self._events:
            return None
        torch.cuda.synchronize()
        ms = [a.elapsed_time(b) for a, b in self._events]
        self._events = []
        return sum(ms) / len(ms)

    def remove(self):
        for h in self._hooks:
            h.remove()
