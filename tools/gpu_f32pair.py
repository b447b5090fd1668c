"""A/B of the CTA-pair variant of ob_gemm_f32 (ob_debug_set key 8) on a B200: same products with and without pairs,
accuracy against fp64, bitwise agreement between the two, and timings.  Appends one JSON line per case to
gpurun_out/f32pair.jsonl as it goes (a cut-off run keeps what it measured)."""
import json
import os
import sys
import time

t_start = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import onebit_b200 as ob  # noqa: E402,F401
from onebit_b200._cabi import lib  # noqa: E402
from onebit_b200.matmul import bmm_nt  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "f32pair.jsonl"), "a")
KEY_PAIR = 8


def emit(**kw):
    kw["t"] = round(time.time() - t_start, 1)
    LOG.write(json.dumps(kw) + "\n")
    LOG.flush()
    print(kw, flush=True)


def timeit(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


g = torch.Generator(device=dev).manual_seed(0)
R = lambda *s: torch.randn(*s, device=dev, generator=g)  # noqa: E731
emit(case="start", device=torch.cuda.get_device_name(0))

cases = [
    # name, a (.., M, K) view, b (.., N, K) view, bias
    ("NT small 512x256x256", lambda: (R(512, 256), R(256, 256), None)),
    ("NT ragged 1000x500x300 bias", lambda: (R(1000, 300), R(500, 300), R(500))),
    ("NN (B mn-major) 700x256x516", lambda: (R(700, 516), R(516, 256).t(), None)),
    ("TN (A,B mn-major, split-K) 5004x256x9000", lambda: (R(9000, 5004).t(), R(9000, 256).t(), None)),
    ("extra TK (A mn-major, B k-major) 3000x512x700", lambda: (R(700, 3000).t(), R(512, 700), None)),
    ("extra NT batch 2x3 600x300x200 bias", lambda: (R(2, 3, 600, 200), R(2, 3, 300, 200), R(300))),
    ("extra NT broadcast B over batch 2x3 600x300x200", lambda: (R(2, 3, 600, 200), R(1, 3, 300, 200), None)),
    ("extra NT tails 257x129x129", lambda: (R(257, 132)[:, :129], R(129, 132)[:, :129], None)),
    ("NT vocabulary fwd 25536x5004x256 bias", lambda: (R(25536, 256), R(5004, 256), R(5004))),
    ("NN vocabulary dx 25536x256x5004", lambda: (R(25536, 5004), R(5004, 256).t(), None)),
    ("TN vocabulary dW 5004x256x25536", lambda: (R(25536, 5004).t(), R(25536, 256).t(), None)),
    ("NT 1x1 conv 76608x512x256 bias", lambda: (R(76608, 256), R(512, 256), R(512))),
]
only = sys.argv[1:] or None
for name, make in cases:
    if only and not any(o in name for o in only):
        continue
    a, b, bias = make()
    lib.ob_debug_set(KEY_PAIR, 0)             # the script restores the library default (pairs on) only by exiting
    single = bmm_nt(a, b, bias=bias)
    torch.cuda.synchronize()
    lib.ob_debug_set(KEY_PAIR, 1)
    pair = bmm_nt(a, b, bias=bias)
    torch.cuda.synchronize()
    row = dict(case=name, bitwise_equal=bool(torch.equal(single, pair)),
               max_abs_diff=(single - pair).abs().max().item())
    if a.shape[-2] * b.shape[-2] <= 6_000_000 or "vocabulary fwd" in name:
        ref = a.double() @ b.double().transpose(-1, -2)
        if bias is not None:
            ref = ref + bias.double()
        scale = ref.abs().max()
        row["err_single"] = ((single.double() - ref).abs().max() / scale).item()
        row["err_pair"] = ((pair.double() - ref).abs().max() / scale).item()
        del ref
    lib.ob_debug_set(KEY_PAIR, 0)
    row["us_single"] = round(timeit(lambda: bmm_nt(a, b, bias=bias)), 1)
    lib.ob_debug_set(KEY_PAIR, 1)
    row["us_pair"] = round(timeit(lambda: bmm_nt(a, b, bias=bias)), 1)
    if "batch 2x3 600" in name and bias is not None:          # accumulation into an existing output (vector atomic adds)
        outs = []
        for on in (0, 1):
            lib.ob_debug_set(KEY_PAIR, on)
            acc = torch.ones(2, 3, 600, 300, device=dev)
            bmm_nt(a, b, out=acc, accumulate=True)
            outs.append(acc)
        torch.cuda.synchronize()
        row["accumulate_equal"] = bool(torch.equal(outs[0], outs[1]))
    lib.ob_debug_set(KEY_PAIR, 0)
    emit(**row)
    del a, b, bias, single, pair
emit(case="done")
