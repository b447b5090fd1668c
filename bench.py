#!/usr/bin/env python
"""bench.py - the measurement contract ("Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train|gemm] [--impl ours|reference]

Workloads (one "step" = one pass of the hot path over one batch of synthetic input):

  train  (default) BASELINE.json configs[2]/[3]: the Conformer co-training step of the reference
         (onebit_asr/train.py:83-120: 2-bit + 1-bit + stochastic-precision passes, CTC + attention + KL losses,
         one backward, clip, AdamW) with every routed projection on the B200 layer; batch 64 x 1600 frames x 80 mel
         PER GPU (weak scaling: global batch 512 at 8 GPUs), NCCL all-reduce of the gradients overlapped with
         backward.  metric = audio-seconds per second (1 frame = 10 ms).  At N = 1 the line also carries the
         BitLinear int8 GEMM microbenchmark of configs[1] ("bitlinear_gemm").
  infer  BASELINE.json configs[4]: packed 2-bit weights (PackedQuantizedLinear), batches of 1..256 utterances x 1000
         frames, encoder forward at precision 2 + CTC head + greedy CTC decode on the device; metric = audio-s/s.
  gemm   BASELINE.json configs[1] alone: act-quant + ternary x int8 tcgen05 GEMM at M = 65536, K = N = 2048;
         metric = BitLinear int8 TOPS.  `--sweep` adds the whole K/N sweep.

`--impl reference` times the UNMODIFIED reference (baseline/_ref: its modules byte-compiled by oracle/build_ref.py) on the host
cores: train.py's own run_epoch on BASELINE configs[0] (batch 4 x 1000 frames, default dims) - a bounded sample of the step the
GPU arm times; that process never imports the product package.  The GPU arm's `cpu_baseline` is the same thing run in a
child process.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FRAME_S = 0.010          # kaldi fbank frame shift (src/data/dataset.py:124-128): 1600 frames = 16 s of audio
TRAIN = dict(batch=64, frames=1600, mel=80, vocab=5004, tokens=64)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--workload", default="train", choices=["train", "gemm", "infer"])
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--batch", type=int, default=TRAIN["batch"], help="utterances per GPU (train)")
    p.add_argument("--frames", type=int, default=TRAIN["frames"])
    p.add_argument("--dropout", type=float, default=0.1)
    p.add_argument("--tokens", type=int, default=65536, help="GEMM microbench M")
    p.add_argument("--in-features", type=int, default=2048)
    p.add_argument("--out-features", type=int, default=2048)
    p.add_argument("--bitwidth", type=int, default=2)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--global-batch", type=int, default=0,
                   help="train: fix the GLOBAL batch (BASELINE configs[3]: 512) instead of the per-GPU batch: each rank takes "
                        "global/N utterances, in micro-batches of at most --max-micro-batch with gradient accumulation "
                        "(strong scaling); 0 = weak scaling with --batch utterances per GPU")
    p.add_argument("--max-micro-batch", type=int, default=128)
    p.add_argument("--no-small-m", action="store_true", help="skip the small-batch (M <= 256) table inside the train line")
    p.add_argument("--no-sweep-summary", action="store_true", help="skip the configs[1] K/N sweep summary inside the train line")
    p.add_argument("--no-gemm", action="store_true", help="skip the GEMM microbench inside the train line")
    p.add_argument("--no-share-frontend", dest="share_frontend", action="store_false",
                   help="recompute the (bitwidth-independent, dropout-free) conv subsampling in each of the three passes, as the "
                        "reference does; by default it is computed once per step and shared - identical loss and gradients")
    p.add_argument("--sweep", action="store_true", help="gemm workload: also run the configs[1] K/N sweep")
    p.add_argument("--torch-nonrouted", default="",
                   help="A/B switch: comma list of {attn,conv,linear,tail,decoder,frontend,ctc} to run on torch's own fp32 kernels instead of the "
                        "library's (sets OB_TORCH_NONROUTED before the package is imported)")
    p.add_argument("--allocator-headroom-gib", type=float, default=6.0,
                   help="train: free block kept in torch's caching allocator after the first warm-up step so that a shifted request "
                        "size never reaches cudaMalloc inside the timed steps (training.reserve_allocator_headroom); 0 disables")
    p.add_argument("--no-stack-passes", action="store_true",
                   help="run the three co-training encoder passes one after the other (as train.py does) instead of side by "
                        "side on one stacked batch - same loss and gradients (tests), 3x the launches")
    p.add_argument("--foreach-adamw", action="store_true",
                   help="use torch's default multi-tensor AdamW instead of the fused single-kernel one (same update rule)")
    p.add_argument("--tf32-nonrouted", action="store_true",
                   help="NOT the reference's numerics: let the non-routed fp32 matmuls (attention, vocabulary projections) use "
                        "TF32 tensor cores; reported with this flag in `config`, never the default")
    return p.parse_args()


# ------------------------------------------------------------------------------------------ helpers
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()

        def num(s):
            try:
                return float(s)
            except ValueError:
                return None
        sm = sorted(v for v in (num(r[0]) for r in self.rows if r) if v is not None)
        mx = [v for v in (num(r[1]) for r in self.rows if len(r) > 1) if v is not None]
        pw = [v for v in (num(r[2]) for r in self.rows if len(r) > 2) if v is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names)
                   if any(len(r) > 3 + i and r[3 + i].lower() == "active" for r in self.rows)]
        return {"sm_mhz": int(sm[len(sm) // 2]) if sm else None, "sm_max_mhz": int(max(mx)) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(sm)}


def dist_setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def timed_region(world, fn, steps):
    """barrier + sync, K steps between CUDA events on the current stream, max over ranks; returns ms."""
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = t.item()
    return ms


def ncu_traffic(key):
    """DRAM bytes per launch measured by ncu for a named kernel/shape (profiles/ncu_traffic.json), else None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[key]["bytes"]
    except Exception:
        return None


def int8_peak(peaks):
    return 2.0 * peaks["bf16_tflops"], (f"2 x {peaks['source']} bf16 burst {peaks['bf16_tflops']} TFLOP/s (int8 dense is nominally "
                                        "2x bf16; MEASURED_PEAKS.json has no int8 figure)")


# ------------------------------------------------------------------------------------------ GEMM microbench
def gemm_microbench(args, steps, sweep=False):
    """configs[1]: forward hot path (activation quantiser + ternary x int8 GEMM) on device-resident operands."""
    import onebit_b200 as ob
    from onebit_b200 import quant as obq
    peaks = load_peaks()
    dev = torch.device("cuda", torch.cuda.current_device())
    M, K, N, bw = args.tokens, args.in_features, args.out_features, args.bitwidth
    torch.manual_seed(0)
    layer = ob.QuantizedLinear(K, N).to(dev)
    xb = torch.randn(M, K, device=dev, generator=torch.Generator(device=dev).manual_seed(1234)).to(torch.bfloat16)
    packed, _ = layer.packed_weight(bw)

    def step():
        q, s = ob.act_quant_int8(xb)
        return obq.gemm_fwd(q, s, packed, layer.alpha, layer.bias, N, torch.bfloat16)
    for _ in range(3):
        step()
    step_ms = timed_region(1, step, steps) / steps
    q, s = ob.act_quant_int8(xb)

    def gemm_only():
        obq.gemm_fwd(q, s, packed, layer.alpha, layer.bias, N, torch.bfloat16)
    for _ in range(3):
        gemm_only()
    gemm_ms = timed_region(1, gemm_only, steps) / steps

    def act_only():
        ob.act_quant_int8(xb)
    for _ in range(3):
        act_only()
    act_ms = timed_region(1, act_only, steps) / steps
    ops = 2.0 * M * N * K
    gemm_bytes = M * K + N * K / 4 + 2.0 * M * N + 4 * M + 4 * N
    peak, peak_src = int8_peak(peaks)
    tops = ops / (gemm_ms * 1e-3) / 1e12
    act_gbs = (M * K * 2 + M * K + 4 * M) / (act_ms * 1e-3) / 1e9
    out = {"shape": {"M": M, "K": K, "N": N, "bitwidth": bw, "in": "bf16", "out": "bf16"},
           "step_ms": round(step_ms, 4), "step_tops": round(ops / (step_ms * 1e-3) / 1e12, 1),
           "roofline": {"kernel": "gemm_expand_kernel<kFwdI8, 256, 5, bf16, cta_group::2> (ternary x int8, tcgen05 kind::i8)",
                        "bound": "tensor", "achieved": round(tops, 1), "peak": round(peak, 1), "unit": "TFLOP/s",
                        "frac": round(tops / peak, 4),
                        "traffic": ncu_traffic(f"gemm_fwd_{M}x{K}x{N}_bf16"), "peak_source": peak_src,
                        "ms": round(gemm_ms, 4), "arith_intensity_op_per_byte": round(ops / gemm_bytes, 1),
                        "algorithmic_bytes": int(gemm_bytes)},
           "act_quant": {"kernel": "act_quant_reg_kernel<bf16,16>", "bound": "hbm", "achieved": round(act_gbs, 1),
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(act_gbs / peaks["hbm_gbs"], 4),
                         "ms": round(act_ms, 4)}}
    try:   # cuBLASLt int8 (a library GEMM, not the product) as a cross-check of the int8 denominator
        a8 = torch.randint(-127, 127, (8192, 8192), device=dev, dtype=torch.int8)
        b8 = torch.randint(-127, 127, (8192, 8192), device=dev, dtype=torch.int8).t()
        for _ in range(3):
            torch._int_mm(a8, b8)
        t = timed_region(1, lambda: torch._int_mm(a8, b8), 10) / 10
        out["roofline"]["cublaslt_int_mm_8192_tops"] = round(2 * 8192 ** 3 / (t * 1e-3) / 1e12, 1)
    except Exception as e:  # pragma: no cover
        out["roofline"]["cublaslt_int_mm_8192_tops"] = f"unavailable: {type(e).__name__}"
    if sweep:
        rows = []
        for Ms in (4096, 65536):
            for Ks in (256, 512, 1024, 2048):
                for Ns in (256, 512, 1024, 2048):
                    torch.manual_seed(0)
                    l2 = ob.QuantizedLinear(Ks, Ns).to(dev)
                    pk, _ = l2.packed_weight(bw)
                    nbuf = max(1, int(400e6 // (Ms * Ks + 2 * Ms * Ns)))      # rotate operands: working set > L2
                    qs = [ob.act_quant_int8(torch.randn(Ms, Ks, device=dev, dtype=torch.bfloat16)) for _ in range(nbuf)]
                    g = torch.cuda.CUDAGraph()                                # graph replay: device time, no launch gaps
                    ys = [torch.empty(Ms, Ns, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        obq.gemm_fwd(qs[0][0], qs[0][1], pk, l2.alpha, l2.bias, Ns, torch.bfloat16)
                        with torch.cuda.graph(g, stream=side):
                            for i in range(20):
                                qq, ss = qs[i % nbuf]
                                obq.gemm_fwd(qq, ss, pk, l2.alpha, l2.bias, Ns, torch.bfloat16)
                    torch.cuda.current_stream().wait_stream(side)
                    g.replay()
                    t = timed_region(1, g.replay, 3) / 3 / 20
                    by = Ms * Ks + Ns * Ks / 4 + 2.0 * Ms * Ns + 4 * Ms + 4 * Ns
                    rows.append({"M": Ms, "K": Ks, "N": Ns, "us": round(t * 1e3, 2),
                                 "tops": round(2.0 * Ms * Ns * Ks / (t * 1e-3) / 1e12, 1),
                                 "gbs": round(by / (t * 1e-3) / 1e9, 1)})
                    del ys
        out["sweep"] = rows
    return out


def _graph_time_us(fn_of_i, n_per_graph, reps=5):
    """Device time per call of fn_of_i(i), i = 0..n_per_graph-1, replayed from a CUDA graph (no host launch gaps)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn_of_i(0)
        with torch.cuda.graph(g, stream=side):
            for i in range(n_per_graph):
                fn_of_i(i)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    return timed_region(1, g.replay, reps) / reps / n_per_graph * 1e3


def _device_time_us(fn_of_i, n_per_graph):
    """Device time per call: CUDA-graph replay; if the capture fails for any reason, eager launches between events (which also
    charge the host-side enqueue to the kernel) so that the bench line survives."""
    try:
        return _graph_time_us(fn_of_i, n_per_graph)
    except Exception:  # noqa: BLE001
        torch.cuda.synchronize()
        k = [0]

        def call():
            fn_of_i(k[0])
            k[0] += 1
        return timed_region(1, call, n_per_graph) / n_per_graph * 1e3


def sweep_summary(rows, peaks):
    """configs[1] K/N sweep condensed: every shape against the roof that binds it (int8 ridge = int8 peak / measured HBM)."""
    peak_i8, _ = int8_peak(peaks)
    ridge = peak_i8 * 1e12 / (peaks["hbm_gbs"] * 1e9)
    out = {"ridge_op_per_byte": round(ridge, 1), "int8_peak_tops": round(peak_i8, 1), "hbm_peak_gbs": peaks["hbm_gbs"], "by_M": {}}
    for Ms in sorted({r["M"] for r in rows}):
        sel = [r for r in rows if r["M"] == Ms]
        for r in sel:
            ops = 2.0 * r["M"] * r["N"] * r["K"]
            by = r["M"] * r["K"] + r["N"] * r["K"] / 4 + 2.0 * r["M"] * r["N"] + 4 * r["M"] + 4 * r["N"]
            r["bound"] = "tensor" if ops / by >= ridge else "hbm"
            r["frac"] = round(r["tops"] / peak_i8 if r["bound"] == "tensor" else r["gbs"] / peaks["hbm_gbs"], 4)
        reg = {}
        for bound in ("tensor", "hbm"):
            fr = sorted(r["frac"] for r in sel if r["bound"] == bound)
            if fr:
                reg[bound] = {"shapes": len(fr), "min_frac": fr[0], "median_frac": fr[len(fr) // 2], "max_frac": fr[-1]}
        out["by_M"][str(Ms)] = {"regimes": reg,
                                "rows": [{k: r[k] for k in ("K", "N", "us", "tops", "gbs", "bound", "frac")} for r in sel]}
    return out


def small_batch_table(peaks):
    """north_star: "at small batch, the packed-weight GEMV-like regime is also reported as achieved GB/s".  M token rows in
    {1, 8, 64, 256} x the model's three routed shapes (+ 2048^2); small M runs the weight-streaming DP4A kernel
    (csrc/ob_gemv.cu) where it beats the 128-row tcgen05 tile (M * K <= 32 K codes), both timings are listed.  Device time per launch from a CUDA-graph replay over 8 rotating weight/activation sets (the layers of a
    model are distinct: 8 x 64 KB..1 MB stays in L2, as it does in a real 108-layer forward at this batch); bytes = packed W
    (N K / 4) + int8 activations (M K) + fp32 output (4 M N) + scales and bias."""
    import onebit_b200 as ob
    from onebit_b200 import _cabi, quant as obq
    dev = torch.device("cuda", torch.cuda.current_device())
    rows = []
    for K, N in ((256, 256), (256, 1024), (1024, 256), (2048, 2048)):
        torch.manual_seed(0)
        layers = [ob.QuantizedLinear(K, N).to(dev) for _ in range(8)]
        pks = [l.packed_weight(2)[0] for l in layers]
        for M in (1, 8, 64, 256):
            xs = [torch.randn(M, K, device=dev) for _ in range(8)]
            qs = [ob.act_quant_int8(x) for x in xs]

            def gemm(i):
                j = i % 8
                obq.gemm_fwd(qs[j][0], qs[j][1], pks[j], layers[j].alpha, layers[j].bias, N, torch.float32)

            def layer_fwd(i):
                j = i % 8
                q, sc = ob.act_quant_int8(xs[j])
                obq.gemm_fwd(q, sc, pks[j], layers[j].alpha, layers[j].bias, N, torch.float32)
            us = _graph_time_us(gemm, 64)
            us_layer = _graph_time_us(layer_fwd, 64)
            us_tc = us_dp4a = None
            if M <= 64:                                        # the same call forced onto either kernel, for comparison
                try:
                    _cabi.lib.ob_debug_set(_cabi.DBG_SMALL_M, 1)
                    us_tc = round(_graph_time_us(gemm, 64), 2)
                    _cabi.lib.ob_debug_set(_cabi.DBG_SMALL_M, 2)
                    us_dp4a = round(_graph_time_us(gemm, 64), 2)
                finally:
                    _cabi.lib.ob_debug_set(_cabi.DBG_SMALL_M, 0)
            by = N * K / 4 + M * K + 4.0 * M * N + 4 * M + 4 * N
            dp4a = us_dp4a is not None and abs(us - us_dp4a) <= abs(us - us_tc)      # which one the dispatcher picked
            rows.append({"M": M, "K": K, "N": N, "kernel": "gemv_tern_i8 (DP4A)" if dp4a else "gemm_expand (tcgen05)",
                         "us": round(us, 2), "gbs": round(by / us / 1e3, 1), "frac_hbm": round(by / us / 1e3 / peaks["hbm_gbs"], 4),
                         "weight_gbs": round(N * K / 4 / us / 1e3, 1), "layer_fwd_us": round(us_layer, 2),
                         "tcgen05_tile_us": us_tc, "dp4a_us": us_dp4a, "algorithmic_bytes": int(by)})
    return {"note": "packed-weight GEMV-like regime; at these sizes (16 KB - 1 MB of weights) a launch is latency-bound: "
                    "the floor is the ~2 us launch + one dependent HBM/L2 round trip, not bandwidth", "rows": rows}


# ------------------------------------------------------------------------------------------ train workload
def train_workload_config(args, world):
    """The workload both arms are quoted on (the reference arm runs a bounded sample of it, stated in its `cpu_baseline.sample`)."""
    B = args.global_batch // world if args.global_batch else args.batch
    return {"workload": "Conformer BitLinear co-training step (BASELINE configs[2]; configs[3] = global batch 512): 12 blocks, "
                        "d_model 256, d_ff 1024, 4 heads, V=5004, 3 passes (2-bit, 1-bit, stochastic precision) + "
                        "CTC/attention/KL losses + clip + AdamW",
            "batch_per_gpu": B, "global_batch": B * world, "frames": args.frames, "mel": TRAIN["mel"], "dropout": args.dropout,
            "l2": "no explicit flush: the activations one step streams (tens of GB) exceed the 126 MB L2 many times over",
            "parallelism": f"dp{world}"}


def make_batch(B, T, seed, device=None, pinned=False):
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(B, T, TRAIN["mel"], generator=g)
    batch = {"feats": feats.pin_memory() if pinned else feats,
             "feat_lens": torch.full((B,), T, dtype=torch.long),
             "tokens": torch.randint(4, TRAIN["vocab"], (B, TRAIN["tokens"]), generator=g),
             "token_lens": torch.full((B,), TRAIN["tokens"], dtype=torch.long)}
    if device is not None:
        host = batch
        batch = {k: v.to(device) for k, v in host.items()}
        batch["feat_lens_cpu"], batch["token_lens_cpu"] = host["feat_lens"], host["token_lens"]
    return batch


def hot_kernel_rooflines(peaks, M):
    """Each kernel of the layer at the model's widest routed shape (256 -> 1024, M tokens), timed alone with CUDA
    events over a rotating operand set that exceeds L2; algorithmic bytes per DESIGN.md."""
    import onebit_b200 as ob
    from onebit_b200 import _cabi, quant as obq
    lib = _cabi.lib
    dev = torch.device("cuda", torch.cuda.current_device())
    K, N = 256, 1024
    torch.manual_seed(0)
    layer = ob.QuantizedLinear(K, N).to(dev)
    pk, pkt = layer.packed_weight(2)
    nbuf = 4
    xs = [torch.randn(M, K, device=dev) for _ in range(nbuf)]
    gys = [torch.randn(M, N, device=dev) for _ in range(nbuf)]
    qs = [ob.act_quant_int8(x) for x in xs]
    ys = [torch.empty(M, N, device=dev) for _ in range(nbuf)]
    dys = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
    dxs = [torch.empty(M, K, device=dev) for _ in range(nbuf)]
    colsum = torch.empty(lib.ob_bwd_colsum_blocks(M), N, device=dev)
    gw, ga, gb = torch.empty(N, K, device=dev), torch.empty((), device=dev), torch.empty(N, device=dev)
    nbytes = lib.ob_bwd_dw_workspace_bytes(M, N, K)
    ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    cur = lambda: torch.cuda.current_stream().cuda_stream        # noqa: E731  (looked up per call: the timing graphs capture on a side stream)
    a = layer.alpha
    i = [0]

    def nxt():
        i[0] = (i[0] + 1) % nbuf
        return i[0]
    fns = {
        "act_quant_i8": (lambda j: lib.ob_act_quant_i8(xs[j].data_ptr(), 0, M, K, qs[j][0].data_ptr(), qs[j][1].data_ptr(), cur()),
                         4 * M * K + M * K + 4 * M, 0.0),
        "gemm_fwd": (lambda j: lib.ob_gemm_tern_i8_fwd(qs[j][0].data_ptr(), qs[j][1].data_ptr(), pk.data_ptr(), a.data_ptr(), 1,
                                                       layer.bias.data_ptr(), M, N, K, ys[j].data_ptr(), 0, cur()),
                     M * K + N * K / 4 + 4.0 * M * N + 4 * M + 4 * N, 2.0 * M * N * K),
        # round 2: grad_W converts the int8 codes in shared memory, so the prep pass no longer reads q or writes a bf16 copy of it
        "bwd_prep": (lambda j: lib.ob_bwd_prep(gys[j].data_ptr(), 0, qs[j][1].data_ptr(), qs[j][0].data_ptr(), M, N, K,
                                               dys[j].data_ptr(), None, colsum.data_ptr(), cur()),
                     6.0 * M * N + 4 * M, 0.0),
        "bwd_dx": (lambda j: lib.ob_bwd_dx(dys[j].data_ptr(), qs[j][1].data_ptr(), pkt.data_ptr(), a.data_ptr(), 1, M, N, K,
                                           dxs[j].data_ptr(), 0, cur()),
                   2.0 * M * N + N * K / 4 + 4.0 * M * K + 4 * M, 2.0 * M * N * K),
        "bwd_dw": (lambda j: lib.ob_bwd_dw_q8(dys[j].data_ptr(), qs[j][0].data_ptr(), colsum.data_ptr(), layer.weight.data_ptr(),
                                              a.data_ptr(), 1, 2, M, N, K, gw.data_ptr(), ga.data_ptr(), gb.data_ptr(),
                                              ws.data_ptr(), nbytes, cur()),
                   2.0 * M * N + 1.0 * M * K + 8.0 * N * K, 2.0 * M * N * K),
    }
    # kernels around the layer (same token count): FFN mid-section, LayerNorm, attention chain, weight quantiser
    from onebit_b200.attention import rel_attention_probs  # noqa: F401
    h_mid = [torch.randn(M, N, device=dev) for _ in range(2)]
    q_mid = torch.empty(M, N, device=dev, dtype=torch.int8)
    s_mid = torch.empty(M, device=dev)
    gh = torch.empty(M, N, device=dev)
    ln_w, ln_b = torch.ones(K, device=dev), torch.zeros(K, device=dev)
    ln_y, ln_stats = torch.empty(M, K, device=dev), torch.empty(2, M, device=dev)
    ln_dx, ln_dp = torch.empty(M, K, device=dev), torch.empty(2, K, device=dev)
    ln_ws = torch.empty(lib.ob_layernorm_bwd_workspace_bytes(K), device=dev, dtype=torch.uint8)
    Bq, Hh, Tt = max(1, M // 399), 4, 399
    att = [torch.randn(Bq, Hh, Tt, Tt, device=dev) for _ in range(4)]
    drop_thr = int(round(0.1 * 2 ** 16))              # in-kernel Philox dropout, p = 0.1 (no stored mask)
    att_mask = torch.ones(Bq, Tt, Tt, device=dev, dtype=torch.bool)
    att_y, att_o = torch.empty_like(att[0]), torch.empty_like(att[0])
    nat = Bq * Hh * Tt * Tt
    pk2, pkt2 = torch.empty_like(pk), torch.empty_like(pkt)
    fns.update({
        "swish_drop_quant": (lambda j: lib.ob_swish_drop_quant(h_mid[j % 2].data_ptr(), None, 1.0 / 0.9, 1234, 4 * j, drop_thr, M, N,
                                                                q_mid.data_ptr(), s_mid.data_ptr(), cur()), 5.0 * M * N + 4 * M, 0.0),
        "swish_drop_bwd": (lambda j: lib.ob_swish_drop_bwd(gys[j].data_ptr(), h_mid[j % 2].data_ptr(), None, 1.0 / 0.9, 1234, 4 * j,
                                                            drop_thr, M * N, gh.data_ptr(), cur()), 12.0 * M * N, 0.0),
        "layernorm_fwd": (lambda j: lib.ob_layernorm_fwd(xs[j].data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), 1e-5, M, K, ln_y.data_ptr(),
                                                         ln_stats[0].data_ptr(), ln_stats[1].data_ptr(), cur()), 8.0 * M * K + 8 * M, 0.0),
        "layernorm_bwd": (lambda j: lib.ob_layernorm_bwd(dxs[j].data_ptr(), xs[j].data_ptr(), ln_stats[0].data_ptr(), ln_stats[1].data_ptr(),
                                                         ln_w.data_ptr(), M, K, ln_dx.data_ptr(), ln_dp[0].data_ptr(), ln_dp[1].data_ptr(),
                                                         ln_ws.data_ptr(), cur()), 12.0 * M * K + 8 * M, 0.0),
        "relattn_softmax_fwd": (lambda j: lib.ob_relattn_softmax_fwd(att[j % 2].data_ptr(), att[2 + j % 2].data_ptr(), att_mask.data_ptr(),
                                                                     None, 1.0 / 0.9, 1234, 4 * j, drop_thr, 0.125, Bq, Hh, Tt, Tt,
                                                                     att_y.data_ptr(), att_o.data_ptr(), cur()), 16.0 * nat, 0.0),
        "relattn_softmax_bwd": (lambda j: lib.ob_relattn_softmax_bwd(att[j % 2].data_ptr(), att_y.data_ptr(), None, 1.0 / 0.9, 1234, 4 * j,
                                                                     drop_thr, 0.125, Bq, Hh, Tt, Tt, att[2].data_ptr(), att[3].data_ptr(), cur()),
                                16.0 * nat, 0.0),
        "weight_quant_pack": (lambda j: lib.ob_weight_quant_pack(layer.weight.data_ptr(), a.data_ptr(), 1, N, K, 2, pk2.data_ptr(),
                                                                 pkt2.data_ptr(), cur()), 4.5 * N * K, 0.0),
    })
    # fp32 tensor-core GEMM (3 x tf32 split) at the shapes the model uses it on, and the convolution-module kernels
    from onebit_b200.matmul import bmm_nt
    dh = 64
    Tp = (Tt + 3) // 4 * 4
    qh = [torch.randn(Bq, Tt, Hh * dh, device=dev) for _ in range(2)]
    heads = lambda t: t.view(Bq, Tt, Hh, dh).permute(0, 2, 1, 3)  # noqa: E731
    sc = [torch.empty(Bq, Hh, Tt, Tp, device=dev)[..., :Tt] for _ in range(2)]
    pr = [torch.rand(Bq, Hh, Tt, Tp, device=dev)[..., :Tt] for _ in range(2)]
    mix = torch.empty(Bq, Tt, Hh * dh, device=dev)
    V = TRAIN["vocab"]
    w_voc, b_voc = torch.randn(V, K, device=dev) * 0.05, torch.zeros(V, device=dev)
    logits = torch.empty(M, V, device=dev)
    gvoc = torch.empty(V, K, device=dev)
    w_pw, y_pw = torch.randn(2 * K, K, device=dev) * 0.05, torch.empty(M, 2 * K, device=dev)
    nsc = float(Bq * Hh * Tt * Tt)
    fns.update({
        "f32gemm_scores_qk": (lambda j: bmm_nt(heads(qh[j % 2]), heads(qh[1 - j % 2]), out=sc[j % 2]),
                              4.0 * nsc + 8.0 * Bq * Tt * Hh * dh, 2.0 * nsc * dh),
        "f32gemm_probs_v": (lambda j: bmm_nt(pr[j % 2], heads(qh[j % 2]).transpose(-1, -2), out=heads(mix)),
                            4.0 * nsc + 8.0 * Bq * Tt * Hh * dh, 2.0 * nsc * dh),
        "f32gemm_probsT_do": (lambda j: bmm_nt(pr[j % 2].transpose(-1, -2), heads(qh[j % 2]).transpose(-1, -2), out=heads(mix)),
                              4.0 * nsc + 8.0 * Bq * Tt * Hh * dh, 2.0 * nsc * dh),
        "f32gemm_vocab_fwd": (lambda j: bmm_nt(xs[j], w_voc, out=logits, bias=b_voc),
                              4.0 * M * K + 4.0 * V * K + 4.0 * M * V, 2.0 * M * V * K, "tensor"),
        "f32gemm_vocab_dw": (lambda j: bmm_nt(logits.t(), xs[j].t(), out=gvoc),
                             4.0 * M * K + 4.0 * V * K + 4.0 * M * V, 2.0 * M * V * K, "tensor"),
        "f32gemm_pw1_fwd": (lambda j: bmm_nt(xs[j], w_pw, out=y_pw), 4.0 * M * K + 8.0 * K * K + 8.0 * M * K, 4.0 * M * K * K),
    })
    Bc, Tc = Bq, Tt
    Mc = Bc * Tc
    cv_a = [torch.randn(Bc, Tc, 2 * K, device=dev) for _ in range(2)]
    cv_w, cv_b = torch.randn(K, 31, device=dev) * 0.1, torch.zeros(K, device=dev)
    cv_d = [torch.randn(Bc, Tc, K, device=dev) for _ in range(2)]
    cv_s, cv_gd, cv_ga = torch.empty(Bc, Tc, K, device=dev), torch.empty(Bc, Tc, K, device=dev), torch.empty(Bc, Tc, 2 * K, device=dev)
    cv_stats, cv_ggb = torch.zeros(2, K, device=dev), torch.empty(2, K, device=dev)
    cv_stats[1].fill_(1.0)
    cv_gw, cv_gb = torch.empty(K, 31, device=dev), torch.empty(K, device=dev)
    cv_ws = torch.empty(lib.ob_convmod_workspace_bytes(Bc, Tc, K), device=dev, dtype=torch.uint8)
    fns.update({
        "glu_dwconv_bn_fwd": (lambda j: lib.ob_glu_dwconv_bn_fwd(cv_a[j % 2].data_ptr(), cv_w.data_ptr(), cv_b.data_ptr(), Bc, Tc, K, 31,
                                                                 1e-5, 1, cv_s.data_ptr(), cv_stats[0].data_ptr(), cv_stats[1].data_ptr(),
                                                                 cv_ws.data_ptr(), cur()), 12.0 * Mc * K, 0.0),
        "bn_swish_fwd": (lambda j: lib.ob_bn_swish_fwd(cv_d[j % 2].data_ptr(), cv_stats[0].data_ptr(), cv_stats[1].data_ptr(), ln_w.data_ptr(),
                                                       ln_b.data_ptr(), Mc, K, 1, cv_s.data_ptr(), cur()), 8.0 * Mc * K, 0.0),
        "bn_swish_bwd": (lambda j: lib.ob_bn_swish_bwd(cv_d[1 - j % 2].data_ptr(), cv_d[j % 2].data_ptr(), cv_stats[0].data_ptr(),
                                                       cv_stats[1].data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), Mc, K, 1, cv_gd.data_ptr(),
                                                       cv_ggb.data_ptr(), cv_ws.data_ptr(), cur()), 20.0 * Mc * K, 0.0),
        "glu_dwconv_bwd": (lambda j: lib.ob_glu_dwconv_bwd(cv_d[j % 2].data_ptr(), cv_a[j % 2].data_ptr(), cv_w.data_ptr(), Bc, Tc, K, 31,
                                                           cv_ga.data_ptr(), cv_gw.data_ptr(), cv_gb.data_ptr(), cv_ws.data_ptr(), cur()),
                           32.0 * Mc * K, 0.0),
    })
    out = {}
    for name, spec in fns.items():
        fn, nbytes_alg, flops = spec[:3]
        for _ in range(3):
            fn(nxt())
        # device time per call from a CUDA-graph replay of 20 calls over the rotating operand sets: several entries are more than
        # one launch (grad_W = GEMM + memset + finaliser) and their host-side enqueue (tensor-map encodes, ctypes) would otherwise
        # be what a 30 us kernel is timed by
        ms = _device_time_us(lambda i: fn(i % nbuf), 20) / 1e3
        gbs = nbytes_alg / (ms * 1e-3) / 1e9
        out[name] = {"ms": round(ms, 4), "bound": "hbm", "achieved": round(gbs, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": round(gbs / peaks["hbm_gbs"], 4), "algorithmic_bytes": int(nbytes_alg),
                     "tflops": round(flops / (ms * 1e-3) / 1e12, 1) if flops else None}
        if len(spec) > 3 and spec[3] == "tensor":
            # compute-bound shapes of the 3 x tf32 GEMM: three tf32 MMAs per algorithmic multiply-add; tf32 peak = half the
            # measured dense bf16 rate (no tf32 figure in MEASURED_PEAKS.json)
            tf32 = 3.0 * flops / (ms * 1e-3) / 1e12
            peak = 0.5 * peaks["bf16_tflops"]
            out[name].update({"bound": "tensor", "achieved": round(tf32, 1), "peak": round(peak, 1), "unit": "TFLOP/s (tf32 MMAs issued)",
                              "frac": round(tf32 / peak, 4), "fp32_equivalent_tflops": round(tf32 / 3.0, 1)})
    return {"shape": {"M": M, "K": K, "N": N, "attention": [Bq, Hh, Tt, Tt]}, "kernels": out}


def layer_core_rooflines(peaks, M, Mg):
    """The quantised layer's own kernels at the token counts they run at in the stacked co-training step, for the model's three
    routed shapes: the stack holds M = 3 x batch x frames/4 rows; the LayerNorm quantiser, the backward prep pass and grad_W (both
    bitwidth groups in one launch) see all of them, the forward and grad_x GEMMs run once per bitwidth group - quoted at the
    larger group, Mg = 2M/3 rows (`layer_kernels` above is quoted at the smaller one).
    Same method as hot_kernel_rooflines: device time from a graph replay over rotating operand sets larger than L2."""
    import onebit_b200 as ob
    from onebit_b200 import _cabi, fused
    lib = _cabi.lib
    dev = torch.device("cuda", torch.cuda.current_device())
    cur = lambda: torch.cuda.current_stream().cuda_stream        # noqa: E731  (looked up per call: the timing graphs capture on a side stream)
    out = {}
    thr = int(round(0.1 * 65536))
    ik = 65536.0 / (65536 - thr)
    for K, N in ((256, 1024), (1024, 256), (256, 256)):
        torch.manual_seed(0)
        layer = ob.QuantizedLinear(K, N).to(dev)
        pk, pkt = layer.packed_weight(2)
        a, nb = layer.alpha, 3
        xs = [torch.randn(M, K, device=dev) for _ in range(nb)]
        gys = [torch.randn(M, N, device=dev) for _ in range(nb)]
        qs = [ob.act_quant_int8(x) for x in xs]
        ys = [torch.empty(M, N, device=dev) for _ in range(nb)]
        dys = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
        dxs = [torch.empty(M, K, device=dev) for _ in range(nb)]
        lnw, lnb = torch.ones(K, device=dev), torch.zeros(K, device=dev)
        stats = torch.empty(2, M, device=dev)
        colsum = torch.empty(lib.ob_bwd_colsum_blocks(M), N, device=dev)
        gw, ga, gb = torch.empty(N, K, device=dev), torch.empty((), device=dev), torch.empty(N, device=dev)
        nbytes = lib.ob_bwd_dw_workspace_bytes(M, N, K)
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        fns = {
            "ln_quant_fwd": (lambda j: lib.ob_layernorm_quant_fwd(xs[j].data_ptr(), lnw.data_ptr(), lnb.data_ptr(), 1e-5, M, K, qs[j][0].data_ptr(),
                                                                  qs[j][1].data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(), cur()),
                             5.0 * M * K + 12 * M),
            "gemm_fwd": (lambda j: lib.ob_gemm_tern_i8_fwd(qs[j][0].data_ptr(), qs[j][1].data_ptr(), pk.data_ptr(), a.data_ptr(), 1,
                                                           layer.bias.data_ptr(), Mg, N, K, ys[j].data_ptr(), 0, cur()),
                         Mg * K + N * K / 4 + 4.0 * Mg * N + 4 * Mg + 4 * N),
            "gemm_fwd_tail": (lambda j: lib.ob_gemm_tern_i8_fwd_tail(qs[j][0].data_ptr(), qs[j][1].data_ptr(), pk.data_ptr(), a.data_ptr(), 1,
                                                                     layer.bias.data_ptr(), Mg, N, K, gys[j].data_ptr(), None, 0.5 * ik, 7, 4 * j, thr,
                                                                     0, ys[j].data_ptr(), cur()),
                              Mg * K + N * K / 4 + 8.0 * Mg * N + 4 * Mg + 4 * N),
            "bwd_prep": (lambda j: lib.ob_bwd_prep(gys[j].data_ptr(), 0, qs[j][1].data_ptr(), qs[j][0].data_ptr(), M, N, K, dys[j].data_ptr(),
                                                   None, colsum.data_ptr(), cur()), 6.0 * M * N + 4 * M),
            "bwd_prep_tail": (lambda j: lib.ob_bwd_prep_fused(gys[j].data_ptr(), 1, None, None, 0.5 * ik, 7, 4 * j, thr, 0, qs[j][1].data_ptr(),
                                                              qs[j][0].data_ptr(), M, N, K, dys[j].data_ptr(), None, colsum.data_ptr(), cur()),
                              6.0 * M * N + 4 * M),
            "bwd_dx": (lambda j: lib.ob_bwd_dx(dys[j].data_ptr(), qs[j][1].data_ptr(), pkt.data_ptr(), a.data_ptr(), 1, Mg, N, K, dxs[j].data_ptr(), 0, cur()),
                       2.0 * Mg * N + N * K / 4 + 4.0 * Mg * K + 4 * Mg),
            # both bitwidth groups of the stack in one launch + one finaliser (rows [0, 2M/3) at 2 bits, the rest at 1 bit)
            "bwd_dw": (lambda j: lib.ob_bwd_dw_q8_groups(dys[j].data_ptr(), qs[j][0].data_ptr(), colsum.data_ptr(), layer.weight.data_ptr(),
                                                         a.data_ptr(), 1, (2 * M) // 3, M, N, K, gw.data_ptr(), ga.data_ptr(), gb.data_ptr(),
                                                         ws.data_ptr(), nbytes, cur()),
                       2.0 * M * N + 1.0 * M * K + 8.0 * N * K),
        }
        if N % 256 == 0:
            fns["bwd_prep_swish"] = (lambda j: lib.ob_bwd_prep_fused(gys[j].data_ptr(), 2, None, ys[j].data_ptr(), ik, 7, 4 * j, thr, 0,
                                                                     qs[j][1].data_ptr(), qs[j][0].data_ptr(), M, N, K, dys[j].data_ptr(), None,
                                                                     colsum.data_ptr(), cur()), 10.0 * M * N + 4 * M)
        i = [0]
        rows = {}
        for name, (fn, nbytes_alg) in fns.items():
            def call(j, fn=fn):
                rc = fn(j % nb)
                if rc != 0:
                    raise RuntimeError(_cabi.last_error())
            for j in range(3):
                call(j)
            ms = _device_time_us(call, 18) / 1e3                     # device time (graph replay), see hot_kernel_rooflines
            gbs = nbytes_alg / (ms * 1e-3) / 1e9
            rows[name] = {"us": round(ms * 1e3, 1), "gbs": round(gbs, 1), "frac": round(gbs / peaks["hbm_gbs"], 3)}
        out[f"{K}->{N}"] = rows
        del xs, gys, qs, ys, dys, dxs, ws
    return {"M": M, "M_group": Mg, "rows": {"ln_quant_fwd, bwd_prep*, bwd_dw": M, "gemm_fwd*, bwd_dx": Mg}, "peak_gbs": peaks["hbm_gbs"],
            "shapes": out}


def ctc_kernel_rooflines(peaks, B, T, V, L, blank=3):
    """The CTC loss kernels (csrc/ob_ctc.cu) at the step's shape through the C ABI: forward (row log-sum-exp + alpha/beta recursion
    + mean) and backward (gradient pass); the 0.5 GB logits exceed L2 by themselves."""
    from onebit_b200 import _cabi
    lib = _cabi.lib
    dev = torch.device("cuda", torch.cuda.current_device())
    st = torch.cuda.current_stream().cuda_stream
    torch.manual_seed(0)
    x = torch.randn(B, T, V, device=dev)
    dx = torch.empty(B, T, V, device=dev)
    targets = torch.randint(4, V, (B, L), device=dev)
    in_lens = torch.full((B,), T, dtype=torch.long, device=dev)
    tgt_lens = torch.full((B,), L, dtype=torch.long, device=dev)
    Sp = lib.ob_ctc_state_pitch(L)
    lse, ab = torch.empty(B * T, device=dev), torch.empty(2, B, T, Sp, device=dev)
    nll, loss, gout = torch.empty(B, device=dev), torch.empty((), device=dev), torch.ones((), device=dev)

    def fwd():
        rc = lib.ob_ctc_loss_fwd(x.data_ptr(), V, in_lens.data_ptr(), targets.data_ptr(), L, tgt_lens.data_ptr(), B, T, V, L, blank,
                                 lse.data_ptr(), ab[0].data_ptr(), ab[1].data_ptr(), nll.data_ptr(), loss.data_ptr(), st)
        if rc != 0:
            raise RuntimeError(_cabi.last_error())

    def bwd():
        rc = lib.ob_ctc_loss_bwd(x.data_ptr(), V, in_lens.data_ptr(), targets.data_ptr(), L, tgt_lens.data_ptr(), B, T, V, L, blank,
                                 lse.data_ptr(), ab[0].data_ptr(), ab[1].data_ptr(), nll.data_ptr(), gout.data_ptr(), dx.data_ptr(), V, st)
        if rc != 0:
            raise RuntimeError(_cabi.last_error())
    out = {}
    for name, fn, nbytes_alg in (("ctc_loss_fwd", fwd, 4.0 * B * T * V + 8.0 * B * T * Sp + 4.0 * B * T),
                                 ("ctc_loss_bwd", bwd, 8.0 * B * T * V + 8.0 * B * T * Sp)):
        for _ in range(3):
            fn()
        ms = timed_region(1, fn, 10) / 10
        gbs = nbytes_alg / (ms * 1e-3) / 1e9
        out[name] = {"ms": round(ms, 4), "bound": "hbm", "achieved": round(gbs, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": round(gbs / peaks["hbm_gbs"], 4), "algorithmic_bytes": int(nbytes_alg), "tflops": None}
    out["ctc_loss_fwd"]["note"] = ("three launches: row log-sum-exp (HBM-bound), the alpha/beta recursion (T serial frames per "
                                   "utterance: latency-bound), the mean")
    out["ctc_loss_fwd"]["shape"] = out["ctc_loss_bwd"]["shape"] = [B, T, V, L]
    return out


def cpu_baseline_child(args):
    """cpu_baseline of the GPU arm = the reference arm run in a CHILD process (this process holds the product and must not
    import the reference; the child must not import the product).  Returns the child's `cpu_baseline` object."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1",
           "--dropout", str(args.dropout), "--workload", "train"]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=ROOT)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
        return json.loads(line)["cpu_baseline"]
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": "audio-s/s", "cores": os.cpu_count(), "kind": "reference",
                "sample": f"child process failed: {type(e).__name__}: {e}"}


def run_train(args, world, rank):
    import onebit_b200 as ob
    from onebit_b200 import _cabi
    from onebit_b200.dp import GradAllReducer
    from onebit_b200 import routes
    from onebit_b200.training import StepConfig, reserve_allocator_headroom, train_step
    peaks = load_peaks()
    dev = torch.device("cuda", torch.cuda.current_device())
    T = args.frames
    if args.global_batch:
        if args.global_batch % world:
            raise SystemExit(f"--global-batch {args.global_batch} is not divisible by {world} ranks")
        B = args.global_batch // world                      # strong scaling: the global batch is fixed (configs[3]: 512)
    else:
        B = args.batch                                      # weak scaling: the per-GPU batch is fixed
    n_micro = -(-B // args.max_micro_batch)
    if B % n_micro:
        raise SystemExit(f"per-GPU batch {B} does not split into {n_micro} equal micro-batches")
    Bm = B // n_micro
    if args.tf32_nonrouted:
        torch.backends.cuda.matmul.allow_tf32 = True
    torch.manual_seed(0)                                   # same weights and (CPU RNG) precision masks on every rank
    model = ob.ConformerASR(TRAIN["mel"], TRAIN["vocab"], enc_dropout=args.dropout, dec_dropout=args.dropout).train().to(dev)
    model.use_packed_code_arena()                          # all 108 layers x 2 bitwidths re-quantised by ONE launch per step
    torch.cuda.manual_seed(1234 + rank)                    # dropout streams differ per replica (weights were drawn on the CPU)
    opt = torch.optim.AdamW(model.parameters(), lr=5e-4, betas=(0.9, 0.98), weight_decay=1e-2,
                            fused=not getattr(args, "foreach_adamw", False))
    sync = GradAllReducer(model.parameters(), buffers=model.buffers()) if world > 1 else None
    if sync is not None:
        sync.timing = True
    cfg = StepConfig(share_frontend=args.share_frontend, stack_passes=args.share_frontend and not args.no_stack_passes)
    batch = [make_batch(Bm, T, 1000 + rank * 16 + i, device=dev) for i in range(n_micro)]
    batch = batch[0] if n_micro == 1 else batch
    warm = max(args.warmup, 3)

    def step():
        return train_step(model, batch, opt, cfg, grad_sync=sync)[0]
    headroom = 0
    for i in range(warm):
        step()
        if i == 0:                                          # the steady allocations exist now: add the slack block
            headroom = reserve_allocator_headroom(dev, args.allocator_headroom_gib)
    mallocs0 = torch.cuda.memory_stats().get("num_device_alloc", 0)
    routes.reset()
    if sync is not None:
        sync.exposed_ms()                                   # drop the warm-up samples
    l0 = _cabi.lib.ob_launch_count()
    sampler = ClockSampler(torch.cuda.current_device()).start()
    window = os.environ.get("OB_NCU_WINDOW") == "1"     # `ncu --profile-from-start off` then sees only the timed steps
    if window:
        torch.cuda.profiler.start()
    ms = timed_region(world, step, args.steps)
    if window:
        torch.cuda.profiler.stop()
    clocks = sampler.stop()
    launches = _cabi.lib.ob_launch_count() - l0
    timed_mallocs = torch.cuda.memory_stats().get("num_device_alloc", 0) - mallocs0     # cudaMallocs inside the timed steps
    taken = routes.counts()                                 # which implementation every op around the layer took (CUDA tensors)
    exposed = sync.exposed_ms() if sync is not None else None
    if sync is not None:
        sync.timing = False
    ms_per_step = ms / args.steps
    audio_s = world * B * T * FRAME_S
    value = audio_s / (ms_per_step * 1e-3)

    if window:
        return {"metric": "conformer_train_audio_sec_per_sec", "value": round(value, 1), "unit": "audio-s/s",
                "ms_per_step": round(ms_per_step, 2), "gpu_launches": int(launches), "note": "OB_NCU_WINDOW run (profiling aid)"}
    # end to end through the public API: host (pinned) batch -> device each step, loss read back each step
    hosts = [make_batch(Bm, T, 2000 + rank * 16 + i, pinned=True) for i in range(n_micro)]

    def e2e_step():
        mbs = []
        for host in hosts:
            b = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
            b["feat_lens_cpu"], b["token_lens_cpu"] = host["feat_lens"], host["token_lens"]
            mbs.append(b)
        return train_step(model, mbs[0] if n_micro == 1 else mbs, opt, cfg, grad_sync=sync)[0].item()
    e2e_step()
    e2e_steps = max(2, min(args.steps, 5))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    last = 0.0
    for _ in range(e2e_steps):
        last = e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([e2e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = t.item()
    h2d = sum(v.numel() * v.element_size() for host in hosts for v in host.values())
    e2e = {"value": round(audio_s / (e2e_ms * 1e-3), 1), "unit": "audio-s/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": 4, "ms_per_step": round(e2e_ms, 2), "last_loss": round(last, 4),
           "api": "onebit_b200.training.train_step(model, host_batch, AdamW) incl. H2D of the batch and loss.item()"}

    out = {"metric": "conformer_train_audio_sec_per_sec", "value": round(value, 1), "unit": "audio-s/s", "n_gpus": world,
           "steps": args.steps, "warmup": warm, "ms_per_step": round(ms_per_step, 2), "higher_is_better": True,
           "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None,
           "dtype": "int8 x ternary fwd (int32 accum), bf16 bwd (fp32 accum); non-routed matmuls fp32 via 3 x tf32 split "
                    "on tcgen05 (error <= 1e-5 of max, fp32 accumulate), other non-routed ops fp32",
           "data": "synthetic", "impl": "ours",
           "config": train_workload_config(args, world),
           "run_info": {"audio_s_per_step": audio_s, "micro_batches_per_step": n_micro, "micro_batch": Bm,
                      "allreduce_exposed_ms": None if exposed is None else round(exposed, 3), "tf32_nonrouted": bool(args.tf32_nonrouted),
                      "torch_nonrouted": args.torch_nonrouted or None,
                      "optimizer": "torch.optim.AdamW(lr 5e-4, betas (0.9, 0.98), wd 1e-2" +
                                   (")" if args.foreach_adamw else ", fused=True) - the reference's update rule, single-kernel form"),
                      "share_frontend": bool(args.share_frontend),
                      "stack_passes": bool(cfg.stack_passes),
                      "stack_passes_note": "the teacher / student / stochastic-precision encoder passes share every weight and are "
                                           "evaluated side by side on one stacked batch: each routed layer runs once per bitwidth "
                                           "group, BatchNorm statistics stay per pass; same loss and gradients as three separate "
                                           "passes (tests/test_conformer_cpu.py, tests/test_gpu_conformer.py)",
                      "share_frontend_note": "conv subsampling (no dropout, no bitwidth) evaluated once per step for the three passes: "
                                             "common-subexpression sharing inside the step, same loss and gradients "
                                             "(tests/test_conformer_cpu.py); --no-share-frontend restores the 3x evaluation",
                      "allocator_headroom_gib": round(headroom / 2 ** 30, 1),
                      "cuda_mallocs_in_timed_steps": int(timed_mallocs)},
           "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "routes": taken,
           "peak_mem_gib": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}
    if rank == 0:
        M = Bm * (((T - 1) // 2 - 1) // 2)
        try:
            hk = hot_kernel_rooflines(peaks, M)
        except Exception as e:  # noqa: BLE001   (the headline must survive a failure of the per-kernel side tables)
            hk = {"shape": {"M": M, "K": 256, "N": 1024}, "kernels": {}, "error": f"{type(e).__name__}: {e}"}
        # dominant kernel of the layer inside the step = the one with the largest per-layer time at this shape
        core = ("act_quant_i8", "gemm_fwd", "bwd_prep", "bwd_dx", "bwd_dw")
        dom = max(((k, v) for k, v in hk["kernels"].items() if k in core), key=lambda kv: kv[1]["ms"],
                  default=("unmeasured", {"bound": "hbm", "achieved": None, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": None}))
        out["roofline"] = dict(kernel=dom[0], shape=hk["shape"],
                               traffic=ncu_traffic(f"{dom[0]}_{M}x{hk['shape']['K']}x{hk['shape']['N']}_f32"),
                               **{k: dom[1][k] for k in ("bound", "achieved", "peak", "unit", "frac")},
                               peak_source=f"{peaks['source']} HBM copy bandwidth (MEASURED_PEAKS.json)",
                               note="dominant kernel of the BitLinear layer (the north_star path) at the model's widest routed "
                                    "shape; every kernel of the step, incl. the fp32 tensor-core GEMM of the non-routed matmuls "
                                    "(the largest single family of the step), is in layer_kernels")
        out["layer_kernels"] = hk
        if cfg.stack_passes:
            try:
                stk = layer_core_rooflines(peaks, 3 * M, 2 * M)
                out["layer_kernels_stacked"] = stk
                # The step runs the layer on the stacked batch, so the line's roofline is the layer kernel with the largest device
                # time there (widest routed shape): achieved = algorithmic bytes / device time of one launch, both from this table.
                wide = stk["shapes"]["256->1024"]
                name, row = max(wide.items(), key=lambda kv: kv[1]["us"])
                rows_at = stk["M_group"] if name.startswith(("gemm_fwd", "bwd_dx")) else stk["M"]
                out["roofline"] = dict(kernel=name, shape={"M": rows_at, "K": 256, "N": 1024},
                                       traffic=ncu_traffic(f"{name}_{rows_at}x256x1024_f32"), bound="hbm", achieved=row["gbs"],
                                       peak=peaks["hbm_gbs"], unit="GB/s", frac=row["frac"],
                                       peak_source=f"{peaks['source']} HBM copy bandwidth (MEASURED_PEAKS.json)",
                                       note="the kernel of the BitLinear layer (the north_star path) with the largest device time per launch "
                                            "at the token counts of the stacked step (layer_kernels_stacked, widest routed shape); every "
                                            "kernel of the step, incl. the fp32 tensor-core GEMM of the non-routed matmuls, is in layer_kernels")
            except Exception as e:  # noqa: BLE001
                out["layer_kernels_stacked"] = {"error": f"{type(e).__name__}: {e}"}
        try:                                                # added late in round 1: a failure here must not cost the bench line
            hk["kernels"].update(ctc_kernel_rooflines(peaks, Bm, ((T - 1) // 2 - 1) // 2, TRAIN["vocab"], TRAIN["tokens"]))
        except Exception as e:  # noqa: BLE001
            hk["kernels"]["ctc_loss"] = {"error": f"{type(e).__name__}: {e}"}
        if world == 1:
            # SURVEY.md section 8(d) also asks for ONE precision-2 pass (forward + backward + AdamW), next to the full step
            from onebit_b200.training import att_ce_loss, ctc_loss_from_logits, make_att_targets

            batch0 = batch if n_micro == 1 else batch[0]

            def one_pass(batch=batch0):
                opt.zero_grad(set_to_none=True)
                enc, mask, ctc = model(batch, 2)
                t_inp, t_out, t_pad = make_att_targets(batch["tokens"], cfg.bos_id, cfg.eos_id, cfg.pad_id)
                logits = model.decode_logits(enc, mask, t_inp, t_pad)
                lens = torch.clamp(batch["feat_lens_cpu"] // 4, max=enc.size(1))
                loss = 0.8 * att_ce_loss(logits, t_out, cfg.pad_id, 0.1) + 0.2 * ctc_loss_from_logits(
                    ctc, lens, batch["tokens"], batch["token_lens_cpu"], cfg.blank_id)
                loss.backward()
                torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=cfg.max_grad_norm)
                opt.step()
            for _ in range(2):
                one_pass()
            ms1 = timed_region(1, one_pass, 3) / 3
            out["variants"] = {"one_precision2_pass": {"ms_per_step": round(ms1, 2),
                                                       "audio_s_per_s": round(Bm * T * FRAME_S / (ms1 * 1e-3), 1)}}
        if world == 1 and not args.no_gemm:
            out["bitlinear_gemm"] = gemm_microbench(args, 20, sweep=not args.no_sweep_summary)
            if "sweep" in out["bitlinear_gemm"]:
                out["bitlinear_gemm"]["sweep_summary"] = sweep_summary(out["bitlinear_gemm"].pop("sweep"), peaks)
        if world == 1 and not args.no_small_m:
            try:
                out["small_batch"] = small_batch_table(peaks)
            except Exception as e:  # noqa: BLE001
                out["small_batch"] = {"error": f"{type(e).__name__}: {e}"}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_child(args)
    return out


def run_infer(args, world, rank):
    """configs[4]: batched inference with packed 2-bit weights and device-side greedy CTC decoding."""
    import onebit_b200 as ob
    from onebit_b200 import _cabi
    from onebit_b200.inference import pack_model_for_inference, transcribe_greedy
    dev = torch.device("cuda", torch.cuda.current_device())
    T = 1000
    torch.manual_seed(0)
    model = ob.ConformerASR(TRAIN["mel"], TRAIN["vocab"]).to(dev)
    pack_model_for_inference(model, 2)
    rows = []
    clocks, launches, headline = None, 0, None
    for B in (1, 8, 64, 256):
        feats = torch.randn(B, T, TRAIN["mel"], generator=torch.Generator().manual_seed(B + rank))
        host = {"feats": feats.pin_memory(), "feat_lens": torch.full((B,), T, dtype=torch.long)}
        batch = {k: v.to(dev) for k, v in host.items()}

        def step():
            return transcribe_greedy(model, batch, precision=2)

        def e2e_step():
            b = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
            toks, n = transcribe_greedy(model, b, precision=2)
            return toks.cpu(), n.cpu()
        for _ in range(3):
            step()
        l0 = _cabi.lib.ob_launch_count()
        sampler = ClockSampler(torch.cuda.current_device()).start() if B == 256 else None
        ms = timed_region(world, step, args.steps) / args.steps
        if sampler is not None:
            clocks = sampler.stop()
            launches = _cabi.lib.ob_launch_count() - l0
        e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            e2e_step()
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) / 3 * 1e3
        # the same pass replayed from a CUDA graph (one driver call instead of ~1300 launches)
        from onebit_b200.inference import GraphedTranscriber
        runner = GraphedTranscriber(model, B, T, TRAIN["mel"], precision=2)
        for _ in range(3):
            runner(batch["feats"], batch["feat_lens"])
        graph_ms = timed_region(world, lambda: runner(batch["feats"], batch["feat_lens"]), args.steps) / args.steps
        del runner
        row = {"batch": B, "ms": round(ms, 3), "graph_ms": round(graph_ms, 3),
               "graph_audio_s_per_s": round(world * B * T * FRAME_S / (graph_ms * 1e-3), 1),
               "utt_per_s": round(world * B / (ms * 1e-3), 1),
               "audio_s_per_s": round(world * B * T * FRAME_S / (ms * 1e-3), 1),
               "e2e_audio_s_per_s": round(world * B * T * FRAME_S / (e2e_ms * 1e-3), 1),
               "h2d_bytes": feats.numel() * 4, "d2h_bytes": B * (((T - 1) // 2 - 1) // 2) * 4 + B * 4}
        rows.append(row)
        headline = row
    out = {"metric": "conformer_infer_audio_sec_per_sec", "value": headline["audio_s_per_s"], "unit": "audio-s/s", "n_gpus": world,
           "steps": args.steps, "warmup": 3, "ms_per_step": headline["ms"], "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "int8 x ternary (int32 accum); non-routed ops fp32 as in the reference", "data": "synthetic",
           "impl": "ours",
           "config": {"workload": "Conformer batched inference (BASELINE configs[4]): packed 2-bit weights, precision 2, 1000 frames, "
                                  "greedy CTC decode on the device; headline = batch 256", "batches": rows,
                      "parallelism": f"dp{world}"},
           "e2e": {"value": headline["e2e_audio_s_per_s"], "unit": "audio-s/s", "h2d_bytes_per_step": headline["h2d_bytes"],
                   "d2h_bytes_per_step": headline["d2h_bytes"], "api": "onebit_b200.inference.transcribe_greedy on a host-pinned batch, "
                   "tokens copied back to the host"},
           "gpu_launches": int(launches), "clocks": clocks}
    return out


def run_gemm(args, world, rank):
    res = gemm_microbench(args, args.steps, sweep=args.sweep and rank == 0)
    M, K, N, bw = args.tokens, args.in_features, args.out_features, args.bitwidth
    # every rank runs the same GEMM (weak scaling, no collective on this path): aggregate = world x per-rank
    ms = timed_region(world, lambda: None, 1)  # barrier only
    del ms
    import onebit_b200 as ob
    from onebit_b200 import _cabi, quant as obq
    dev = torch.device("cuda", torch.cuda.current_device())
    layer = ob.QuantizedLinear(K, N).to(dev)
    x_host = torch.randn(M, K).to(torch.bfloat16).pin_memory()
    y_host = torch.empty(M, N, dtype=torch.bfloat16).pin_memory()
    x_dev = torch.empty(M, K, device=dev, dtype=torch.bfloat16)

    def e2e_step():
        x_dev.copy_(x_host, non_blocking=True)
        with torch.no_grad():
            y = layer(x_dev, bw)
        y_host.copy_(y, non_blocking=True)
    for _ in range(3):
        e2e_step()
    l0 = _cabi.lib.ob_launch_count()
    sampler = ClockSampler(torch.cuda.current_device()).start()
    e2e_ms = timed_region(world, e2e_step, max(3, args.steps)) / max(3, args.steps)
    clocks = sampler.stop()
    launches = _cabi.lib.ob_launch_count() - l0
    ops = 2.0 * M * N * K
    out = {"metric": "bitlinear_int8_tops", "value": round(world * res["step_tops"], 1), "unit": "TOPS", "n_gpus": world,
           "steps": args.steps, "warmup": 3, "ms_per_step": res["step_ms"], "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "int8 x ternary (int32 accum)", "data": "synthetic", "impl": "ours",
           "config": {"workload": f"BitLinear fwd (act-quant + GEMM) M={M} K={K} N={N} bitwidth={bw} (BASELINE configs[1])",
                      "l2": "operands (q 134 MB, y 268 MB) exceed the 126 MB L2; no explicit flush", "parallelism": f"dp{world}"},
           "roofline": res["roofline"], "act_quant": res["act_quant"],
           "e2e": {"value": round(world * ops / (e2e_ms * 1e-3) / 1e12, 2), "unit": "TOPS", "h2d_bytes_per_step": M * K * 2,
                   "d2h_bytes_per_step": M * N * 2, "ms_per_step": round(e2e_ms, 3),
                   "api": "QuantizedLinear.forward(x_bf16, 2) with host-pinned input and output"},
           "gpu_launches": int(launches), "clocks": clocks}
    if "sweep" in res:
        out["sweep"] = res["sweep"]
    return out


def run_reference(args, world, rank):
    """Reference arm: the UNMODIFIED reference from baseline/_ref on the host cores (rank 0 only).  Never imports the product."""
    if rank != 0:
        return None
    from oracle import ref_loader
    if not ref_loader.available():
        return {"impl": "reference", "unavailable": "baseline/_ref not built (python oracle/build_ref.py needs /root/reference)"}
    cores = os.cpu_count()
    if args.workload == "train":
        from oracle import ref_bench
        r = ref_bench.run(args.steps, args.warmup, args.dropout)
        c1 = r["config"]
        v, dt, unit, metric = r["audio_s_per_s"], r["s_per_step"], "audio-s/s", "conformer_train_audio_sec_per_sec"
        cfg = train_workload_config(args, world)
        steps_done, warm = r["steps_timed"], args.warmup
        sample_txt = (f"bounded sample of that step = BASELINE configs[0]: {c1['batch']} utterances x {c1['frames']} frames, "
                      f"{c1['tokens']} tokens, default dims, V={c1['vocab']}, dropout {args.dropout}, weights seed 0 / inputs seed 1; "
                      f"the reference's own train.py:run_epoch (3 passes + losses + backward + clip + AdamW) from baseline/_ref, "
                      f"torch-CPU fp32, {r['threads']} threads; {dt:.2f} s/step over {steps_done} steps"
                      + (f"; one precision-2 pass fwd+bwd+AdamW {r['one_pass']['s_per_step']:.2f} s "
                         f"= {r['one_pass']['audio_s_per_s']:.1f} audio-s/s" if "one_pass" in r else ""))
        extra = {"one_precision2_pass": r.get("one_pass"), "full_step": {"s_per_step": round(dt, 3), "audio_s_per_s": round(v, 2)},
                 "sample_config": c1, "last_loss": r["last_loss"]}
    else:
        ref = ref_loader.load()
        torch.set_num_threads(cores)
        torch.manual_seed(0)
        rows = 2048
        layer = ref.quant.QuantizedLinear(args.in_features, args.out_features)
        x = torch.randn(rows, args.in_features)
        steps_done, warm = max(1, args.steps), max(1, args.warmup)
        with torch.no_grad():
            for _ in range(warm):
                layer(x, args.bitwidth)
            t0 = time.perf_counter()
            for _ in range(steps_done):
                layer(x, args.bitwidth)
        dt = (time.perf_counter() - t0) / steps_done
        v = 2.0 * rows * args.in_features * args.out_features / dt / 1e12
        unit, metric = "TOPS", "bitlinear_int8_tops"
        cfg = {"workload": f"BitLinear fwd (act-quant + GEMM) M={args.tokens} K={args.in_features} N={args.out_features} "
                           f"bitwidth={args.bitwidth} (BASELINE configs[1])",
               "l2": "operands (q 134 MB, y 268 MB) exceed the 126 MB L2; no explicit flush", "parallelism": f"dp{world}"}
        sample_txt = (f"{rows} of {args.tokens} token rows per step through the reference's own QuantizedLinear.forward "
                      f"(quant.py:120-127: re-quantise W, fp32 F.linear) from baseline/_ref, torch-CPU, {cores} threads")
        extra = {}
    return {"metric": metric, "value": round(v, 4), "unit": unit, "n_gpus": world, "steps": steps_done, "warmup": warm,
            "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference", "config": cfg,
            "cpu_baseline": {"value": round(v, 4), "unit": unit, "cores": cores, "kind": "reference", "sample": sample_txt, **extra},
            "e2e": {"value": round(v, 4), "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}


def main():
    args = parse()
    if args.torch_nonrouted:
        os.environ["OB_TORCH_NONROUTED"] = args.torch_nonrouted      # read when the package is imported
    # exactly ONE line on stdout: anything libraries print to fd 1 meanwhile (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        out = run_reference(args, int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")))
        if out is not None:
            emit(out)
        return
    if args.gpus > 1 and int(os.environ.get("WORLD_SIZE", "1")) != args.gpus:
        sys.stderr.write(f"bench.py --gpus {args.gpus} must be launched with one rank per GPU: python -m torch.distributed.run "
                         f"--nnodes=1 --nproc-per-node {args.gpus} --master-addr 127.0.0.1 bench.py --gpus {args.gpus} ...\n")
        sys.exit(2)
    world, rank, _ = dist_setup()
    out = {"train": run_train, "gemm": run_gemm, "infer": run_infer}[args.workload](args, world, rank)
    if rank == 0:
        emit(out)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
