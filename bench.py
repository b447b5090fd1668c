#!/usr/bin/env python
"""bench.py - the measurement contract ("Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload gemm|layer] [--impl ours|reference]

One "step" is one pass of the hot path over one batch of synthetic input.  Workloads:

  gemm   BASELINE.json configs[1]: the BitLinear forward at M = 65536 tokens, K = N = 2048, bitwidth 2
         (act-quant + ternary x int8 tcgen05 GEMM, bf16 output).  metric = BitLinear int8 TOPS.
  layer  the same layer forward+backward (adds bwd_prep, grad_x, grad_W + fused STE); metric = TFLOP/s-equivalent.

Multi-GPU (torchrun): every rank runs the same per-rank workload (weak scaling, no data-path collective for the
forward GEMM); the layer workload all-reduces the latent-weight gradients over NCCL each step.
`--impl reference` times the CPU restatement of the reference path (oracle/) on the host cores.
Prints ONE JSON line on rank 0.
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


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--workload", default="gemm", choices=["gemm", "layer"])
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--tokens", type=int, default=65536)
    p.add_argument("--in-features", type=int, default=2048)
    p.add_argument("--out-features", type=int, default=2048)
    p.add_argument("--bitwidth", type=int, default=2)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--sweep", action="store_true", help="also run the configs[1] K/N sweep (rank 0) and add it to the line")
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
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""

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

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower() == "active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def dist_setup(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def timed_region(world, fn, steps):
    """barrier + sync, K steps between CUDA events on the current stream, max over ranks (ms)."""
    import torch.distributed as dist
    if world > 1:
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
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = t.item()
    return ms


# ------------------------------------------------------------------------------------------ CPU baseline
def cpu_reference_layer(args, sample_rows, steps, backward):
    """The reference's path (fp32 weight quantiser + F.linear, quant.py:120-127) restated in oracle/, on the host."""
    from oracle.torch_oracle import OracleQuantizedLinear
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    layer = OracleQuantizedLinear(args.in_features, args.out_features, act_bits=32)   # Oracle-A: the pure reference
    x = torch.randn(sample_rows, args.in_features)
    gy = torch.randn(sample_rows, args.out_features)

    def step():
        if backward:
            xr = x.requires_grad_(True)
            layer(xr, args.bitwidth).backward(gy)
        else:
            with torch.no_grad():
                layer(x, args.bitwidth)
    step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    flops = 2.0 * sample_rows * args.in_features * args.out_features * (3 if backward else 1)
    return flops / dt / 1e12, dt


# ------------------------------------------------------------------------------------------ workloads
def run_ours(args, world, rank):
    import onebit_b200 as ob
    from onebit_b200 import _cabi, quant as obq
    peaks = load_peaks()
    dev = torch.device("cuda", torch.cuda.current_device())
    M, K, N, bw = args.tokens, args.in_features, args.out_features, args.bitwidth
    torch.manual_seed(0)
    layer = ob.QuantizedLinear(K, N).to(dev)
    x = torch.randn(M, K, device=dev, generator=torch.Generator(device=dev).manual_seed(1234 + rank))
    xb = x.to(torch.bfloat16)
    packed, packed_t = layer.packed_weight(bw)
    backward = args.workload == "layer"
    launches_per_step = 0

    if not backward:
        # hot path of the forward: activation quantiser + ternary x int8 GEMM (bf16 in -> bf16 out)
        def step():
            q, s = ob.act_quant_int8(xb)
            return obq.gemm_fwd(q, s, packed, layer.alpha, layer.bias, N, torch.bfloat16)
        launches_per_step = 2
        flops_per_step = 2.0 * M * N * K
    else:
        gy = torch.randn(M, N, device=dev, generator=torch.Generator(device=dev).manual_seed(99))
        xr = x.requires_grad_(True)
        grads = [layer.weight, layer.alpha, layer.bias]

        def step():
            obq._ActQuantCache.clear()
            for p in grads:
                p.grad = None
            xr.grad = None
            y = layer(xr, bw)
            y.backward(gy)
            if world > 1:
                import torch.distributed as dist
                flat = torch.cat([p.grad.reshape(-1) for p in grads])
                dist.all_reduce(flat)
        launches_per_step = 2 + 1 + 1 + 3     # act quant, fwd gemm | prep, dx, dw + ste + tail
        flops_per_step = 6.0 * M * N * K

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    ms = timed_region(world, step, args.steps)
    clocks = sampler.stop()
    ms_per_step = ms / args.steps
    value = world * flops_per_step / (ms_per_step * 1e-3) / 1e12

    # ---- dominant kernel alone (forward GEMM), CUDA events on its launch stream
    q, s = ob.act_quant_int8(xb)
    def gemm_only():
        obq.gemm_fwd(q, s, packed, layer.alpha, layer.bias, N, torch.bfloat16)
    for _ in range(3):
        gemm_only()
    gemm_ms = timed_region(1, gemm_only, args.steps) / args.steps
    gemm_tops = 2.0 * M * N * K / (gemm_ms * 1e-3) / 1e12
    gemm_bytes = M * K + N * K / 4 + 2.0 * M * N + 4 * M + 4 * N
    int8_peak = 2.0 * peaks["bf16_tflops"]
    # act quantiser alone (HBM-bound)
    def act_only():
        ob.act_quant_int8(xb)
    for _ in range(3):
        act_only()
    act_ms = timed_region(1, act_only, args.steps) / args.steps
    act_gbs = (M * K * 2 + M * K + 4 * M) / (act_ms * 1e-3) / 1e9
    ai = 2.0 * M * N * K / gemm_bytes
    roofline = {"kernel": "gemm_expand_kernel<kFwdI8> (ternary x int8, tcgen05 kind::i8)", "bound": "tensor",
                "achieved": round(gemm_tops, 1), "peak": round(int8_peak, 1), "unit": "TFLOP/s",
                "frac": round(gemm_tops / int8_peak, 4), "traffic": None,
                "peak_source": f"2 x {peaks['source']} bf16 burst {peaks['bf16_tflops']} (int8 dense is 2x bf16 nominally; "
                               "no int8 figure in MEASURED_PEAKS.json)",
                "gemm_ms": round(gemm_ms, 4), "arith_intensity_op_per_byte": round(ai, 1),
                "hbm_gbs_of_gemm": round(gemm_bytes / (gemm_ms * 1e-3) / 1e9, 1),
                "act_quant": {"bound": "hbm", "achieved": round(act_gbs, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                              "frac": round(act_gbs / peaks["hbm_gbs"], 4), "ms": round(act_ms, 4)}}
    try:   # cuBLASLt int8 cross-check for the int8 denominator (library GEMM, not the product)
        a8 = torch.randint(-127, 127, (8192, 8192), device=dev, dtype=torch.int8)
        b8 = torch.randint(-127, 127, (8192, 8192), device=dev, dtype=torch.int8).t()
        for _ in range(3):
            torch._int_mm(a8, b8)
        t = timed_region(1, lambda: torch._int_mm(a8, b8), 10) / 10
        roofline["int_mm_8192_tops"] = round(2 * 8192 ** 3 / (t * 1e-3) / 1e12, 1)
    except Exception as e:  # pragma: no cover
        roofline["int_mm_8192_tops"] = f"unavailable: {type(e).__name__}"

    # ---- end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
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
    e2e_steps = max(3, min(args.steps, 10))
    e2e_ms = timed_region(world, e2e_step, e2e_steps) / e2e_steps
    e2e = {"value": round(world * 2.0 * M * N * K / (e2e_ms * 1e-3) / 1e12, 2), "unit": "TOPS",
           "h2d_bytes_per_step": x_host.numel() * 2, "d2h_bytes_per_step": y_host.numel() * 2,
           "ms_per_step": round(e2e_ms, 3), "api": "QuantizedLinear.forward(x_bf16, 2) on host-pinned input/output"}

    out = {"metric": "bitlinear_int8_tops" if not backward else "bitlinear_fwd_bwd_tflops",
           "value": round(value, 2), "unit": "TOPS" if not backward else "TFLOP/s", "n_gpus": world,
           "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8 x ternary (int32 accum)",
           "data": "synthetic", "impl": "ours",
           "config": {"workload": f"BitLinear {'fwd' if not backward else 'fwd+bwd'} M={M} K={K} N={N} bitwidth={bw} "
                                  "(BASELINE configs[1], largest sweep point)",
                      "tokens_per_gpu": M, "in_features": K, "out_features": N, "bitwidth": bw,
                      "l2": "operands (q 134 MB, y 268 MB) exceed the 126 MB L2; no explicit flush",
                      "parallelism": f"dp{world}"},
           "roofline": roofline, "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "clocks": clocks}

    if args.sweep and rank == 0:
        sweep = []
        for Ms in (4096, 65536):
            for Ks in (256, 512, 1024, 2048):
                for Ns in (256, 512, 1024, 2048):
                    torch.manual_seed(0)
                    l2 = ob.QuantizedLinear(Ks, Ns).to(dev)
                    pk, _ = l2.packed_weight(bw)
                    qs, ss = ob.act_quant_int8(torch.randn(Ms, Ks, device=dev, dtype=torch.bfloat16))
                    f = lambda: obq.gemm_fwd(qs, ss, pk, l2.alpha, l2.bias, Ns, torch.bfloat16)
                    for _ in range(3):
                        f()
                    t = timed_region(1, f, 20) / 20
                    by = Ms * Ks + Ns * Ks / 4 + 2.0 * Ms * Ns + 4 * Ms + 4 * Ns
                    sweep.append({"M": Ms, "K": Ks, "N": Ns, "us": round(t * 1e3, 2),
                                  "tops": round(2.0 * Ms * Ns * Ks / (t * 1e-3) / 1e12, 1),
                                  "gbs": round(by / (t * 1e-3) / 1e9, 1)})
        out["sweep"] = sweep

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rows = 2048
        tf, dt = cpu_reference_layer(args, rows, 3, backward)
        out["cpu_baseline"] = {"value": round(tf, 4), "unit": out["unit"], "cores": os.cpu_count(), "kind": "port",
                               "sample": f"{rows} of {M} tokens, same K/N/bitwidth, fp32 torch-CPU restatement of "
                                         f"quant.py:120-127 (Oracle-A), {dt:.3f} s/step"}
    return out


def run_reference(args, world, rank):
    """Reference arm: the reference's own CPU implementation of the path (oracle port), all host threads."""
    if rank != 0:
        return None
    backward = args.workload == "layer"
    rows = 2048
    steps = max(1, min(args.steps, 5))
    tf, dt = cpu_reference_layer(args, rows, steps, backward)
    unit = "TOPS" if not backward else "TFLOP/s"
    M, K, N, bw = args.tokens, args.in_features, args.out_features, args.bitwidth
    return {"metric": "bitlinear_int8_tops" if not backward else "bitlinear_fwd_bwd_tflops", "value": round(tf, 4),
            "unit": unit, "n_gpus": world, "steps": steps, "warmup": 1, "ms_per_step": round(dt * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": f"BitLinear {'fwd' if not backward else 'fwd+bwd'} M={M} K={K} N={N} bitwidth={bw} "
                                   "(BASELINE configs[1], largest sweep point)",
                       "tokens_per_gpu": M, "in_features": K, "out_features": N, "bitwidth": bw,
                       "parallelism": "cpu"},
            "cpu_baseline": {"value": round(tf, 4), "unit": unit, "cores": os.cpu_count(), "kind": "port",
                             "sample": f"{rows} of {M} tokens per step (bounded sample), torch-CPU restatement of the "
                                       "reference layer (oracle/torch_oracle.py, Oracle-A)"},
            "e2e": {"value": round(tf, 4), "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def main():
    args = parse()
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        out = run_reference(args, int(os.environ.get("WORLD_SIZE", "1")), rank)
        if out is not None:
            print(json.dumps(out), flush=True)
        return
    world, rank, _ = dist_setup(args)
    out = run_ours(args, world, rank)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
