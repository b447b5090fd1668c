"""Pretty-print a bench.py JSON line: python tools/show_bench.py file.json"""
import json
import sys

d = json.load(open(sys.argv[1]))
print(d.get("metric"), d.get("value"), d.get("unit"), "ms/step", d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"),
      "cpu", (d.get("cpu_baseline") or {}).get("value"), "launches", d.get("gpu_launches"))
for k, v in (d.get("layer_kernels") or {}).get("kernels", {}).items():
    unit = "GB/s" if v.get("bound", "hbm") == "hbm" else "TF/s"
    print(f"  {k:22s} {v['ms'] * 1e3:8.1f} us {v['achieved']:8.1f} {unit} {v['frac']:.3f}" + (f"  ({v['tflops']} TF/s)" if v.get("tflops") else ""))
g = d.get("bitlinear_gemm")
if g:
    print("  gemm", g["roofline"]["achieved"], g["roofline"]["frac"], "act", g["act_quant"]["achieved"])
