"""profiles/ncu_traffic.json from the `ncu --set full` raw csv of tools/gpu_layer_ncu.py: DRAM bytes (read + write) per launch, keyed
the way bench.py looks them up (`<bench name>_<M>x<K>x<N>_<dtype>`).  The launches are matched by ORDER: gpu_layer_ncu.py runs a fixed
sequence twice per shape (the second repetition is read), and every match is checked against the expected kernel name.

    python tools/ncu_traffic_json.py gpurun_out/r02b_layer_kernels_ncu_full_raw.csv r02b_layer_kernels_ncu_full_raw.csv"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, kept = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(path)))
hdr, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(hdr)}
scale_b = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def dram(r):
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[col[m]].replace(",", "")) * scale_b.get(units[col[m]], 1)
    return tot


launches = [(re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("ob::", ""), r) for r in rows[2:] if len(r) >= len(hdr)]
SEQ = [("act_quant_i8", "act_quant_reg_kernel"), ("ln_quant_fwd", "ln_quant_fwd_kernel"), ("gemm_fwd", "gemm_expand_kernel<0"),
       ("bwd_prep", "bwd_prep_kernel<float, 0>"), ("bwd_dx", "gemm_expand_kernel<1"), ("bwd_dw", "dw_pair_kernel"),
       ("bwd_dw", "dw_finalize_kernel"), ("swish_drop_quant", "swish_drop_quant"), ("gemm_fwd_tail", "gemm_expand_kernel<0"),
       ("bwd_prep_swish", "bwd_prep_kernel<float, 2>")]
out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full --clock-control none` of tools/gpu_layer_ncu.py "
                   "(every launch behind a 256 MB cache flush; second repetition of each shape). Outputs smaller than the 126 MB L2 are only "
                   "partly written back when the kernel ends, so traffic can be below the algorithmic bytes. bwd_dw = GEMM + finaliser."}
pos = 0
for (M, K, N) in ((25536, 256, 1024), (76608, 256, 1024)):
    pos += len(SEQ)                                   # skip the warm-up repetition
    for name, expect in SEQ:
        kname, r = launches[pos]
        assert expect in kname, (pos, name, expect, kname)
        shape = f"{M}x{N}x{K}" if name == "gemm_fwd_tail" else f"{M}x{K}x{N}"      # the tail GEMM runs 1024 -> 256
        key = f"{name}_{shape}_f32"
        e = out.setdefault(key, {"bytes": 0, "source": kept, "kernel": ""})
        e["bytes"] += int(dram(r))
        e["kernel"] = (e["kernel"] + " + " if e["kernel"] else "") + kname
        pos += 1
# large GEMM: (act_quant bf16, gemm) x 2 repetitions
pos += 2
kname, r = launches[pos]
assert "act_quant_reg_kernel" in kname, kname
out["act_quant_i8_65536x2048x2048_bf16"] = {"bytes": int(dram(r)), "source": kept, "kernel": kname}
kname, r = launches[pos + 1]
assert "gemm_expand_kernel<0" in kname, kname
out["gemm_fwd_65536x2048x2048_bf16"] = {"bytes": int(dram(r)), "source": kept, "kernel": kname}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
for k, v in out.items():
    if k != "_comment":
        print(f"{k:48s} {v['bytes'] / 1e6:9.1f} MB  {v['kernel']}")
