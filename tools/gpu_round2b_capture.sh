#!/usr/bin/env bash
# Evidence run for the second half of round 2 on one B200, in two parts (gpurun copies back at most 64 MiB: the .ncu-rep stays in /tmp,
# only its csv exports travel):  gpurun --timeout 1200 -- 'bash tools/gpu_round2b_capture.sh bench'   tests, the bench workloads (default
# line with the CPU baseline, the reference arm, PDL off, the GEMM sweep, inference), a torch-profiler breakdown;
#                                gpurun --timeout 1800 -- 'bash tools/gpu_round2b_capture.sh ncu'     the ncu launch list of the bench
# command and one `ncu --set full` capture of the layer's kernels.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02b
PART=${1:-bench}
if [ "$PART" = "bench" ]; then
timeout 900 python -m pytest tests -q -m gpu > ${O}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a ${O}_pytest.log; tail -2 ${O}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > ${O}_bench_train_1gpu.json 2> ${O}_bench_train_1gpu.err; echo "bench rc=$?"
python tools/show_bench.py ${O}_bench_train_1gpu.json 2>/dev/null | head -3
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_reference_arm.json 2> ${O}_bench_reference_arm.err; echo "reference arm rc=$?"
OB_PDL=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gemm --no-small-m > ${O}_bench_train_1gpu_nopdl.json 2>> ${O}_bench_train_1gpu.err
python tools/show_bench.py ${O}_bench_train_1gpu_nopdl.json 2>/dev/null | head -1
timeout 300 python bench.py --workload gemm --sweep --steps 20 > ${O}_bench_gemm_sweep.json 2> ${O}_bench_gemm.err; echo "gemm rc=$?"
timeout 300 python bench.py --workload infer --steps 10 > ${O}_bench_infer.json 2> ${O}_bench_infer.err; echo "infer rc=$?"
timeout 200 python tools/gpu_train_probe.py 64 1600 share > ${O}_torch_profiler_train_B64.txt 2>&1
exit 0
fi
OB_NCU_WINDOW=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 9000 --csv \
  --log-file ${O}_launches_train.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gemm --no-small-m > ${O}_ncu_launches.log 2>&1
python tools/ncu_launches.py ${O}_launches_train.csv > ${O}_launches_train_summary.txt 2>&1; head -30 ${O}_launches_train_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_expand|dw_|bwd_prep|act_quant|ln_quant|gemv|swish_drop_quant" \
  -o /tmp/r02b_layer python tools/gpu_layer_ncu.py > ${O}_ncu_layer.log 2>&1
ncu -i /tmp/r02b_layer.ncu-rep --page raw --csv > ${O}_layer_kernels_ncu_full_raw.csv 2>/dev/null
ncu -i /tmp/r02b_layer.ncu-rep --page details --kernel-name regex:"dw_pair|bwd_prep|gemm_expand" 2>/dev/null | grep -v "^\s*$" | cut -c1-160 > ${O}_layer_kernels_ncu_details.txt
python tools/ncu_traffic.py ${O}_layer_kernels_ncu_full_raw.csv | tail -30
gzip -f ${O}_launches_train.csv
ls -la gpurun_out | tail -20; du -sh gpurun_out
