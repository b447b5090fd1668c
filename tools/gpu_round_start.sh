#!/usr/bin/env bash
# First GPU call of a round: everything that was changed after the previous round's GPU budget ran out, in the order that
# matters.  Usage (from the repo root):  gpurun --timeout 900 -- 'bash tools/gpu_round_start.sh'
# Outputs land in gpurun_out/round_start_*; copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
O=gpurun_out/round_start
# 1. parity suite through the C ABI
timeout 400 python -m pytest tests -x -q -m gpu > ${O}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a ${O}_pytest.log; tail -3 ${O}_pytest.log
# 2. the bench line (cuda_mallocs_in_timed_steps must read 0 with the allocator headroom; routes.torch must be empty)
timeout 300 python bench.py > ${O}_train.json 2> ${O}_train.err; echo "bench rc=$?"
python tools/show_bench.py ${O}_train.json 2>/dev/null | head -3
# 3. per-step times with and without the headroom (the late cudaMalloc stall of profiles/r01_step_times.json)
OB_HEADROOM_GIB=6 timeout 60 python tools/gpu_step_times.py 16 > ${O}_step_times_headroom.json 2>> ${O}_train.err
timeout 60 python tools/gpu_step_times.py 16 > ${O}_step_times_plain.json 2>> ${O}_train.err
# 4. ncu: launch list of the bench command, then one full capture each of the CTA-pair fp32 GEMM and the CTC kernels
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${O}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gemm > ${O}_ncu_launches.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'f32_gemm_kernel|ctc_' -c 8 -o ${O}_pair_ctc \
  python tools/gpu_ctc.py > ${O}_ncu_ctc.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:f32_gemm_kernel -c 6 -o ${O}_pair_gemm \
  python tools/gpu_f32pair.py vocabulary > ${O}_ncu_pair.log 2>&1
ls -la gpurun_out | tail -15
