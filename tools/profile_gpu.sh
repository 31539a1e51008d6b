#!/bin/bash
# Run on the B200 box (gpurun): ncu launch list of a few bench steps + one --set full capture of the top kernel.
# Outputs land in gpurun_out/ (scratch); summaries are copied to profiles/ by tools/summarize_profile.py here.
set -x
mkdir -p gpurun_out
TAG=${1:-r1}
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-3000} -c ${COUNT:-1400} --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 120 -c 3 \
    -o gpurun_out/gemm_${TAG} -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
