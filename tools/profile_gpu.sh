#!/bin/bash
# Run on the B200 box (gpurun): ncu launch list of a few bench steps + --set full captures of the top kernels.
# Outputs land in gpurun_out/ (scratch); tools/summarize_profile.py turns them into profiles/ncu_summary_<tag>.{md,json}.
set -x
mkdir -p gpurun_out
TAG=${1:-r1}
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-2600} -c ${COUNT:-1400} --csv \
    --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
if [ "${FULL:-1}" = "1" ]; then
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 120 -c 3 \
    -o gpurun_out/gemm_${TAG} -f $BENCH > gpurun_out/ncu_gemm_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_ -s 60 -c 4 \
    -o gpurun_out/conv_${TAG} -f $BENCH > gpurun_out/ncu_conv_${TAG}.log 2>&1
fi
ls -la gpurun_out | tail -8
