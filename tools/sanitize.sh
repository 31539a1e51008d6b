#!/bin/bash
# compute-sanitizer over the per-kernel checks (run on the B200 box): memcheck and racecheck, one check per process, each
# under its own timeout; the summaries go to profiles/sanitizer_<tag>.txt.  usage: tools/sanitize.sh <tag> [checks...]
tag=${1:-r2}; shift
checks=${@:-"gemm_tc_single_tile_k64 gemm_tc_multi_tile_edges gemm_tc_epilogues layernorm_fwd_bwd attention_fwd_bwd preprocess_fwd_bwd loss_kernels adam_kernel generator_conv_kernels generator_native_small vit_loss_backward_s16"}
out=profiles/sanitizer_${tag}.txt
: > $out
for tool in memcheck racecheck; do
  for c in $checks; do
    echo "=== $tool $c" | tee -a $out
    SPLICE_B200_GRAPHS=0 timeout 900 compute-sanitizer --tool $tool --print-limit 5 --error-exitcode 9 \
        python tools/gpu_checks.py --run $c > /tmp/san_$c.log 2>&1
    rc=$?
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error:|hazard" /tmp/san_$c.log | head -8 | tee -a $out
    grep -E '^\{' /tmp/san_$c.log | python -c "import sys,json; [print('check ok:', json.loads(l).get('ok'), 'secs', round(json.loads(l).get('secs',0),1)) for l in sys.stdin]" | tee -a $out
    echo "exit code $rc" | tee -a $out
  done
done
