#!/bin/bash
# compute-sanitizer over the hand-written mbarrier / TMEM / TMA kernels: memcheck, racecheck and synccheck on the
# fused-kernel parity tests (training forward, dgrad chain, multicast clusters, in-kernel encoders, narrow net), bounded by timeouts.
tag=${1:-r02s}
out=gpurun_out
mkdir -p $out
K="test_fused_training_forward_matches_layered or test_fused_dgrad_chain_matches_layered or test_weight_multicast or inkernel or test_fused_forward_render"
for tool in memcheck racecheck synccheck; do
  timeout -s KILL 1000 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_tc_gpu.py -q -m gpu -x -k "$K" \
    > $out/${tag}_sanitizer_$tool.log 2>&1; echo "$tool rc=$?" | tee -a $out/${tag}_status.txt
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|hazard" $out/${tag}_sanitizer_$tool.log | tail -6
done
true
