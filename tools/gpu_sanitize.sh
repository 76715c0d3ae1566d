#!/bin/bash
# Runs on the GPU box (under gpurun): compute-sanitizer over tools/sanitize_run.py -> gpurun_out/r02_compute_sanitizer.txt
OUT=gpurun_out/r02_compute_sanitizer.txt
mkdir -p gpurun_out
echo "# compute-sanitizer (B200, tools/sanitize_run.py: plan fast+exact, finer_fit, build_grid (both rasterisers, fp32+fp64), solve_dense fast on both, K4 kernels, selftest) -- round 2 kernels" > $OUT
echo "## memcheck  (H=17, 24 mixed states)" >> $OUT
timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_run.py 17 24 2>&1 | grep -E "sanitize_run done|ERROR SUMMARY|Invalid|Error" | head -8 >> $OUT
echo "## memcheck  (H=50, 12 mixed states: ring window, hand-overs, side stream)" >> $OUT
timeout 700 compute-sanitizer --tool memcheck python tools/sanitize_run.py 50 12 2>&1 | grep -E "sanitize_run done|ERROR SUMMARY|Invalid|Error" | head -8 >> $OUT
echo "## racecheck --racecheck-report analysis (H=17, 8 mixed states)" >> $OUT
timeout 700 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_run.py 17 8 2>&1 | grep -E "sanitize_run done|RACECHECK SUMMARY|hazard|Error" | head -8 >> $OUT
echo "## synccheck (H=17, 8 mixed states)" >> $OUT
timeout 500 compute-sanitizer --tool synccheck python tools/sanitize_run.py 17 8 2>&1 | grep -E "sanitize_run done|ERROR SUMMARY|Barrier|Error" | head -8 >> $OUT
cat $OUT
