#!/bin/bash
# Runs on the GPU box (under gpurun): full ncu capture of fast32_kernel and of predict_layers_kernel (one launch each, H=50 and H=17).
# usage: tools/gpu_profile32.sh <tag>     -> gpurun_out/<tag>_*.{ncu-rep,so,log}
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
cp rl_mpc_lanemerging_b200/libmpcb200.so gpurun_out/${TAG}_lib.so
for H in 50 17; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fast32 -s 2 -c 1 -o gpurun_out/${TAG}_fast32_h${H} \
      python tools/prof_run.py $H 4096 fast > gpurun_out/${TAG}_ncu32_h${H}.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:predict_layers -s 2 -c 1 -o gpurun_out/${TAG}_predict_h50 \
    python tools/prof_run.py 50 4096 fast > gpurun_out/${TAG}_ncupred_h50.log 2>&1
