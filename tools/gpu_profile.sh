#!/bin/bash
# Runs on the GPU box (under gpurun): launch list of the bench command + one full capture of the dominant kernel.
# usage: tools/gpu_profile.sh <tag>     -> gpurun_out/<tag>_*.{csv,ncu-rep,so,json}
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
cp rl_mpc_lanemerging_b200/libmpcb200.so gpurun_out/${TAG}_lib.so
for H in 50 17; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_h${H}.csv \
      python bench.py --steps 2 --warmup 3 --horizon $H --no-cpu-baseline > gpurun_out/${TAG}_launches_h${H}.log 2>&1
  # fast_pull launches per plan(): main, full-row re-solve; skip the first plan() of prof_run, take the main launch of the second
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fast_pull -s 2 -c 1 -o gpurun_out/${TAG}_fast_h${H} \
      python tools/prof_run.py $H 4096 fast > gpurun_out/${TAG}_ncu_h${H}.log 2>&1
done
timeout 600 python bench.py > gpurun_out/${TAG}_bench_h50.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --horizon 17 > gpurun_out/${TAG}_bench_h17.json 2>> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench_h50.json
