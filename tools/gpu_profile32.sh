#!/bin/bash
# Runs on the GPU box (under gpurun): launch list of the bench command + one full ncu capture of each kernel of the path.
# usage: tools/gpu_profile32.sh <tag>     -> gpurun_out/<tag>_*.{csv,ncu-rep,so,log}      (then: python tools/summarise_profiles.py <tag> r02)
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
cp rl_mpc_lanemerging_b200/libmpcb200.so gpurun_out/${TAG}_lib.so
for H in 50 17; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_h${H}.csv \
      python bench.py --steps 2 --warmup 3 --horizon $H --no-cpu-baseline --no-sweep --env-ticks 0 --train-ticks 0 > gpurun_out/${TAG}_launches_h${H}.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fast32 -s 2 -c 1 -o gpurun_out/${TAG}_fast32_h${H} \
      python tools/prof_run.py $H 4096 fast > gpurun_out/${TAG}_ncu32_h${H}.log 2>&1
done
# prof_run makes 3 plans; per plan: predict_layers, fast32 (A), fast32 (B), fast_pull (hand-overs), fast_pull (full row), exact
timeout 600 ncu --set full --clock-control none --import-source on -k regex:predict_layers -s 2 -c 1 -o gpurun_out/${TAG}_predict_h50 \
    python tools/prof_run.py 50 4096 fast > gpurun_out/${TAG}_ncupred_h50.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fast_pull -s 4 -c 1 -o gpurun_out/${TAG}_fallback_h50 \
    python tools/prof_run.py 50 4096 fast > gpurun_out/${TAG}_ncufb_h50.log 2>&1
MPC_RASTER_ROWS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rasterise -s 1 -c 1 -o gpurun_out/${TAG}_rasterise_h50 \
    python tools/prof_run.py 50 256 grid > gpurun_out/${TAG}_ncuras_h50.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rasterise_rows -s 1 -c 1 -o gpurun_out/${TAG}_rasterise_rows_h50 \
    python tools/prof_run.py 50 256 grid > gpurun_out/${TAG}_ncurasrows_h50.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fast_pull -s 2 -c 1 -o gpurun_out/${TAG}_dense32_h50 \
    python tools/prof_run.py 50 1024 dense32 > gpurun_out/${TAG}_ncudense32_h50.log 2>&1
