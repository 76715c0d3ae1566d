#!/bin/bash
# Runs on the GPU box (under gpurun): launch list of the bench command + one full ncu capture of each kernel of the path.
# usage: tools/gpu_profile32.sh <tag> [1|2]   -> gpurun_out/<tag>_*.{csv,ncu-rep,so,log}      (then: python tools/summarise_profiles.py <tag> r02)
# Two parts (gpurun brings back at most 64 MiB per call): 1 = launch lists, fast32 at H=50 / H=17, predictor;
# 2 = hand-over kernel, the two rasterisers, K2 on dense grids.
set -u
TAG=${1:-r02}
PART=${2:-1}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
if [ "$PART" = "1" ]; then
cp rl_mpc_lanemerging_b200/libmpcb200.so gpurun_out/${TAG}_lib.so
for H in 50 17; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_h${H}.csv \
      python bench.py --steps 2 --warmup 3 --horizon $H --no-cpu-baseline --no-sweep --env-ticks 0 --train-ticks 0 > gpurun_out/${TAG}_launches_h${H}.log 2>&1
  timeout 600 $NCU -k regex:fast32 -s 2 -c 1 -o gpurun_out/${TAG}_fast32_h${H} \
      python tools/prof_run.py $H 4096 fast > gpurun_out/${TAG}_ncu32_h${H}.log 2>&1
done
# prof_run makes 3 plans; per plan: predict_layers, fast32 (A), fast32 (B), fast_pull (hand-overs), fast_pull (full row), exact
timeout 600 $NCU -k regex:predict_layers -s 2 -c 1 -o gpurun_out/${TAG}_predict_h50 \
    python tools/prof_run.py 50 4096 fast > gpurun_out/${TAG}_ncupred_h50.log 2>&1
else
timeout 600 $NCU -k regex:fast_pull -s 4 -c 1 -o gpurun_out/${TAG}_fallback_h50 \
    python tools/prof_run.py 50 4096 fast > gpurun_out/${TAG}_ncufb_h50.log 2>&1
MPC_RASTER_ROWS=0 timeout 600 $NCU -k regex:rasterise -s 1 -c 1 -o gpurun_out/${TAG}_rasterise_h50 \
    python tools/prof_run.py 50 256 grid > gpurun_out/${TAG}_ncuras_h50.log 2>&1
timeout 600 $NCU -k regex:rasterise_rows -s 1 -c 1 -o gpurun_out/${TAG}_rasterise_rows_h50 \
    python tools/prof_run.py 50 256 grid > gpurun_out/${TAG}_ncurasrows_h50.log 2>&1
# K2: per solve fast32 (A), fast32 (B), fast_pull (hand-overs), fast_pull (full row), exact; the second solve's first launch
timeout 600 $NCU -k regex:fast32 -s 2 -c 1 -o gpurun_out/${TAG}_dense32_h50 \
    python tools/prof_run.py 50 1024 dense32 > gpurun_out/${TAG}_ncudense32_h50.log 2>&1
fi
