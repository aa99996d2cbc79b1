#!/bin/bash
# ncu launch lists (device time per launch) of the non-headline configs: bash profiles/gpu_launchlists.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
for w in vd vd64 cpso pso cma; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_${w}.csv \
     python profiles/prof_cfg.py $w > gpurun_out/${tag}_launches_${w}.log 2>&1
done
python profiles/prof_cfg.py cma_time > gpurun_out/${tag}_cma_time.log 2>&1
python profiles/prof_cfg.py eigh_time > gpurun_out/${tag}_eigh_time.log 2>&1
cat gpurun_out/${tag}_cma_time.log gpurun_out/${tag}_eigh_time.log
