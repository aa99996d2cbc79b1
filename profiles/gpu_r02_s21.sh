#!/bin/bash
# round 2, session 21: in-run kernel durations (CUPTI through torch.profiler) of the C3 / C5 / C4 chains
mkdir -p gpurun_out
for c in cpso pso vd de; do python profiles/prof_timeline.py $c 200 2>&1 | grep -v Warning >> gpurun_out/r02s21_timelines.txt; done
python profiles/prof_timeline.py cma 20 2>&1 | grep -v Warning >> gpurun_out/r02s21_timelines.txt
cat gpurun_out/r02s21_timelines.txt
