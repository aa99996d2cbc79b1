#!/bin/bash
# round 2, session 11: vd_update with one column per thread (1024 threads at N = 1024)
tag=r02s11
mkdir -p gpurun_out
for f in test_gpu_es test_gpu_sizes test_gpu_parity; do
  ( timeout 1200 python -m pytest tests/$f.py -m gpu -q 2>&1 | tail -60 ) > gpurun_out/${tag}_pytest_$f.log
  echo "$f: $(tail -1 gpurun_out/${tag}_pytest_$f.log)"
done
python profiles/vd_clocks.py > gpurun_out/${tag}_vd_clocks.txt 2>&1
cat gpurun_out/${tag}_vd_clocks.txt
python profiles/prof_cfg.py slopes > gpurun_out/${tag}_slopes.txt 2>&1
cat gpurun_out/${tag}_slopes.txt
