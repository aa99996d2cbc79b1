#!/bin/bash
# round 2, session 18: eigensolver threshold back to the rounding-noise bound; PSO PLAIN variant with one group of
# rows prefetched ahead
tag=r02s18
mkdir -p gpurun_out
for f in test_gpu_es test_gpu_sizes test_gpu_parity test_gpu_l3; do
  ( timeout 1200 python -m pytest tests/$f.py -m gpu -q 2>&1 | tail -40 ) > gpurun_out/${tag}_pytest_$f.log
  echo "$f: $(tail -1 gpurun_out/${tag}_pytest_$f.log)"
done
python profiles/prof_cfg.py slopes > gpurun_out/${tag}_slopes.txt 2>&1
cat gpurun_out/${tag}_slopes.txt
