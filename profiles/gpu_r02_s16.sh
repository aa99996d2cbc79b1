#!/bin/bash
# round 2, session 16: vectorised restart reset, Jacobi quiet-sweep threshold 5e-5, zgen off by default;
# full GPU suite + slopes + eigensolver sweeps
tag=r02s16
mkdir -p gpurun_out
for f in test_gpu_parity test_gpu_es test_gpu_sizes test_gpu_jit test_gpu_l3 test_parallel; do
  ( timeout 1200 python -m pytest tests/$f.py -m gpu -q 2>&1 | tail -40 ) > gpurun_out/${tag}_pytest_$f.log
  echo "$f: $(tail -1 gpurun_out/${tag}_pytest_$f.log)"
done
python profiles/prof_cfg.py slopes > gpurun_out/${tag}_slopes.txt 2>&1
cat gpurun_out/${tag}_slopes.txt
python profiles/prof_cfg.py eigh_time 2>&1 | grep "N=256\|N=128" > gpurun_out/${tag}_eigh_time.txt; cat gpurun_out/${tag}_eigh_time.txt
SP_EIGH_QUIET=1e-6 python profiles/prof_cfg.py cma_time 2>&1 | head -1
