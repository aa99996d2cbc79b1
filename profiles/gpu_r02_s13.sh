#!/bin/bash
# round 2, session 13 (2 GPUs): CMA-ES after the Jacobi threshold change (1 GPU part), then the driver's multi-GPU
# command at N = 2: headline + c3_sharded + c5_seeds keys, and the multi-GPU pytest
tag=r02s13
mkdir -p gpurun_out
for f in test_gpu_es test_gpu_sizes; do
  ( timeout 1200 python -m pytest tests/$f.py -m gpu -q 2>&1 | tail -40 ) > gpurun_out/${tag}_pytest_$f.log
  echo "$f: $(tail -1 gpurun_out/${tag}_pytest_$f.log)"
done
python profiles/prof_cfg.py cma_time > gpurun_out/${tag}_cma_time.txt 2>&1; cat gpurun_out/${tag}_cma_time.txt
python profiles/prof_cfg.py eigh_time > gpurun_out/${tag}_eigh_time.txt 2>&1; grep "N=256" gpurun_out/${tag}_eigh_time.txt
( timeout 900 python -m pytest tests/test_parallel.py -m gpu -q 2>&1 | tail -15 ) > gpurun_out/${tag}_pytest_parallel_2gpu.log
tail -3 gpurun_out/${tag}_pytest_parallel_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_2gpu.json 2> gpurun_out/${tag}_bench_2gpu.err
tail -c 2500 gpurun_out/${tag}_bench_2gpu.json; tail -5 gpurun_out/${tag}_bench_2gpu.err
du -sh gpurun_out
