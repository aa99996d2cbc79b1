#!/bin/bash
# round 2, session 1: new parity tests (sizes, L3), DGEMM probe, reference arm, launch lists of the ES chains
tag=r02s1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
nproc > gpurun_out/${tag}_nproc.txt
( timeout 1500 python -m pytest tests/test_gpu_sizes.py tests/test_gpu_l3.py -m gpu -q 2>&1 | tail -60 ) > gpurun_out/${tag}_pytest_new.log
( timeout 300 python profiles/dgemm_probe.py gpurun_out/${tag}_dgemm.json 2>&1 | tail -5 ) > gpurun_out/${tag}_dgemm.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
for w in vd cma cpso; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches_${w}.csv \
   python profiles/prof_cfg.py $w > gpurun_out/${tag}_launches_${w}.log 2>&1
done
timeout 600 python profiles/prof_cfg.py slopes > gpurun_out/${tag}_slopes.txt 2>&1
tail -5 gpurun_out/${tag}_pytest_new.log; cat gpurun_out/${tag}_dgemm.log | tail -3; cat gpurun_out/${tag}_slopes.txt
