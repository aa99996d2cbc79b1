#!/bin/bash
# round 2, session 6: full re-measurement after the container was re-created (earlier gpurun_out/ lost):
# GPU test files in separate processes, smoke, bench line (+configs), launch lists and full ncu captures of the
# DE / VD-CMA / CMA-ES / PSO kernels.
tag=r02s6
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
for f in test_gpu_parity test_gpu_es test_gpu_sizes test_gpu_jit test_gpu_l3 test_parallel; do
  ( timeout 1200 python -m pytest tests/$f.py -m gpu -q 2>&1 | tail -40 ) > gpurun_out/${tag}_pytest_$f.log
  echo "$f: $(tail -1 gpurun_out/${tag}_pytest_$f.log)"
done
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/${tag}_smoke.log
cat gpurun_out/${tag}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 4000 gpurun_out/${tag}_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
   python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/${tag}_launches_run.log 2>&1
for c in vd cma cpso pso; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_$c.csv \
     python profiles/prof_cfg.py $c > gpurun_out/${tag}_launches_$c.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:de_pool_kernel -s 10 -c 2 -f -o gpurun_out/${tag}_de_pool \
   python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/${tag}_ncu_de.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"vd_" -s 12 -c 8 -f -o gpurun_out/${tag}_vd \
   python profiles/prof_cfg.py vd > gpurun_out/${tag}_ncu_vd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cma_|jacobi" -s 24 -c 14 -f -o gpurun_out/${tag}_cma \
   python profiles/prof_cfg.py cma > gpurun_out/${tag}_ncu_cma.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pso_" -s 4 -c 4 -f -o gpurun_out/${tag}_cpso \
   python profiles/prof_cfg.py cpso > gpurun_out/${tag}_ncu_cpso.log 2>&1
ls -la gpurun_out
