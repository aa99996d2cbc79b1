#!/bin/bash
# round 2, session 8: tests after the VD-CMA injection fix, launch lists + small ncu captures of the VD-CMA /
# CMA-ES / PSO / CPSO chains, compute-sanitizer memcheck + racecheck over small runs of every method.
tag=r02s8
mkdir -p gpurun_out
for f in test_gpu_parity test_gpu_es test_gpu_sizes; do
  ( timeout 1200 python -m pytest tests/$f.py -m gpu -q 2>&1 | tail -60 ) > gpurun_out/${tag}_pytest_$f.log
  echo "$f: $(tail -1 gpurun_out/${tag}_pytest_$f.log)"
done
for c in vd cma cpso pso; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_$c.csv \
     python profiles/prof_cfg.py $c > gpurun_out/${tag}_launches_$c.log 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vd_sample_eval -s 2 -c 1 -f -o gpurun_out/${tag}_vd_sample \
   python profiles/prof_cfg.py vd > gpurun_out/${tag}_ncu_vd.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"vd_wsum|vd_update|rank_" -s 6 -c 4 -f -o gpurun_out/${tag}_vd_rest \
   python profiles/prof_cfg.py vd >> gpurun_out/${tag}_ncu_vd.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"cma_" -s 12 -c 8 -f -o gpurun_out/${tag}_cma \
   python profiles/prof_cfg.py cma > gpurun_out/${tag}_ncu_cma.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"pso_generation" -s 3 -c 1 -f -o gpurun_out/${tag}_pso \
   python profiles/prof_cfg.py pso > gpurun_out/${tag}_ncu_pso.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 10 python profiles/sanitize_small.py > gpurun_out/${tag}_memcheck.log 2>&1
grep -E "ERROR SUMMARY" gpurun_out/${tag}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python profiles/sanitize_small.py > gpurun_out/${tag}_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|ERROR SUMMARY" gpurun_out/${tag}_racecheck.log
du -sh gpurun_out
