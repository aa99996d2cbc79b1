#!/bin/bash
# round 2, session 17: clean rebuild of every object (the per-object dependency files were missing for the units
# built before the Makefile change): full GPU suite, smoke, bench, slopes, eigensolver sweeps, memcheck log,
# launch list + ncu summary of the headline kernel
tag=r02s17
mkdir -p gpurun_out
for f in test_gpu_parity test_gpu_es test_gpu_sizes test_gpu_jit test_gpu_l3 test_parallel; do
  ( timeout 1200 python -m pytest tests/$f.py -m gpu -q 2>&1 | tail -40 ) > gpurun_out/${tag}_pytest_$f.log
  echo "$f: $(tail -1 gpurun_out/${tag}_pytest_$f.log)"
done
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > gpurun_out/${tag}_smoke.log; cat gpurun_out/${tag}_smoke.log
python profiles/prof_cfg.py eigh_time 2>&1 | grep "N=256\|N=128" > gpurun_out/${tag}_eigh_time.txt; cat gpurun_out/${tag}_eigh_time.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02s17_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "value_dirty_flush", "value_l2_resident", "value_kernel_only")})
print(d["e2e"]["value"], d["roofline"]["frac"], {k: round(v["us_per_generation"], 1) for k, v in d.get("configs", {}).items()})
print(d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("kind"))
PY
timeout 600 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
tail -c 400 gpurun_out/${tag}_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
   python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/${tag}_launches_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:de_pool_kernel -s 10 -c 2 -f -o /tmp/${tag}_de_pool \
   python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/${tag}_ncu_de.log 2>&1
python profiles/summarize_ncu.py /tmp/${tag}_de_pool.ncu-rep gpurun_out/${tag}_de_pool_ncu_summary.json 65536 >> gpurun_out/${tag}_ncu_de.log 2>&1
for c in cma cpso; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_$c.csv \
     python profiles/prof_cfg.py $c > /dev/null 2>&1
done
( timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python profiles/sanitize_small.py 2>&1 | tail -60 ) > gpurun_out/${tag}_memcheck.log
grep -E "ERROR SUMMARY" gpurun_out/${tag}_memcheck.log
( timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python profiles/sanitize_small.py 2>&1 | tail -80 ) > gpurun_out/${tag}_racecheck.log
grep -E "RACECHECK SUMMARY" gpurun_out/${tag}_racecheck.log
du -sh gpurun_out
