#!/bin/bash
# round 2, session 7: DE kernel v10 (Philox2x32 16-bit crossover pieces, pair claims, record table, compile-time
# objective): parity, bench, launch list, one full ncu capture; logs of the failing ES/size tests.
# gpurun_out must stay below 64 MiB: few captured launches per report.
tag=r02s7
mkdir -p gpurun_out
for f in test_gpu_parity test_gpu_es test_gpu_sizes test_gpu_l3; do
  ( timeout 1200 python -m pytest tests/$f.py -m gpu -q 2>&1 | tail -120 ) > gpurun_out/${tag}_pytest_$f.log
  echo "$f: $(tail -1 gpurun_out/${tag}_pytest_$f.log)"
done
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/${tag}_smoke.log
cat gpurun_out/${tag}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 1500 gpurun_out/${tag}_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02s7_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "value_dirty_flush", "value_l2_resident", "value_kernel_only", "event_pair_overhead_us")})
print(d["e2e"]["value"], d["roofline"]["frac"], {k: v["us_per_generation"] for k, v in d.get("configs", {}).items()})
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
   python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/${tag}_launches_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:de_pool_kernel -s 10 -c 2 -f -o gpurun_out/${tag}_de_pool \
   python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/${tag}_ncu_de.log 2>&1
SP_DE_GENERIC_OBJ=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras > gpurun_out/${tag}_bench_generic_obj.json 2>> gpurun_out/${tag}_bench.err
tail -c 600 gpurun_out/${tag}_bench_generic_obj.json
du -sh gpurun_out
