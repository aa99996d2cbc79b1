#!/bin/bash
# round 2, session 9: PSO kernel v10 (PLAIN / compile-time objective instantiations, r1+r2 from one Philox call);
# launch lists + ncu summaries of the VD-CMA / CMA-ES / PSO / CPSO chains (reports are summarised ON the box and
# deleted: gpurun_out must stay below 64 MiB), racecheck log.
tag=r02s9
mkdir -p gpurun_out
for f in test_gpu_parity test_gpu_l3 test_parallel; do
  ( timeout 1200 python -m pytest tests/$f.py -m gpu -q 2>&1 | tail -60 ) > gpurun_out/${tag}_pytest_$f.log
  echo "$f: $(tail -1 gpurun_out/${tag}_pytest_$f.log)"
done
for c in vd cma cpso pso; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_$c.csv \
     python profiles/prof_cfg.py $c > gpurun_out/${tag}_launches_$c.log 2>&1
done
cap() {  # cap <name> <kernel regex> <skip> <count> <units> <prof_cfg arg>
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/${tag}_$1 \
     python profiles/prof_cfg.py $6 > gpurun_out/${tag}_ncu_$1.log 2>&1
  python profiles/summarize_ncu.py /tmp/${tag}_$1.ncu-rep gpurun_out/${tag}_$1_ncu_summary.json $5 >> gpurun_out/${tag}_ncu_$1.log 2>&1
  rm -f /tmp/${tag}_$1.ncu-rep
}
cap vd_sample vd_sample_eval 2 1 16384 vd
cap vd_wsum vd_wsum 2 1 16384 vd
cap vd_update vd_update 2 1 16384 vd
cap rank "rank_" 4 2 16384 vd
cap pso pso_generation 3 1 32768 pso
cap cma_sample cma_sample 2 1 4096 cma
cap cma_cov cma_cov 2 2 4096 cma
cap jacobi jacobi 2 1 4096 cma
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02s9_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "value_dirty_flush", "value_l2_resident")})
print(d["e2e"]["value"], d["roofline"]["frac"], {k: v["us_per_generation"] for k, v in d.get("configs", {}).items()})
PY
( timeout 900 compute-sanitizer --tool racecheck --print-limit 60 python profiles/sanitize_small.py 2>&1 | tail -400 ) > gpurun_out/${tag}_racecheck.log
grep -E "RACECHECK SUMMARY|ERROR SUMMARY" gpurun_out/${tag}_racecheck.log
du -sh gpurun_out
