#!/bin/bash
# round 2, session 5: VD-CMA iteration 3 (single-barrier update, compile-time sampling variants, Philox-7 normals),
# e2e path, new bench.py; test files in separate processes; memcheck of the small ES configurations
tag=r02s5
mkdir -p gpurun_out
rm -f gpurun_out/size_parity.jsonl gpurun_out/l3_stats.jsonl
for f in test_gpu_sizes test_gpu_es test_gpu_jit test_gpu_l3 test_gpu_parity test_parallel; do
  ( timeout 900 python -m pytest tests/$f.py -m gpu -q 2>&1 | tail -40 ) > gpurun_out/${tag}_pytest_$f.log
  tail -1 gpurun_out/${tag}_pytest_$f.log
done
cat > /tmp/es_small.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import stochopy_b200 as sb
for m in ("vdcma", "cmaes"):
    r = sb.optimize.minimize(sb.factory.rosenbrock, [[-2.0, 2.0]] * 5, method=m, options=dict(maxiter=6, popsize=12, seed=3))
    print(m, "small", r.nit, r.fun)
    r = sb.optimize.minimize(sb.factory.rastrigin, [[-5.12, 5.12]] * 300, method=m, options=dict(maxiter=3, popsize=1000, seed=3, constraints="Penalize"))
    print(m, "300 penalize", r.nit, r.fun)
    r = sb.optimize.minimize(sb.factory.rosenbrock, [[-5.12, 5.12]] * 300, method=m, options=dict(maxiter=3, popsize=1000, seed=3, dtype="float32", return_all=True))
    print(m, "300 f32 return_all", r.nit, r.fun)
PY
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python /tmp/es_small.py > gpurun_out/${tag}_memcheck_es.log 2>&1
grep -E "ERROR SUMMARY|small|300" gpurun_out/${tag}_memcheck_es.log | head
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches_vd.csv \
   python profiles/prof_cfg.py vd > gpurun_out/${tag}_launches_vd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"vd_sample_eval|vd_wsum|vd_update" -s 3 -c 3 -f -o gpurun_out/${tag}_vd \
   python profiles/prof_cfg.py vd > gpurun_out/${tag}_ncu_vd.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 3000 gpurun_out/${tag}_bench.json
