#!/bin/bash
# round 2, session 27: seeds on persistent worker streams -- sequential vs 2 / 4 at a time (1 GPU)
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_parallel.py -m gpu -q -k "concurrent" 2>&1 | tail -2 )
python - <<'PY' 2>&1 | grep -v -i warn | tee gpurun_out/r02s27_seeds_concurrent.txt
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import stochopy_b200 as sb
from stochopy_b200 import parallel
b = [[-5.12, 5.12]] * 1024
o = dict(popsize=16384, dtype="float32", xtol=-1.0, ftol=-1.0e300)
seeds = list(range(8))
for conc in (1, 2, 4, 1):
    parallel.minimize_seeds(sb.factory.ackley, b, list(range(max(conc, 1))), method="vdcma", options=dict(o, maxiter=3), concurrent=conc)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = parallel.minimize_seeds(sb.factory.ackley, b, seeds, method="vdcma", options=dict(o, maxiter=100), concurrent=conc)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"vdcma C5, 8 seeds x 100 generations, {conc} at a time: {dt*1e3:.1f} ms, {8*100*16384/dt:.3e} evals/s, best {r['fun']:.6f}", flush=True)
o = dict(popsize=65536, dtype="float32", updating="deferred", xtol=-1.0, ftol=-1.0e300)
for conc in (1, 4):
    parallel.minimize_seeds(sb.factory.rosenbrock, [[-5.12, 5.12]] * 128, list(range(conc)), method="de", options=dict(o, maxiter=5), concurrent=conc)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = parallel.minimize_seeds(sb.factory.rosenbrock, [[-5.12, 5.12]] * 128, seeds, method="de", options=dict(o, maxiter=300), concurrent=conc)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"de headline config, 8 seeds x 300 generations, {conc} at a time: {dt*1e3:.1f} ms, {8*299*65536/dt:.3e} evals/s", flush=True)
PY
