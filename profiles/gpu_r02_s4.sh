#!/bin/bash
# round 2, session 4: memcheck of the VD-CMA chain on the configurations that faulted
tag=r02s4
mkdir -p gpurun_out
cat > /tmp/vd_small.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import stochopy_b200 as sb
r = sb.optimize.minimize(sb.factory.rosenbrock, [[-2.0, 2.0]] * 5, method="vdcma", options=dict(maxiter=6, popsize=12, seed=3))
print("small", r.nit, r.fun)
r = sb.optimize.minimize(lambda x: float(np.sum(x * x)), [[-2.0, 2.0]] * 5, method="vdcma", options=dict(maxiter=4, popsize=12, seed=3))
print("host", r.nit, r.fun)
r = sb.optimize.minimize(sb.factory.rosenbrock, [[-5.12, 5.12]] * 300, method="vdcma", options=dict(maxiter=3, popsize=1000, seed=3, _probe=lambda it, b, c: None))
print("300", r.nit, r.fun)
PY
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/vd_small.py > gpurun_out/${tag}_memcheck_vd.log 2>&1
tail -60 gpurun_out/${tag}_memcheck_vd.log
