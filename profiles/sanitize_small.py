"""Small end-to-end runs of every method for compute-sanitizer (memcheck / racecheck):
chained DE / PSO generations (parity-indexed scratch regions, self-resetting done_blocks, PDL
prologues), CPSO lazy + eager restarts, the ES kernel chains, NA, return_all streaming."""
import os
import sys

sys.path.insert(0, os.getcwd())
import numpy as np  # noqa: E402

import stochopy_b200 as sb  # noqa: E402

off = dict(xtol=-1.0, ftol=-1.0e300)
b = lambda n: [[-5.12, 5.12]] * n  # noqa: E731
for dt in ("float32", "float64"):
    r = sb.optimize.minimize(sb.factory.rosenbrock, b(128), method="de",
                             options=dict(maxiter=12, popsize=1500, seed=1, dtype=dt, updating="deferred", strategy="best1bin", **off))
    print("de pool", dt, r.nit, r.fun)
    r = sb.optimize.minimize(sb.factory.rastrigin, b(300), method="de",
                             options=dict(maxiter=6, popsize=600, seed=1, dtype=dt, updating="deferred", strategy="rand2bin",
                                          constraints="Random", **off))
    print("de ring", dt, r.nit, r.fun)
    r = sb.optimize.minimize(sb.factory.sphere, b(6), method="de",
                             options=dict(maxiter=20, popsize=40, seed=1, dtype=dt, updating="deferred", return_all=True))
    print("de small + history", dt, r.nit, r.fun)
    r = sb.optimize.minimize(sb.factory.styblinski_tang, b(64), method="pso",
                             options=dict(maxiter=12, popsize=3001, seed=1, dtype=dt, updating="deferred", **off))
    print("pso chained", dt, r.nit, r.fun)
    r = sb.optimize.minimize(sb.factory.rastrigin, b(10), method="cpso",
                             options=dict(maxiter=30, popsize=500, seed=8, dtype=dt, competitivity=1.0, updating="deferred", **off))
    print("cpso restarts", dt, r.nit, r.fun)
    r = sb.optimize.minimize(sb.factory.rastrigin, b(6), method="cpso",
                             options=dict(maxiter=120, popsize=300, seed=8, dtype=dt, competitivity=1.0, updating="deferred", **off))
    print("cpso lazy", dt, r.nit, r.fun)
    for m in ("cmaes", "vdcma"):
        r = sb.optimize.minimize(sb.factory.rosenbrock, b(5), method=m, options=dict(maxiter=8, popsize=12, seed=3, dtype=dt))
        print(m, "small", dt, r.nit, r.fun)
        r = sb.optimize.minimize(sb.factory.rastrigin, b(130), method=m,
                                 options=dict(maxiter=3, popsize=600, seed=3, dtype=dt, constraints="Penalize"))
        print(m, "130 penalize", dt, r.nit, r.fun)
    r = sb.optimize.minimize(sb.factory.sphere, b(4), method="na", options=dict(maxiter=6, popsize=16, seed=3, dtype=dt))
    print("na", dt, r.nit, r.fun)
r = sb.optimize.minimize(lambda x: float(np.sum(x * x)), b(5), method="de", options=dict(maxiter=5, popsize=16, seed=3, updating="deferred"))
print("host objective", r.nit, r.fun)
