"""In-run (warm, not serialised) kernel durations of one configuration through torch.profiler / CUPTI:
per-kernel average device time, the sum over a generation, and the wall time per generation beside it -- what is
left between the two is launch gaps and host time.
    python profiles/prof_timeline.py cpso|pso|vd|cma|de [generations]"""
import collections
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import torch
from torch.profiler import ProfilerActivity, profile

import stochopy_b200 as sb

which = sys.argv[1]
gens = int(sys.argv[2]) if len(sys.argv) > 2 else 200
off = dict(xtol=-1.0, ftol=-1.0e300)
B = 5.12
cfg = {
    "de": (sb.factory.rosenbrock, 128, "de", dict(popsize=65536, dtype="float32", strategy="best1bin", updating="deferred")),
    "pso": (sb.factory.styblinski_tang, 64, "pso", dict(popsize=32768, dtype="float32", updating="deferred")),
    "cpso": (sb.factory.styblinski_tang, 64, "cpso", dict(popsize=32768, dtype="float32", updating="deferred", competitivity=1.0)),
    "vd": (sb.factory.ackley, 1024, "vdcma", dict(popsize=16384, dtype="float32")),
    "cma": (sb.factory.rosenbrock, 256, "cmaes", dict(popsize=4096)),
}[which]
fun, n, method, o = cfg


def run(it):
    return sb.optimize.minimize(fun, [[-B, B]] * n, method=method, options=dict(o, maxiter=it, seed=0, **off))


run(10)
torch.cuda.synchronize()
t0 = time.perf_counter()
run(gens)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(gens)
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type is not None and "cuda" in str(ev.device_type).lower() and ev.device_time_total > 0:
        k = ev.name.split("(")[0][:70]
        tot[k][0] += 1
        tot[k][1] += ev.device_time_total
busy = sum(v[1] for v in tot.values())
print(f"{which}: {gens} generations, wall {wall / gens * 1e6:.1f} us/gen (unprofiled run), device busy {busy / gens:.1f} us/gen")
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"  {c:6d} x {t / c:8.2f} us  = {t / gens:7.2f} us/gen  {k}")
