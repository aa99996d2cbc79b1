import sys, os
sys.path.insert(0, os.getcwd())
import stochopy_b200 as sb
which = sys.argv[1]
off = dict(xtol=-1.0, ftol=-1.0e300)
if which == "vd":
    sb.optimize.minimize(sb.factory.ackley, [[-5.12, 5.12]] * 1024, method="vdcma", options=dict(maxiter=4, popsize=16384, seed=0, dtype="float32", **off))
elif which == "cma":
    sb.optimize.minimize(sb.factory.rosenbrock, [[-5.12, 5.12]] * 256, method="cmaes", options=dict(maxiter=4, popsize=4096, seed=0, **off))
elif which == "pso":
    sb.optimize.minimize(sb.factory.styblinski_tang, [[-5.12, 5.12]] * 64, method="pso", options=dict(maxiter=6, popsize=32768, seed=0, dtype="float32", updating="deferred", **off))
