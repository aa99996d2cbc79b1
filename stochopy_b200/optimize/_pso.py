"""PSO front-end: CPSO without the competitive restart (stochopy/optimize/pso/_pso.py:9-122)."""
from ._cpso import minimize as cpso
from ._helpers import register

from ._common import device_scope

__all__ = ["minimize"]


@device_scope
def minimize(
    fun,
    bounds,
    x0=None,
    args=(),
    maxiter=100,
    popsize=10,
    inertia=0.7298,
    cognitivity=1.49618,
    sociability=1.49618,
    seed=None,
    xtol=1.0e-8,
    ftol=1.0e-8,
    constraints=None,
    updating="immediate",
    workers=1,
    backend=None,
    return_all=False,
    verbosity=1.0,
    callback=None,
    dtype="float64",
    device=None,
    rng="philox",
):
    """Particle Swarm Optimization on the GPU; arguments as stochopy.optimize.pso.minimize."""
    return cpso(fun, bounds, x0, args, maxiter, popsize, inertia, cognitivity, sociability, None, seed, xtol, ftol,
                constraints, updating, workers, backend, return_all, verbosity, callback, dtype, device, rng)


register("pso", minimize)
