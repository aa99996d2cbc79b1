"""Oracle: PSO / competitive PSO (reference: stochopy/optimize/cpso/, pso/)."""
import numpy as np

from .common import History, lhs_from_draws, result, select_async, select_sync
from .objectives import evaluate
from .streams import MTStream


def shrink_factor(x, v, lower, upper):
    """cpso/_constraints.py:22-42: largest beta<=... that keeps x+beta*v inside;
    min over violated coordinates of (bound-x)/v, 1 when nothing is violated."""
    trial = x + v
    lo = trial < lower
    hi = trial > upper
    beta = None
    if lo.any():
        beta = ((lower[lo] - x[lo]) / v[lo]).min()
    if hi.any():
        bu = ((upper[hi] - x[hi]) / v[hi]).min()
        beta = bu if beta is None else min(beta, bu)
    return 1.0 if beta is None else beta


def move(X, V, pbest, gbest, w, c1, c2, r1, r2, constraints, lower, upper):
    """Velocity + position update, _cpso.py:324-329 with cpso/_constraints.py.
    Returns new (X, V); works on a population (2-D) or one particle (1-D)."""
    V = w * V + c1 * r1 * (pbest - X) + c2 * r2 * (gbest - X)
    if constraints == "Shrink":
        if X.ndim == 2:
            beta = np.array([shrink_factor(x, v, lower, upper) for x, v in zip(X, V)])
            V = V * beta[:, None]
        else:
            V = V * shrink_factor(X, V, lower, upper)
    elif constraints is not None:
        raise KeyError(constraints)
    return X + V, V


def restart_plan(it, X, gbest, pbestfit, gamma, delta, maxiter):
    """Competitive restart decision, _cpso.py:405-420.  Returns the rows to
    reset (worst first: pbestfit.argsort()[:-nw-1:-1]) or an empty array."""
    P, N = X.shape
    radius = max(np.linalg.norm(X[i] - gbest) for i in range(P)) / np.sqrt(4.0 * N)
    if radius < delta:
        nw = int((P - 1.0) / (1.0 + np.exp(1.0 / 0.09 * (it / maxiter - gamma + 0.5))))
        if nw > 0:
            return pbestfit.argsort()[: -nw - 1 : -1]
    return np.empty(0, dtype=np.int64)


def restart_apply(rows, fresh, X, V, pbest, pbestfit):
    """_cpso.py:421-424."""
    V[rows] = 0.0
    X[rows] = fresh
    pbest[rows] = fresh
    pbestfit[rows] = 1.0e30


def swarm_delta(P, maxiter):
    """_cpso.py:216."""
    return np.log(1.0 + 0.003 * P) / np.max((0.2, np.log(0.01 * maxiter)))


def minimize(
    fun,
    bounds,
    x0=None,
    maxiter=100,
    popsize=10,
    inertia=0.7298,
    cognitivity=1.49618,
    sociability=1.49618,
    competitivity=1.0,
    seed=None,
    xtol=1.0e-8,
    ftol=1.0e-8,
    constraints=None,
    updating="immediate",
    return_all=False,
    verbosity=1.0,
    callback=None,
    stream=None,
    dtype=np.float64,
):
    """Driver of _cpso.py:182-321.  competitivity=None gives plain PSO (_pso.py:99)."""
    dt = np.dtype(dtype)
    bounds = np.asarray(bounds, dtype=dt)
    N = len(bounds)
    P = popsize
    lower, upper = bounds.T
    stream = stream if stream is not None else MTStream(seed)
    w, c1, c2 = dt.type(inertia), dt.type(cognitivity), dt.type(sociability)
    gamma = competitivity
    if gamma:
        delta = swarm_delta(P, maxiter)

    if x0 is not None:
        X = np.array(x0, dtype=dt)
    else:
        X = lhs_from_draws(*stream.lhs(P, N), bounds).astype(dt)
    V = np.zeros((P, N), dtype=dt)
    pbest = X.copy()
    pfit = evaluate(fun, X).astype(dt)
    pbestfit = pfit.copy()
    b = int(np.argmin(pbestfit))
    gfit, gbest = pbestfit[b], X[b].copy()

    hist = History(return_all, maxiter, P, N, verbosity)
    hist.first(X, pfit, gbest, gfit)
    if callback is not None:
        callback(X, dict(x=gbest, fun=gfit, nfev=P, nit=1))

    it = 1
    status = None
    while status is None:
        it += 1
        r1, r2 = stream.pso(it, P, N)
        if updating == "deferred":
            X, V = move(X, V, pbest, gbest, w, c1, c2, r1, r2, constraints, lower, upper)
            pfit = evaluate(fun, X).astype(dt)
            gbest, gfit, status = select_sync(it, X, pfit, gbest, pbest, pbestfit, maxiter, xtol, ftol)
        else:  # _cpso.py:364-402
            for i in range(P):
                X[i], V[i] = move(
                    X[i], V[i], pbest[i], gbest, w, c1, c2, r1[i], r2[i], constraints, lower, upper
                )
                pfit[i] = fun(X[i])
                gbest, gfit, status = select_async(X[i], pfit[i], i, gbest, gfit, pbest, pbestfit, xtol, ftol)
            if status is None and it >= maxiter:
                status = -1
        hist.put(it, X, pfit)
        if callback is not None:
            callback(X, dict(x=gbest, fun=gfit, nfev=it * P, nit=it))
        if status is None and gamma:
            rows = restart_plan(it, X, gbest, pbestfit, gamma, delta, maxiter)
            if len(rows):
                fresh = stream.pso_restart(it, rows, N, lower, upper)
                restart_apply(rows, fresh, X, V, pbest, pbestfit)

    return result(gbest, gfit, status, it * P, it, *hist.upto(it))
