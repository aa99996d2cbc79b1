"""Oracle: differential evolution (reference: stochopy/optimize/de/)."""
import numpy as np

from .common import History, lhs_from_draws, result, select_async, select_sync
from .objectives import evaluate
from .streams import MTStream

DONORS = {"rand1bin": 3, "rand2bin": 5, "best1bin": 2, "best2bin": 4}


def mutate(strategy, d, F, X, gbest):
    """de/_strategy.py:1-38.  ``d`` = donor index array(s), first k rows used."""
    if strategy == "rand1bin":
        return X[d[0]] + F * (X[d[1]] - X[d[2]])
    if strategy == "rand2bin":
        return X[d[0]] + F * (X[d[1]] + X[d[2]] - X[d[3]] - X[d[4]])
    if strategy == "best1bin":
        return gbest + F * (X[d[0]] - X[d[1]])
    if strategy == "best2bin":
        return gbest + F * (X[d[0]] + X[d[1]] - X[d[2]] - X[d[3]])
    raise KeyError(strategy)


def trial_population(X, gbest, strategy, F, CR, r1, donors, irand, repair, lower, upper):
    """Mutation + binomial crossover + bound repair of de_sync, _de.py:333-344.

    repair: None (NoConstraint) or the (P,N) uniform(lower,upper) matrix that
    de/_constraints.py:22-26 draws every generation."""
    P, N = X.shape
    V = mutate(strategy, donors, F, X, gbest)
    forced = np.zeros((P, N), dtype=bool)
    forced[np.arange(P), irand] = True
    U = np.where(forced | (r1 <= CR), V, X)
    if repair is not None:
        U = np.where((U < lower) | (U > upper), repair, U)
    return U


def generation_sync(it, X, gbest, pbestfit, U, candfun, maxiter, xtol, ftol):
    """Selection half of de_sync (_de.py:346-351): mutates X, pbestfit in place.
    Returns gbest, gfit, status; the generation's pfit is ``candfun``."""
    return select_sync(it, U, candfun, gbest, X, pbestfit, maxiter, xtol, ftol)


def minimize(
    fun,
    bounds,
    x0=None,
    maxiter=100,
    popsize=10,
    mutation=0.5,
    recombination=0.9,
    strategy="best1bin",
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
    """Driver of _de.py:176-301 (it starts at 1; nfev = it*P)."""
    dt = np.dtype(dtype)
    bounds = np.asarray(bounds, dtype=dt)
    N = len(bounds)
    P = popsize
    lower, upper = bounds.T
    k = DONORS[strategy]
    if constraints not in (None, "Random"):
        raise KeyError(constraints)
    repair = constraints == "Random"
    stream = stream if stream is not None else MTStream(seed)
    F, CR = dt.type(mutation), dt.type(recombination)

    if x0 is not None:
        X = np.array(x0, dtype=dt)
    else:
        X = lhs_from_draws(*stream.lhs(P, N), bounds).astype(dt)
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
        if updating == "deferred":
            r1, donors, irand, rep = stream.de(it, P, N, k, lower, upper, repair)
            U = trial_population(X, gbest, strategy, F, CR, r1, donors, irand, rep, lower, upper)
            pfit = evaluate(fun, U).astype(dt)
            gbest, gfit, status = generation_sync(it, X, gbest, pbestfit, U, pfit, maxiter, xtol, ftol)
        else:  # _de.py:354-394; status of the *last* individual survives the loop
            r1 = stream.de_rand(P, N)
            for i in range(P):
                donors, irand = stream.de_one(i, P, N, k, lower, upper, repair)
                v = mutate(strategy, donors, F, X, gbest)
                forced = np.zeros(N, dtype=bool)
                forced[irand] = True
                u = np.where(forced | (r1[i] <= CR), v, X[i])
                if repair:
                    u = np.where((u < lower) | (u > upper), stream.repair_one(lower, upper, N), u)
                pfit[i] = fun(u)
                gbest, gfit, status = select_async(u, pfit[i], i, gbest, gfit, X, pbestfit, xtol, ftol)
            if status is None and it >= maxiter:
                status = -1
        hist.put(it, X, pfit)
        if callback is not None:
            callback(X, dict(x=gbest, fun=gfit, nfev=it * P, nit=it))

    return result(gbest, gfit, status, it * P, it, *hist.upto(it))
