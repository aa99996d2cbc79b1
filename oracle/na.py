"""Oracle: Neighbourhood Algorithm (reference: stochopy/optimize/na/_na.py)."""
import numpy as np

from .common import History, lhs_from_draws, result, select_sync
from .objectives import evaluate
from .streams import MTStream


def voronoi_limits(archive, k, walker, j, d1, d2):
    """Bounds of cell k along axis j through ``walker`` (_na.py:290-297).
    d1: squared distance of the walker to cell centre k over the other axes;
    d2: same to every other archived model.  Returns (low, high) in [0,1]."""
    others = np.delete(archive, k, axis=0)
    ck = archive[k, j]
    lim = 0.5 * (ck + others[:, j] + (d1 - d2) / (ck - others[:, j]))
    below = lim <= walker[j]
    above = lim >= walker[j]
    low = max(lim[below].max(), 0.0) if below.sum() else 0.0
    high = min(lim[above].min(), 1.0) if above.sum() else 1.0
    return low, high


def resample(archive, archfit, P, N, nr, span_mask, draw):
    """Gibbs walk inside the Voronoi cells of the best nr archived models
    (_na.py:265-305).  ``draw(i, j, low, high)`` supplies uniform(low, high)."""
    X = np.empty((P, N))
    best = archfit.argsort()[:nr]
    for i in range(P):
        k = best[i % nr]
        X[i] = archive[k]
        others = np.delete(archive, k, axis=0)
        d1 = 0.0
        d2 = ((others[:, 1:] - X[i, 1:]) ** 2).sum(axis=1)
        for j in range(N):
            if not span_mask[j]:
                X[i, j] = 0.0
                continue
            low, high = voronoi_limits(archive, k, X[i], j, d1, d2)
            X[i, j] = draw(i, j, low, high)
            if j < N - 1:
                d1 += (archive[k, j] - X[i, j]) ** 2 - (archive[k, j + 1] - X[i, j + 1]) ** 2
                d2 += (others[:, j] - X[i, j]) ** 2 - (others[:, j + 1] - X[i, j + 1]) ** 2
    return X


def minimize(
    fun,
    bounds,
    x0=None,
    maxiter=100,
    popsize=10,
    nrperc=0.5,
    seed=None,
    xtol=1.0e-8,
    ftol=1.0e-8,
    return_all=False,
    verbosity=1.0,
    callback=None,
    stream=None,
):
    """Driver of _na.py:134-262 (unit-cube normalisation, growing archive)."""
    bounds = np.asarray(bounds, dtype=np.float64)
    N, P = len(bounds), popsize
    lower, upper = bounds.T
    span = upper - lower
    span_mask = span > 0.0
    span = np.where(span_mask, span, 1.0)
    norm = lambda x: np.where(span_mask, (x - lower) / span, upper)
    unnorm = lambda x: np.where(span_mask, x * span + lower, upper)
    evaluate_n = lambda X: evaluate(fun, unnorm(X))
    stream = stream if stream is not None else MTStream(seed)
    nr = max(1, int(nrperc * P))

    X = np.array(x0, dtype=np.float64) if x0 is not None else lhs_from_draws(*stream.lhs(P, N), bounds)
    X = norm(X)
    pbest = X.copy()
    pfit = evaluate_n(X)
    pbestfit = pfit.copy()
    b = int(np.argmin(pbestfit))
    gfit, gbest = pbestfit[b], X[b].copy()
    archive, archfit = X.copy(), pfit.copy()

    # quirk kept: xall[0] holds the *normalised* initial population (_na.py:186-187)
    hist = History(return_all, maxiter, P, N, verbosity)
    hist.first(X, pfit, gbest, gfit)
    if callback is not None:
        callback(unnorm(X), dict(x=unnorm(gbest), fun=gfit, nfev=P, nit=1))

    it = 1
    status = None
    while status is None:
        it += 1
        draw = lambda i, j, lo, hi: stream.na_uniform(it, i, j, lo, hi)
        X = resample(archive, archfit, P, N, nr, span_mask, draw)
        pfit = evaluate_n(X)
        gbest, gfit, status = select_sync(it, X, pfit, gbest, pbest, pbestfit, maxiter, xtol, ftol)
        archive = np.vstack((X, archive))
        archfit = np.concatenate((pfit, archfit))
        hist.put(it, unnorm(X), pfit)
        if callback is not None:
            callback(unnorm(X), dict(x=unnorm(gbest), fun=gfit, nfev=it * P, nit=it))

    return result(unnorm(gbest), gfit, status, it * P, it, *hist.upto(it))
