"""Oracle: shared operators (reference: stochopy/optimize/_common.py:109-194)."""
import numpy as np

# _common.py:12-24
MESSAGES = {
    -8: "TolX",
    -7: "TolFun",
    -6: "TolXUp",
    -5: "EqualFunValues",
    -4: "ConditionCov",
    -3: "NoEffectCoord",
    -2: "NoEffectAxis",
    -1: "maximum number of iterations is reached",
    0: "best solution changes less than xtol",
    1: "best solution value is lower than ftol",
}


def lhs_from_draws(jitter, perms, bounds=None):
    """Latin hypercube from explicit draws, _common.py:109-120.

    Quirk kept: the jitter is 1/P wide while the strata are 2/P apart
    (linspace(-1,1,P,endpoint=False))."""
    P, N = jitter.shape
    dt = jitter.dtype
    cell = jitter / dt.type(P) + np.linspace(-1.0, 1.0, P, endpoint=False).astype(dt)[:, None]
    pop = np.take_along_axis(cell, perms, axis=0)
    if bounds is not None:
        lower, upper = np.transpose(np.asarray(bounds, dtype=dt))
        pop = pop * (dt.type(0.5) * (upper - lower))
        pop = pop + dt.type(0.5) * (upper + lower)
    return pop


def terminate_sync(it, maxiter, dist, bestfun, xtol, ftol):
    """Status ladder of selection_sync, _common.py:135-158 (None == keep going)."""
    if dist <= xtol and bestfun <= ftol:
        return 0
    if bestfun <= ftol:
        return 1
    if it >= maxiter:
        return -1
    return None


def select_sync(it, cand, candfun, xbest, x, xfun, maxiter, xtol, ftol):
    """Greedy synchronous selection with already evaluated candidates.

    _common.py:123-160: strict '<' replacement in place, np.argmin (first
    minimum) for the new best, status from both tolerances."""
    better = candfun < xfun
    xfun[better] = candfun[better]
    x[better] = cand[better]
    b = int(np.argmin(xfun))
    dist = np.linalg.norm(xbest - x[b])
    status = terminate_sync(it, maxiter, dist, xfun[b], xtol, ftol)
    return x[b].copy(), xfun[b], status


def select_async(cand_i, candfun_i, i, xbest, xbestfun, x, xfun, xtol, ftol):
    """One individual of selection_async, _common.py:163-194 ('<=' variants)."""
    status = None
    if candfun_i <= xfun[i]:
        x[i] = cand_i
        xfun[i] = candfun_i
        if candfun_i <= xbestfun:
            near = np.linalg.norm(xbest - cand_i) <= xtol
            low = candfun_i <= ftol
            if near and low:
                status = 0
            elif low:
                status = 1
            xbest = np.array(cand_i, copy=True)
            xbestfun = candfun_i
    return xbest, xbestfun, status


def result(x, fun, status, nfev, nit, xall=None, funall=None):
    res = dict(
        x=x,
        success=status >= 0,
        status=status,
        message=MESSAGES[status],
        fun=fun,
        nfev=nfev,
        nit=nit,
    )
    if xall is not None:
        res["xall"] = xall
        res["funall"] = funall
    return res


class History:
    """xall/funall bookkeeping shared by de/cpso/na (_de.py:221-234, 270-278)."""

    def __init__(self, enabled, maxiter, P, N, verbosity):
        self.enabled = enabled
        if enabled:
            self.nout = int(np.ceil(verbosity * P))
            w = max(1, self.nout)
            self.xall = np.empty((maxiter, w, N))
            self.funall = np.empty((maxiter, w))

    def first(self, X, pfit, gbest, gfit):
        if not self.enabled:
            return
        if self.nout > 0:
            self.xall[0] = X[: self.nout]
            self.funall[0] = pfit[: self.nout]
        else:
            self.xall[0] = gbest
            self.funall[0] = gfit

    def put(self, it, X, pfit):
        if not self.enabled:
            return
        if self.nout > 0:
            self.xall[it - 1] = X[: self.nout]
            self.funall[it - 1] = pfit[: self.nout]
        else:
            b = pfit.argmin()
            self.xall[it - 1] = X[b]
            self.funall[it - 1] = pfit[b]

    def upto(self, it):
        return (self.xall[:it], self.funall[:it]) if self.enabled else (None, None)
