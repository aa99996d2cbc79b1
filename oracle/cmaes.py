"""Oracle: (mu,lambda)-CMA-ES (reference: stochopy/optimize/cmaes/)."""
import numpy as np

from .common import result
from .objectives import evaluate
from .streams import MTStream


def selection_weights(P, muperc):
    """_cmaes.py:184-189: mu, log weights, mueff."""
    mu = int(muperc * P)
    w = np.log(mu + 0.5) - np.log(np.arange(1, mu + 1))
    w /= w.sum()
    mueff = w.sum() ** 2 / np.square(w).sum()
    return mu, w, mueff


def strategy_constants(N, mueff):
    """_cmaes.py:192-205."""
    cc = (4.0 + mueff / N) / (N + 4.0 + 2.0 * mueff / N)
    cs = (mueff + 2.0) / (N + mueff + 5.0)
    c1 = 2.0 / ((N + 1.3) ** 2 + mueff)
    cmu = min(1.0 - c1, 2.0 * (mueff - 2.0 + 1.0 / mueff) / ((N + 2.0) ** 2 + mueff))
    damps = 1.0 + 2.0 * max(0.0, np.sqrt((mueff - 1.0) / (N + 1.0)) - 1.0) + cs
    chind = np.sqrt(N) * (1.0 - 1.0 / (4.0 * N) + 1.0 / (21.0 * N**2))
    return cc, cs, c1, cmu, damps, chind


def sample(xmean, sigma, B, D, Z):
    """_cmaes.py:232-237: arx_i = xmean + sigma * B (D o z_i)."""
    return xmean + sigma * (Z * D) @ B.T


def update(arx, order, mu, w, xmean, sigma, ps, pc, C, invsqrtC, consts, mueff, nfev, P):
    """Mean, paths, covariance and step size, _cmaes.py:272-298.
    Returns xmean, xold, ps, pc, C, sigma, hsig."""
    cc, cs, c1, cmu, damps, chind = consts
    N = xmean.size
    elite = arx[order[:mu]]
    xold = xmean.copy()
    xmean = w @ elite
    ps =(1.0 - cs) * ps + np.sqrt(cs * (2.0 - cs) * mueff) * (invsqrtC @ (xmean - xold)) / sigma
    hsig = np.linalg.norm(ps) / np.sqrt(1.0 - (1.0 - cs) ** (2.0 * nfev / P)) / chind < 1.4 + 2.0 / (N + 1.0)
    pc = (1.0 - cc) * pc
    if hsig:
        pc = pc + np.sqrt(cc * (2.0 - cc) * mueff) * (xmean - xold) / sigma
    art = (elite - xold) / sigma
    extra = 0.0 if hsig else c1 * cc * (2.0 - cc) * C
    C = (1.0 - c1 - cmu) * C
    C = C + cmu * (art.T * w) @ art
    C = C + c1 * np.outer(pc, pc)
    C = C + extra
    sigma = sigma * np.exp((cs / damps) * (np.linalg.norm(ps) / chind - 1.0))
    return xmean, xold, ps, pc, C, sigma, hsig


def decompose(C, eigh=np.linalg.eigh):
    """_cmaes.py:303-309: symmetrise from the upper triangle, eigh, ascending,
    D = sqrt(eigenvalues) (no negative guard), invsqrtC = B diag(1/D) B^T."""
    C = np.triu(C) + np.triu(C, 1).T
    vals, B = eigh(C)
    o = np.argsort(vals)
    vals, B = vals[o], B[:, o]
    D = np.sqrt(vals)
    return C, B, D, (B / D) @ B.T


def converge(it, N, maxiter, xmean, xold, besthist, arfit, order, sigma, insigma, ilim, pc,
             xtol, ftol, diagC, B=None, D=None):
    """Termination ladder, _cmaes.py:360-434 (precedence kept; zero-padded
    history quirks kept: the -5 slice includes a not-yet-written zero and -7
    appends the whole zero-padded history)."""
    i = int(np.floor(np.mod(it, N)))
    sd = np.sqrt(diagC)
    best = arfit[order[0]]
    if it >= maxiter:
        return -1
    if np.linalg.norm(xold - xmean) <= xtol and best < ftol:
        return 0
    if best <= ftol:
        return 1
    if B is not None and (np.abs(0.1 * sigma * B[:, i] * D[i]) < 1.0e-10).all():
        return -2
    if (0.2 * sigma * sd < 1.0e-10).any():
        return -3
    if D is not None and D.max() > 1.0e7 * D.min():
        return -4
    if it >= ilim:
        win = besthist[it - ilim : it + 1]
        if win.max() - win.min() < 1.0e-10:
            return -5
    if (sigma * sd > 1.0e3 * insigma).any():
        return -6
    if it > 2:
        both = np.append(arfit, besthist)
        if both.max() - both.min() < 1.0e-12:
            return -7
    if (sigma * np.append(np.abs(pc), sd.max()) < 1.0e-11 * insigma).all():
        return -8
    return None


class PenaltyState:
    """Mutable state threaded through Penalize (cmaes/_constraints.py:4-82)."""

    def __init__(self, N):
        self.weights = np.zeros(N)
        self.hist = np.ones(1)
        self.valid = False
        self.ini = True


def penalize(arx, xmean, xold, sigma, diagC, mueff, it, st, evaluate_std):
    """Box handling by penalty, cmaes/_constraints.py:4-82.  Returns
    (penalised fitness, clipped population).  Bug kept: the lower clip of the
    mean (line 53) is overwritten by line 54."""
    P, N = arx.shape
    valid = np.clip(arx, -1.0, 1.0)
    fit = evaluate_std(valid)

    q25, q75 = np.percentile(fit, [25.0, 75.0])
    delta = (q75 - q25) / N / diagC.mean() / sigma**2
    if delta == 0:
        delta = st.hist[st.hist > 0.0].min()
    elif not st.valid:
        st.hist = np.empty(0)
        st.valid = True
    if st.hist.size < 20 + (3.0 * N) / P:
        st.hist = np.append(st.hist, delta)
    else:
        st.hist = np.append(st.hist[1:], delta)

    out = (xmean < -1.0) | (xmean > 1.0)
    tx = np.where(xmean > 1.0, 1.0, xmean)
    if st.ini and out.any():
        st.weights = np.full(N, 2.0002 * np.median(st.hist))
        if st.valid and it > 2:
            st.ini = False
    if out.any():
        tx = xmean - tx
        grow = out & (np.abs(tx) > 3.0 * max(1.0, np.sqrt(N / mueff)) * sigma * np.sqrt(diagC))
        grow &= np.sign(tx) == np.sign(xmean - xold)
        st.weights = np.where(grow, st.weights * 1.2 ** min(1.0, mueff / 10.0 / N), st.weights)

    scale = np.exp(0.9 * (np.log(diagC) - np.log(diagC).mean()))
    fit = fit + ((valid - arx) ** 2) @ (st.weights / scale)
    return fit, valid


def minimize(
    fun,
    bounds,
    x0=None,
    maxiter=100,
    popsize=10,
    sigma=0.1,
    muperc=0.5,
    seed=None,
    xtol=1.0e-8,
    ftol=1.0e-8,
    constraints=None,
    return_all=False,
    verbosity=1.0,
    callback=None,
    stream=None,
    eigh=np.linalg.eigh,
    trace=None,
):
    """Driver of _cmaes.py:143-357 (works in the space standardised to [-1,1])."""
    bounds = np.asarray(bounds, dtype=np.float64)
    N, P = len(bounds), popsize
    lower, upper = bounds.T
    xm, xs = 0.5 * (upper + lower), 0.5 * (upper - lower)
    unstd = lambda x: x * xs + xm
    evaluate_std = lambda X: evaluate(fun, unstd(X))
    if constraints not in (None, "Penalize"):
        raise KeyError(constraints)
    stream = stream if stream is not None else MTStream(seed)

    xmean = stream.es_mean0(N) if x0 is None else (np.asarray(x0, dtype=np.float64) - xm) / xs
    xold = np.empty(N)
    mu, w, mueff = selection_weights(P, muperc)
    consts = strategy_constants(N, mueff)
    cc, cs, c1, cmu, damps, chind = consts

    pc, ps = np.zeros(N), np.zeros(N)
    B, D, C, invsqrtC = np.eye(N), np.ones(N), np.eye(N), np.eye(N)
    pen = PenaltyState(N)
    if return_all:
        nout = int(np.ceil(verbosity * P))
        xall = np.empty((maxiter, max(1, nout), N))
        funall = np.empty((maxiter, max(1, nout)))

    nfev = eigeneval = 0
    besthist = np.zeros(maxiter)
    ilim = int(10.0 + 30.0 * N / P)
    insigma = sigma
    it = 0
    status = None
    while status is None:
        it += 1
        Z = stream.es_z(it, P, N)
        arx = sample(xmean, sigma, B, D, Z)
        if constraints == "Penalize":
            arfit, arxvalid = penalize(arx, xmean, xold, sigma, np.diag(C), mueff, it, pen, evaluate_std)
        else:
            arxvalid = arx.copy()
            arfit = evaluate_std(arxvalid)
        nfev += P
        if return_all:
            if nout > 0:
                xall[it - 1] = unstd(arxvalid[:nout])
                funall[it - 1] = arfit[:nout]
            else:
                b = arfit.argmin()
                xall[it - 1] = unstd(arxvalid[b])
                funall[it - 1] = arfit[b]

        order = np.argsort(arfit)
        besthist[it - 1] = arfit[order[0]]
        xmean, xold, ps, pc, C, sigma, hsig = update(
            arx, order, mu, w, xmean, sigma, ps, pc, C, invsqrtC, consts, mueff, nfev, P
        )
        if nfev - eigeneval > P / (c1 + cmu) / N / 10.0:
            eigeneval = nfev
            C, B, D, invsqrtC = decompose(C, eigh)
        status = converge(it, N, maxiter, xmean, xold, besthist, arfit, order, sigma, insigma,
                          ilim, pc, xtol, ftol, np.diag(C), B, D)
        if trace is not None:
            trace.append(dict(it=it, xmean=xmean.copy(), sigma=sigma, C=C.copy(), ps=ps.copy(),
                              pc=pc.copy(), best=arfit[order[0]], hsig=bool(hsig)))
        if callback is not None:
            callback(unstd(arxvalid), dict(x=unstd(arxvalid[order[0]]), fun=arfit[order[0]], nfev=nfev, nit=it))

    xa, fa = (xall[:it], funall[:it]) if return_all else (None, None)
    return result(unstd(arxvalid[order[0]]), arfit[order[0]], status, nfev, it, xa, fa)
