"""Oracle: VD-CMA (reference: stochopy/optimize/vdcma/_vdcma.py).

Covariance model C = D (I + v v^T) D with D = diag(dvec); every update is O(N)
per individual (Akimoto, Auger & Hansen, GECCO 2014)."""
import numpy as np

from .cmaes import PenaltyState, converge, penalize, selection_weights
from .common import result
from .objectives import evaluate
from .streams import MTStream


def strategy_constants(N, mueff):
    """_vdcma.py:192-199 (c1, cmu negative for N < 5: cfactor quirk kept)."""
    cc = (4.0 + mueff / N) / (N + 4.0 + 2.0 * mueff / N)
    cf = (N - 5.0) / 6.0
    c1 = cf * 2.0 / ((N + 1.3) ** 2 + mueff)
    cmu = min(1.0 - c1, cf * 2.0 * (mueff - 2.0 + 1.0 / mueff) / ((N + 2.0) ** 2 + mueff))
    return cc, c1, cmu


def sample(Z, dvec, vvec):
    """_vdcma.py:240-242: y = D (z + (sqrt(1+|v|^2)-1) (z.vn) vn)."""
    nv2 = vvec @ vvec
    vn = vvec / np.sqrt(nv2)
    return dvec * (Z + (np.sqrt(1.0 + nv2) - 1.0) * np.outer(Z @ vn, vn))


def injection(dx, dvec, vvec, g):
    """_vdcma.py:244-246: rescale the last mean shift to the Mahalanobis length
    of a fresh N(0,I) vector g."""
    nv2 = vvec @ vvec
    ddx = dx / dvec
    mnorm = (ddx**2).sum() - (ddx @ vvec) ** 2 / (1.0 + nv2)
    return np.linalg.norm(g) / np.sqrt(mnorm) * dx


def diag_cov(dvec, vvec):
    """_vdcma.py:251-256 without the two dense N x N products (same values:
    the products only add exact zeros)."""
    return (dvec * (1.0 + vvec * vvec)) * dvec


def pvec_qvec(vn, nv2, y, w=None):
    """_vdcma.py:426-441."""
    yv = y @ vn
    if w is None:
        p = y**2 - nv2 / (1.0 + nv2) * (yv * y * vn) - 1.0
        q = yv * y - (0.5 * (yv**2 + 1.0 + nv2)) * vn
        return p, q
    p = w @ (y**2 - nv2 / (1.0 + nv2) * (yv[:, None] * (y * vn)) - 1.0)
    q = w @ (yv[:, None] * y - np.outer(0.5 * (yv**2 + 1.0 + nv2), vn))
    return p, q


def alpha_terms(nv2, vnn):
    """_vdcma.py:318-329."""
    gamma = 1.0 / np.sqrt(1.0 + nv2)
    alpha = np.sqrt(nv2**2 + (1.0 + nv2) / vnn.max() * (2.0 - gamma)) / (2.0 + nv2)
    if alpha < 1.0:
        beta = (4.0 - (2.0 - gamma) / vnn.max()) / (1.0 + 2.0 / nv2) ** 2
    else:
        alpha, beta = 1.0, 0.0
    bsca = 2.0 * alpha**2 - beta
    avec = 2.0 - (bsca + 2.0 * alpha**2) * vnn
    return alpha, bsca, avec, vnn / avec


def natural_gradient(dvec, vn, vnn, nv, nv2, alpha, avec, bsca, invavnn, p, q):
    """_vdcma.py:444-458."""
    r = p - alpha / (1.0 + nv2) * ((2.0 + nv2) * q * vn - nv2 * (vn @ q) * vnn)
    s = r / avec - bsca * (r @ invavnn) / (1.0 + bsca * (vnn @ invavnn)) * invavnn
    ngv = q / nv - alpha / nv * ((2.0 + nv2) * (vn * s) - (s @ vnn) * vn)
    return ngv, dvec * s


def adapt(ary_elite, w, pc, dvec, vvec, c1, cmu, hsig):
    """Restricted covariance update, _vdcma.py:318-372.  Returns new (dvec, vvec)."""
    N = dvec.size
    nv2 = vvec @ vvec
    nv = np.sqrt(nv2)
    vn = vvec / nv
    vnn = vn**2
    alpha, bsca, avec, invavnn = alpha_terms(nv2, vnn)
    if cmu == 0.0:
        p_mu = q_mu = np.zeros(N)
    else:
        p_mu, q_mu = pvec_qvec(vn, nv2, ary_elite / dvec, w)
    if c1 == 0.0:
        p_1 = q_1 = np.zeros(N)
    else:
        p_1, q_1 = pvec_qvec(vn, nv2, pc / dvec)
    p = cmu * p_mu
    q = cmu * q_mu
    if hsig:
        p = p + c1 * p_1
        q = q + c1 * q_1
    if cmu + c1 > 0.0:
        ngv, ngd = natural_gradient(dvec, vn, vnn, nv, nv2, alpha, avec, bsca, invavnn, p, q)
        up = min(1.0, 0.7 * nv / np.sqrt(ngv @ ngv))
        up = min(up, 0.7 * (dvec / np.abs(ngd)).min())
    else:
        ngv = ngd = np.zeros(N)
        up = 1.0
    return dvec + up * ngd, vvec + up * ngv


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
    trace=None,
):
    """Driver of _vdcma.py:144-423."""
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
    cc, c1, cmu = strategy_constants(N, mueff)

    inject = False
    cs, ds = 0.3, np.sqrt(N)
    dx = np.zeros(N)
    ps = 0.0
    dvec = np.ones(N)
    vvec = stream.vd_v0(N) / np.sqrt(N)
    pc = np.zeros(N)
    pen = PenaltyState(N)
    if return_all:
        nout = int(np.ceil(verbosity * P))
        xall = np.empty((maxiter, max(1, nout), N))
        funall = np.empty((maxiter, max(1, nout)))

    nfev = 0
    besthist = np.zeros(maxiter)
    ilim = int(10 + 30 * N / P)
    insigma = sigma
    it = 0
    status = None
    while status is None:
        it += 1
        Z = stream.es_z(it, P, N)
        ary = sample(Z, dvec, vvec)
        if inject:
            dy = injection(dx, dvec, vvec, stream.vd_inject(it, N))
            ary[0] = dy
            ary[1] = -dy
        arx = xmean + sigma * ary
        diagC = diag_cov(dvec, vvec)
        if constraints == "Penalize":
            arfit, arxvalid = penalize(arx, xmean, xold, sigma, diagC, mueff, it, pen, evaluate_std)
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
        dx = w @ arx[order[:mu]] - w.sum() * xmean
        xold = xmean.copy()
        xmean = xmean + dx
        besthist[it - 1] = arfit[order[0]]

        # step size from the rank gap of the injected pair, _vdcma.py:299-307
        if inject:
            gap = int(np.where(order == 1)[0][0]) - int(np.where(order == 0)[0][0])
            ps += cs * (gap / (P - 1.0) - ps)
            sigma *= np.exp(ps / ds)
            hsig = ps < 0.5
        else:
            inject = True
            hsig = True

        pc = (1.0 - cc) * pc
        if hsig:
            pc = pc + np.sqrt(cc * (2.0 - cc) * mueff) * (w @ ary[order[:mu]])

        dvec, vvec = adapt(ary[order[:mu]], w, pc, dvec, vvec, c1, cmu, hsig)

        status = converge(it, N, maxiter, xmean, xold, besthist, arfit, order, sigma, insigma,
                          ilim, pc, xtol, ftol, diagC)
        if trace is not None:
            trace.append(dict(it=it, xmean=xmean.copy(), sigma=sigma, dvec=dvec.copy(),
                              vvec=vvec.copy(), pc=pc.copy(), best=arfit[order[0]]))
        if callback is not None:
            callback(unstd(arxvalid), dict(x=unstd(arxvalid[order[0]]), fun=arfit[order[0]], nfev=nfev, nit=it))

    xa, fa = (xall[:it], funall[:it]) if return_all else (None, None)
    return result(unstd(arxvalid[order[0]]), arfit[order[0]], status, nfev, it, xa, fa)
