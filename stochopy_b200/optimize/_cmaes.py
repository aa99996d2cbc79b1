"""CMA-ES front-end and generation driver (CUDA backend).

Mirrors stochopy/optimize/cmaes/_cmaes.py: ``minimize`` keeps the reference's
keyword signature, defaults and validation (:12-140); the generation loop
(:143-357) is device resident (sp_cma_generation): sampling GEMM, objective,
ranking, paths, rank-mu covariance update, Jacobi eigendecomposition and the
termination ladder never leave the GPU.

Extra options: ``eigh`` = 'device' (default: Jacobi solver on the GPU, eigenvector
sign normalised) or 'host' (numpy's LAPACK ``eigh`` on the N x N matrix -- only
useful with ``rng='numpy'`` to retrace the reference's fixed-seed trajectory, whose
samples depend on the sign LAPACK happens to give each eigenvector).
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib as L
from ._common import HistoryStreamer, Engine, NumpyStream, device_objective, fresh_seed, messages, validate_common, device_scope
from ._helpers import OptimizeResult, register

__all__ = ["minimize"]

_CONSTRAINTS = {None: L.CONS_NONE, "Penalize": L.CONS_PENALIZE}  # cmaes/_constraints.py:85-87


def selection_weights(popsize, muperc):
    """mu, log-rank weights, mueff (_cmaes.py:184-189)."""
    mu = int(muperc * popsize)
    w = np.log(mu + 0.5) - np.log(np.arange(1, mu + 1))
    w /= w.sum()
    return mu, w, w.sum() ** 2 / np.square(w).sum()


class EsHistory:
    """xall / funall of the ES methods (_cmaes.py:212-215, 261-269)."""

    def __init__(self, enabled, maxiter, P, N, verbosity):
        self.enabled = bool(enabled)
        if self.enabled:
            self.nout = int(np.ceil(verbosity * P))
            self.xall = np.empty((maxiter, max(1, self.nout), N))
            self.funall = np.empty((maxiter, max(1, self.nout)))

    def put(self, it, X, fit):
        if not self.enabled:
            return
        if self.nout > 0:
            self.xall[it - 1] = X[: self.nout]
            self.funall[it - 1] = fit[: self.nout]
        else:
            b = int(np.argmin(fit))
            self.xall[it - 1] = X[b]
            self.funall[it - 1] = fit[b]

    def into(self, res, it):
        if self.enabled:
            res.update({"xall": self.xall[:it], "funall": self.funall[:it]})


@device_scope
def minimize(
    fun,
    bounds,
    x0=None,
    args=(),
    maxiter=100,
    popsize=10,
    sigma=0.1,
    muperc=0.5,
    seed=None,
    xtol=1.0e-8,
    ftol=1.0e-8,
    constraints=None,
    workers=1,
    backend=None,
    return_all=False,
    verbosity=1.0,
    callback=None,
    dtype="float64",
    device=None,
    rng="philox",
    eigh="device",
    _probe=None,
):
    """CMA-ES on the GPU; arguments as stochopy.optimize.cmaes.minimize (_cmaes.py:12-30)."""
    validate_common(fun, bounds, None)
    if x0 is not None:
        if np.ndim(x0) != 1 or len(x0) != len(bounds):
            raise ValueError()
    if sigma <= 0.0:
        raise ValueError()
    if not 0.0 < muperc <= 1.0:
        raise ValueError()
    if callback is not None and not hasattr(callback, "__call__"):
        raise ValueError()
    if rng not in {"philox", "numpy"} or eigh not in {"device", "host"}:
        raise ValueError()
    cons = _CONSTRAINTS[constraints]  # KeyError like _cmaes.py:177

    eng = Engine(dtype, device, backend)
    bounds = np.asarray(bounds, dtype=np.float64)
    N, P = len(bounds), int(popsize)
    lower, upper = bounds[:, 0], bounds[:, 1]
    xm, xs = 0.5 * (upper + lower), 0.5 * (upper - lower)  # _cmaes.py:166-171
    unstd = lambda x: x * xs + xm
    obj = device_objective(fun, args)
    stream = NumpyStream(seed) if rng == "numpy" else None
    seed64 = fresh_seed(seed)
    ld = eng.ld(N)

    mu, w, mueff = selection_weights(P, muperc)
    if mu < 1:
        raise ValueError()
    # strategy parameters, _cmaes.py:192-205
    cc = (4.0 + mueff / N) / (N + 4.0 + 2.0 * mueff / N)
    cs = (mueff + 2.0) / (N + mueff + 5.0)
    c1 = 2.0 / ((N + 1.3) ** 2 + mueff)
    cmu = min(1.0 - c1, 2.0 * (mueff - 2.0 + 1.0 / mueff) / ((N + 2.0) ** 2 + mueff))
    damps = 1.0 + 2.0 * max(0.0, np.sqrt((mueff - 1.0) / (N + 1.0)) - 1.0) + cs
    chind = np.sqrt(N) * (1.0 - 1.0 / (4.0 * N) + 1.0 / (21.0 * N**2))

    # initial mean, _cmaes.py:180
    if x0 is not None:
        xmean = eng.upload_vec((np.asarray(x0, dtype=np.float64) - xm) / xs, ld)
    elif stream is not None:
        xmean = eng.upload_vec(stream.mean0(N), ld)
    else:
        u = eng.zeros(ld)
        L.call("sp_random_fill", eng.sp_dt, u.data_ptr(), 1, N, ld, 0, L.PURPOSE_ES_MEAN0, seed64, 0, eng.stream)
        xmean = eng.zeros(ld)
        xmean[:N] = 2.0 * u[:N] - 1.0

    eye = torch.eye(N, dtype=eng.t_dt, device=eng.device)
    bufs = dict(
        xmean=xmean, xold=eng.zeros(ld), pc=eng.zeros(N), ps=eng.zeros(N), C=eye.clone(), B=eye.clone(),
        D=torch.ones(N, dtype=eng.t_dt, device=eng.device), invsqrtC=eye.clone(), arx=eng.rows(P, N), arfit=eng.empty(P),
        Z=eng.rows(P, N), weights=eng.upload_vec(w), xscale=eng.upload_vec(xs, ld), xshift=eng.upload_vec(xm, ld),
        besthist=eng.zeros(max(int(maxiter), 1)), work=eng.zeros(int(L.load().sp_cma_work_scalars(N, P))),
        rank=eng.zeros(P, dtype=torch.int32), bnd_weights=eng.zeros(N),
    )
    hist_cap = int(20 + 3.0 * N / P) + 3
    dfithist = eng.zeros(hist_cap)
    dfithist[0] = 1.0  # dfithist = np.ones(1), _cmaes.py:209
    host = L.EsCtrl()
    host.base.status = L.SP_RUNNING
    host.sigma = host.sigma_gen = float(sigma)
    host.iniphase, host.hist_len = 1, 1
    ctrl = eng.new_struct(host)

    st = L.CmaState()
    st.dtype, st.objective, st.constraint, st.N = eng.sp_dt, (obj if obj is not None else L.SP_OBJ_HOST), cons, N
    st.P, st.ld, st.mu, st.maxiter = P, ld, mu, int(maxiter)
    st.ilim, st.hist_cap = int(10.0 + 30.0 * N / P), hist_cap
    st.cc, st.cs, st.c1, st.cmu, st.damps, st.chind, st.mueff = cc, cs, c1, cmu, damps, chind, mueff
    st.xtol, st.ftol, st.insigma, st.seed = float(xtol), float(ftol), float(sigma), seed64
    for k, t in bufs.items():
        setattr(st, k, t.data_ptr())
    st.dfithist, st.ctrl = dfithist.data_ptr(), ctrl.data_ptr()
    st.host_z, st.host_eigh = int(stream is not None), int(eigh == "host")

    hist = EsHistory(return_all, maxiter, P, N, verbosity)
    observe = hist.enabled or callback is not None
    penal = cons == L.CONS_PENALIZE
    arx, arfit, Cm = bufs["arx"], bufs["arfit"], bufs["C"]

    def valid_rows(rows):  # arxvalid: the clipped population under Penalize (cmaes/_constraints.py:30-31)
        return np.clip(rows, -1.0, 1.0) if penal else rows

    fast = obj is not None and stream is None and eigh == "device" and not observe and _probe is None
    streamer = (HistoryStreamer.maybe(eng, hist, callback, P, N)
                if obj is not None and stream is None and eigh == "device" and _probe is None else None)
    it = 0
    last = max(int(maxiter), 1)
    c = eng.read_ctrl(ctrl, L.EsCtrl)
    while c.base.status == L.SP_RUNNING:
        if fast:
            n = min(16 if it < 16 else 64, last - it)
            L.call("sp_cma_run", C.byref(st), it + 1, n, eng.stream)
            c = eng.read_ctrl(ctrl, L.EsCtrl)
            it = c.base.nit
            continue
        if streamer is not None:  # return_all: snapshots leave through a side stream, no per-generation sync
            for _ in range(min(16, last - it)):
                it += 1
                L.call("sp_cma_generation", C.byref(st), it, eng.stream)
                streamer.push(it, arx, arfit)
            c = eng.read_ctrl(ctrl, L.EsCtrl)
            it = c.base.nit
            continue
        it += 1
        if stream is not None:  # P draws of randn(N) == one randn(P, N), _cmaes.py:234
            eng.upload_rows(stream.normal(P, N), out=bufs["Z"])
        if obj is not None and eigh == "device":
            L.call("sp_cma_generation", C.byref(st), it, eng.stream)
        else:
            L.call("sp_cma_sample", C.byref(st), it, eng.stream)
            if obj is not None:  # the fused path clips inside the kernel; here evaluate a clipped copy
                src = arx.clamp(-1.0, 1.0) if penal else arx
                eng.evaluate(fun, args, obj, src, P, N, arfit, bufs["xscale"], bufs["xshift"])
            else:
                eng.evaluate(fun, args, None, arx, P, N, arfit, bufs["xscale"], bufs["xshift"],
                             to_user=lambda X: unstd(valid_rows(X)), clip=penal)
            L.call("sp_cma_update", C.byref(st), it, eng.stream)
            if eigh == "host":
                if eng.read_ctrl(ctrl, L.EsCtrl).do_eig:  # _cmaes.py:301-309 with numpy's LAPACK
                    Ch = Cm.to("cpu").numpy().astype(np.float64)
                    Ch = np.triu(Ch) + np.triu(Ch, 1).T
                    vals, vecs = np.linalg.eigh(Ch)
                    order = np.argsort(vals)
                    Cm.copy_(torch.from_numpy(Ch.astype(eng.np_dt)))
                    bufs["B"].copy_(torch.from_numpy(np.ascontiguousarray(vecs[:, order]).astype(eng.np_dt)))
                    bufs["D"].copy_(torch.from_numpy(vals[order].astype(eng.np_dt)))
                L.call("sp_cma_finish_generation", C.byref(st), it, eng.stream)
        c = eng.read_ctrl(ctrl, L.EsCtrl)
        if _probe is not None:  # test hook: device state after every generation (tests/test_gpu_sizes.py)
            _probe(it, bufs, c)
        if observe:
            Xh = unstd(valid_rows(eng.download_rows(arx, P, N)))
            fh = arfit.to("cpu").numpy().astype(np.float64)
            hist.put(it, Xh, fh)
            if callback is not None:
                res = OptimizeResult(x=Xh[c.base.gbest_row], fun=c.base.gfit, nfev=int(c.nfev), nit=it)
                hist.into(res, it)
                callback(Xh, res)

    it = c.base.nit
    if c.base.status == L.SP_STATUS_INTERNAL:
        raise L.EngineError("inconsistent device state (ranking without a rank 0)")
    if streamer is not None:
        streamer.finish(hist, it, transform=lambda X: unstd(valid_rows(X)))
    best = arx[c.base.gbest_row, :N].to("cpu").numpy().astype(np.float64)
    res = OptimizeResult(
        x=unstd(valid_rows(best)),
        success=c.base.status >= 0,
        status=int(c.base.status),
        message=messages[int(c.base.status)],
        fun=float(c.base.gfit),
        nfev=int(c.nfev),
        nit=it,
    )
    hist.into(res, it)
    return res


register("cmaes", minimize)
