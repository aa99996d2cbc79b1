"""VD-CMA front-end and generation driver (CUDA backend).

Mirrors stochopy/optimize/vdcma/_vdcma.py: ``minimize`` keeps the reference's
keyword signature, defaults and validation (:13-141); the generation loop
(:144-423) is device resident (sp_vd_generation): row-local sampling fused with
the objective, ranking, O(N) weighted sums over the selected rows, the scalar
step-size path from the injected pair, the natural-gradient step on (v, D) and
the termination ladder.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib as L
from ._cmaes import EsHistory, selection_weights
from ._common import HistoryStreamer, Engine, NumpyStream, device_objective, fresh_seed, messages, validate_common, device_scope
from ._helpers import OptimizeResult, register

__all__ = ["minimize"]

_CONSTRAINTS = {None: L.CONS_NONE, "Penalize": L.CONS_PENALIZE}  # reuses cmaes' Penalize, _vdcma.py:5-6


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
    _probe=None,
):
    """VD-CMA on the GPU; arguments as stochopy.optimize.vdcma.minimize (_vdcma.py:13-31)."""
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
    if rng not in {"philox", "numpy"}:
        raise ValueError()
    cons = _CONSTRAINTS[constraints]

    eng = Engine(dtype, device, backend)
    bounds = np.asarray(bounds, dtype=np.float64)
    N, P = len(bounds), int(popsize)
    lower, upper = bounds[:, 0], bounds[:, 1]
    xm, xs = 0.5 * (upper + lower), 0.5 * (upper - lower)
    unstd = lambda x: x * xs + xm
    obj = device_objective(fun, args)
    stream = NumpyStream(seed) if rng == "numpy" else None
    seed64 = fresh_seed(seed)
    ld = eng.ld(N)

    mu, w, mueff = selection_weights(P, muperc)
    if mu < 1 or P < 2:
        raise ValueError()
    # strategy parameters, _vdcma.py:192-199 (negative c1, cmu for N < 5 kept)
    cc = (4.0 + mueff / N) / (N + 4.0 + 2.0 * mueff / N)
    cf = (N - 5.0) / 6.0
    c1 = cf * 2.0 / ((N + 1.3) ** 2 + mueff)
    cmu = min(1.0 - c1, cf * 2.0 * (mueff - 2.0 + 1.0 / mueff) / ((N + 2.0) ** 2 + mueff))

    # initial mean and v, _vdcma.py:181, 208
    if x0 is not None:
        xmean = eng.upload_vec((np.asarray(x0, dtype=np.float64) - xm) / xs, ld)
    elif stream is not None:
        xmean = eng.upload_vec(stream.mean0(N), ld)
    else:
        u = eng.zeros(ld)
        L.call("sp_random_fill", eng.sp_dt, u.data_ptr(), 1, N, ld, 0, L.PURPOSE_ES_MEAN0, seed64, 0, eng.stream)
        xmean = eng.zeros(ld)
        xmean[:N] = 2.0 * u[:N] - 1.0
    if stream is not None:
        vvec = eng.upload_vec(stream.normal(N) / np.sqrt(N), ld)
    else:
        g = eng.zeros(ld)
        L.call("sp_random_fill", eng.sp_dt, g.data_ptr(), 1, N, ld, 0, L.PURPOSE_VD_V0, seed64, 1, eng.stream)
        vvec = eng.zeros(ld)
        vvec[:N] = g[:N] / np.sqrt(N)
    dvec = eng.zeros(ld)
    dvec[:N] = 1.0

    bufs = dict(
        xmean=xmean, xold=eng.zeros(ld), dx=eng.zeros(N), pc=eng.zeros(N), dvec=dvec, vvec=vvec, vn=eng.zeros(ld),
        diagC=eng.zeros(N), dy=eng.zeros(ld), ginj=eng.zeros(ld), arx=eng.rows(P, N), ary=eng.rows(P, N),
        yvn=eng.zeros(P), arfit=eng.empty(P), weights=eng.upload_vec(w), xscale=eng.upload_vec(xs, ld),
        xshift=eng.upload_vec(xm, ld), besthist=eng.zeros(max(int(maxiter), 1)),
        work=eng.zeros(int(L.load().sp_vd_work_scalars(N, P))), rank=eng.zeros(P, dtype=torch.int32),
        bnd_weights=eng.zeros(N),
    )
    hist_cap = int(20 + 3.0 * N / P) + 3
    dfithist = eng.zeros(hist_cap)
    dfithist[0] = 1.0
    host = L.EsCtrl()
    host.base.status = L.SP_RUNNING
    host.sigma = host.sigma_gen = float(sigma)
    host.iniphase, host.hist_len = 1, 1
    ctrl = eng.new_struct(host)

    st = L.VdState()
    st.dtype, st.objective, st.constraint, st.N = eng.sp_dt, (obj if obj is not None else L.SP_OBJ_HOST), cons, N
    st.P, st.ld, st.mu, st.maxiter = P, ld, mu, int(maxiter)
    st.ilim, st.hist_cap = int(10 + 30 * N / P), hist_cap
    st.cc, st.c1, st.cmu, st.mueff, st.wsum = cc, c1, cmu, mueff, float(w.sum())
    st.xtol, st.ftol, st.insigma, st.seed = float(xtol), float(ftol), float(sigma), seed64
    for k, t in bufs.items():
        setattr(st, k, t.data_ptr())
    st.dfithist, st.ctrl = dfithist.data_ptr(), ctrl.data_ptr()
    st.host_z = int(stream is not None)

    hist = EsHistory(return_all, maxiter, P, N, verbosity)
    observe = hist.enabled or callback is not None
    penal = cons == L.CONS_PENALIZE
    # nobody reads the population: the sampling kernel keeps only y and the fitness (the result row is
    # rebuilt below as xold + sigma_gen * y)
    lean = obj is not None and stream is None and not observe and _probe is None and not penal
    st.lean = int(lean)
    L.call("sp_vd_refresh", C.byref(st), eng.stream)
    arx, arfit = bufs["arx"], bufs["arfit"]
    valid_rows = (lambda r: np.clip(r, -1.0, 1.0)) if penal else (lambda r: r)

    fast = obj is not None and stream is None and not observe and _probe is None
    streamer = HistoryStreamer.maybe(eng, hist, callback, P, N) if obj is not None and stream is None and _probe is None else None
    it = 0
    last = max(int(maxiter), 1)
    c = eng.read_ctrl(ctrl, L.EsCtrl)
    while c.base.status == L.SP_RUNNING:
        if fast:
            n = min(16 if it < 16 else 64, last - it)
            L.call("sp_vd_run", C.byref(st), it + 1, n, eng.stream)
            c = eng.read_ctrl(ctrl, L.EsCtrl)
            it = c.base.nit
            continue
        if streamer is not None:  # return_all: snapshots leave through a side stream, no per-generation sync
            for _ in range(min(16, last - it)):
                it += 1
                L.call("sp_vd_generation", C.byref(st), it, eng.stream)
                streamer.push(it, arx, arfit)
            c = eng.read_ctrl(ctrl, L.EsCtrl)
            it = c.base.nit
            continue
        it += 1
        if stream is not None:  # randn(P, N), then from generation 2 on randn(N) (_vdcma.py:239, 246)
            eng.upload_rows(stream.normal(P, N), out=bufs["ary"])
            if it > 1:
                bufs["ginj"][:N].copy_(torch.from_numpy(stream.normal(N).astype(eng.np_dt)))
        if obj is not None:
            L.call("sp_vd_generation", C.byref(st), it, eng.stream)
        else:
            L.call("sp_vd_sample", C.byref(st), it, 0, eng.stream)
            eng.evaluate(fun, args, None, arx, P, N, arfit, bufs["xscale"], bufs["xshift"],
                         to_user=lambda X: unstd(valid_rows(X)), clip=penal)
            L.call("sp_vd_update", C.byref(st), it, eng.stream)
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
    if lean:  # x = xmean_old + sigma * y in the working precision, as the exact kernel stores it (_vdcma.py:249)
        yb = bufs["ary"][c.base.gbest_row, :N].to("cpu").numpy()
        xo = bufs["xold"][:N].to("cpu").numpy()
        best = (xo + eng.np_dt.type(c.sigma_gen) * yb).astype(np.float64)
    else:
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


register("vdcma", minimize)
