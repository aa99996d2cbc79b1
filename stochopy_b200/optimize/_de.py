"""Differential evolution front-end and generation driver (CUDA backend).

Mirrors stochopy/optimize/de/_de.py: ``minimize`` keeps the reference's keyword
signature, defaults and validation (:13-173); the generation loop (:176-301)
runs on the device through sp_de_generation / sp_de_run.  ``updating`` is
accepted and validated but the device path is always synchronous ("deferred"),
exactly as the reference forces for workers > 1 or backend == "mpi" (:142-145).
"""
import ctypes as C
import warnings

import numpy as np
import torch

from .. import _lib as L
from ._common import Engine, History, HistoryStreamer, NumpyStream, device_objective, fresh_seed, messages, validate_common, device_scope
from ._helpers import OptimizeResult, register

__all__ = ["minimize"]

_CONSTRAINTS = {None: L.CONS_NONE, "Random": L.CONS_RANDOM}  # de/_constraints.py:31-34


@device_scope
def minimize(
    fun,
    bounds,
    x0=None,
    args=(),
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
    workers=1,
    backend=None,
    return_all=False,
    verbosity=1.0,
    callback=None,
    dtype="float64",
    device=None,
    rng="philox",
):
    """Differential Evolution on the GPU; arguments as stochopy.optimize.de.minimize
    (_de.py:13-33).  ``workers``/``backend`` are accepted for compatibility; the
    population is always evaluated on the device."""
    validate_common(fun, bounds, None)
    if x0 is not None:
        if np.ndim(x0) != 2 or np.shape(x0)[1] != len(bounds):
            raise ValueError()
    if popsize < 2:
        raise ValueError()
    if x0 is not None and len(x0) != popsize:
        raise ValueError()
    if not 0.0 <= mutation <= 2.0:
        raise ValueError()
    if not 0.0 <= recombination <= 1.0:
        raise ValueError()
    if updating not in {"immediate", "deferred"}:
        raise ValueError()
    if updating == "immediate":
        # the reference's default is sequential by construction (_common.py:163-194: every individual sees
        # the best found so far, '<=' acceptance); on the device the population moves synchronously, as the
        # reference itself does for workers > 1 / backend="mpi" -- say so instead of switching silently
        warnings.warn("updating='immediate' runs as updating='deferred' on the CUDA backend "
                      "(synchronous generations); pass updating='deferred' to silence this", UserWarning, stacklevel=3)
    strat = L.DE_STRATEGIES[strategy]  # KeyError like _strategy_map[strategy], _de.py:140
    if callback is not None and not hasattr(callback, "__call__"):
        raise ValueError()
    if rng not in {"philox", "numpy"}:
        raise ValueError()
    cons = _CONSTRAINTS[constraints]  # KeyError like _constraints_map[constraints], _de.py:202
    if popsize <= L.DE_DONORS[strategy]:
        raise ValueError()

    eng = Engine(dtype, device, backend)
    bounds = np.asarray(bounds, dtype=np.float64)
    N, P = len(bounds), int(popsize)
    lower, upper = bounds[:, 0].copy(), bounds[:, 1].copy()
    obj = device_objective(fun, args)
    stream = NumpyStream(seed) if rng == "numpy" else None

    ld = eng.ld(N)
    X = [eng.rows_scratch(P, N), eng.rows_scratch(P, N)]
    pbestfit, pfit = eng.empty(P), eng.empty(P)
    gbest = eng.zeros(ld)
    d_lower, d_upper = eng.upload_vec(lower, ld), eng.upload_vec(upper, ld)
    ctrl, scratch = eng.new_ctrl()

    st = L.DeState()
    st.dtype, st.objective, st.strategy, st.constraint = eng.sp_dt, (obj if obj is not None else L.SP_OBJ_HOST), strat, cons
    st.P, st.N, st.maxiter, st.ld = P, N, int(maxiter), ld
    st.F, st.CR, st.xtol, st.ftol = float(mutation), float(recombination), float(xtol), float(ftol)
    st.seed = fresh_seed(seed)
    st.X[0], st.X[1] = X[0].data_ptr(), X[1].data_ptr()
    st.pbestfit, st.pfit, st.gbest = pbestfit.data_ptr(), pfit.data_ptr(), gbest.data_ptr()
    st.lower, st.upper = d_lower.data_ptr(), d_upper.data_ptr()
    st.ctrl, st.scratch = ctrl.data_ptr(), scratch.data_ptr()

    # initial population: x0 (copied, the caller's array is not aliased) or LHS, _de.py:208
    if x0 is not None:
        eng.upload_rows(x0, out=X[0])
    elif stream is not None:
        jitter, perm = stream.lhs(P, N)
        d_j, d_p = eng.upload_rows(jitter), torch.from_numpy(perm).to(eng.device)
        L.call("sp_lhs_init", eng.sp_dt, X[0].data_ptr(), P, N, ld, st.lower, st.upper, 0, d_j.data_ptr(),
               d_p.data_ptr(), eng.stream)
    else:
        L.call("sp_lhs_init", eng.sp_dt, X[0].data_ptr(), P, N, ld, st.lower, st.upper, st.seed, None, None, eng.stream)

    eng.evaluate(fun, args, obj, X[0], P, N, pbestfit)  # _de.py:212
    pfit.copy_(pbestfit)
    L.call("sp_best_init", eng.sp_dt, X[0].data_ptr(), pbestfit.data_ptr(), P, N, ld, gbest.data_ptr(),
           ctrl.data_ptr(), scratch.data_ptr(), eng.stream)

    hist = History(return_all, maxiter, P, N, verbosity)
    observe = hist.enabled or callback is not None

    def snapshot(it, cur):
        c = eng.read_ctrl(ctrl)
        xbest = gbest[:N].to("cpu").numpy().astype(np.float64)
        if not observe:
            return c, xbest
        Xh = eng.download_rows(X[cur], P, N)
        hist.put(it, Xh, pfit.to("cpu").numpy().astype(np.float64), xbest if it == 1 else None, c.gfit)
        if callback is not None:
            res = OptimizeResult(x=xbest, fun=c.gfit, nfev=it * P, nit=it)
            hist.into(res, it)
            callback(Xh, res)
        return c, xbest

    it = 1
    last = max(int(maxiter), 2)
    fast = obj is not None and stream is None and not observe
    if fast:  # nobody looks at generation 1: no host round trip before the first chunk (status cannot be set yet)
        c = L.Ctrl()
        c.status, c.nit = L.SP_RUNNING, 1
    else:
        c, xbest = snapshot(1, 0)
    streamer = HistoryStreamer.maybe(eng, hist, callback, P, N) if obj is not None and stream is None else None
    keep = None
    while c.status == L.SP_RUNNING:
        if fast:  # enqueue a chunk of generations; kernels no-op once ctrl.status is set
            n = min(64 if it < 64 else 256, last - it)
            L.call("sp_de_run", C.byref(st), it + 1, n, eng.stream)
            c = eng.read_ctrl(ctrl)
            it = c.nit
            continue
        if streamer is not None:  # return_all: snapshots leave through a side stream, no per-generation sync
            for _ in range(min(64, last - it)):
                it += 1
                L.call("sp_de_generation", C.byref(st), it, eng.stream)
                streamer.push(it, X[(it & 1) ^ 1], pfit)
            c = eng.read_ctrl(ctrl)
            it = c.nit
            continue
        it += 1
        if stream is not None:
            r1, donors, irand, rep = stream.de(P, N, L.DE_DONORS[strategy], lower, upper, cons == L.CONS_RANDOM)
            keep = (eng.upload_rows(r1), torch.from_numpy(donors).to(eng.device),
                    torch.from_numpy(irand).to(eng.device), None if rep is None else eng.upload_rows(rep))
            st.r1, st.donors, st.irand = keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr()
            st.repair = None if rep is None else keep[3].data_ptr()
        new = (it & 1) ^ 1
        if obj is not None:
            L.call("sp_de_generation", C.byref(st), it, eng.stream)
        else:  # user objective: propose on device, evaluate through fun(x), select on device
            L.call("sp_de_propose", C.byref(st), it, eng.stream)
            eng.evaluate(fun, args, None, X[new], P, N, pfit)
            L.call("sp_select_sync", eng.sp_dt, it, int(maxiter), float(xtol), float(ftol), X[it & 1].data_ptr(),
                   pfit.data_ptr(), X[new].data_ptr(), pbestfit.data_ptr(), P, N, ld, 0, gbest.data_ptr(),
                   ctrl.data_ptr(), scratch.data_ptr(), eng.stream)
        c, xbest = snapshot(it, new)

    it = c.nit
    if streamer is not None:
        streamer.finish(hist, it)
    xbest = gbest[:N].to("cpu").numpy().astype(np.float64)
    res = OptimizeResult(
        x=xbest,
        success=c.status >= 0,
        status=int(c.status),
        message=messages[int(c.status)],
        fun=float(c.gfit),
        nfev=it * P,
        nit=it,
    )
    hist.into(res, it)
    return res


register("de", minimize)
