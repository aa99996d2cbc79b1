"""PSO / competitive PSO front-end and generation driver (CUDA backend).

Mirrors stochopy/optimize/cpso/_cpso.py: ``minimize`` keeps the reference's
keyword signature, defaults and validation (:12-179); the generation loop
(:182-321) runs on the device through sp_pso_generation / sp_cpso_restart.
``updating`` is validated but the device path is always synchronous.
"""
import ctypes as C
import warnings

import numpy as np
import torch

from .. import _lib as L
from ._common import Engine, History, HistoryStreamer, NumpyStream, device_objective, fresh_seed, messages, validate_common, device_scope
from ._helpers import OptimizeResult, register

__all__ = ["minimize"]

_CONSTRAINTS = {None: L.CONS_NONE, "Shrink": L.CONS_SHRINK}  # cpso/_constraints.py:69-72


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
    competitivity=1.0,
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
    """Competitive PSO on the GPU; arguments as stochopy.optimize.cpso.minimize
    (_cpso.py:12-33).  ``competitivity=None`` gives plain PSO."""
    validate_common(fun, bounds, None)
    if x0 is not None:
        if np.ndim(x0) != 2 or np.shape(x0)[1] != len(bounds):
            raise ValueError()
    if popsize < 2:
        raise ValueError()
    if x0 is not None and len(x0) != popsize:
        raise ValueError()
    if not 0.0 <= inertia <= 1.0:
        raise ValueError()
    if not 0.0 <= cognitivity <= 4.0:
        raise ValueError()
    if not 0.0 <= sociability <= 4.0:
        raise ValueError()
    if competitivity is not None and not 0.0 <= competitivity <= 2.0:
        raise ValueError()
    if updating not in {"immediate", "deferred"}:
        raise ValueError()
    if updating == "immediate":
        # the reference's default is sequential by construction (_common.py:163-194: every individual sees
        # the best found so far, '<=' acceptance); on the device the population moves synchronously, as the
        # reference itself does for workers > 1 / backend="mpi" -- say so instead of switching silently
        warnings.warn("updating='immediate' runs as updating='deferred' on the CUDA backend "
                      "(synchronous generations); pass updating='deferred' to silence this", UserWarning, stacklevel=3)
    if callback is not None and not hasattr(callback, "__call__"):
        raise ValueError()
    if rng not in {"philox", "numpy"}:
        raise ValueError()
    cons = _CONSTRAINTS[constraints]  # KeyError like _cpso.py:209

    eng = Engine(dtype, device, backend)
    bounds = np.asarray(bounds, dtype=np.float64)
    N, P = len(bounds), int(popsize)
    lower, upper = bounds[:, 0].copy(), bounds[:, 1].copy()
    obj = device_objective(fun, args)
    stream = NumpyStream(seed) if rng == "numpy" else None
    gamma = competitivity
    restart = bool(gamma)  # `if gamma:` _cpso.py:215,304

    ld = eng.ld(N)
    X, V, pbest = eng.rows(P, N), eng.rows(P, N), eng.rows(P, N)
    pbestfit, pfit = eng.empty(P), eng.empty(P)
    gbest = eng.zeros(ld)
    d_lower, d_upper = eng.upload_vec(lower, ld), eng.upload_vec(upper, ld)
    ctrl, scratch = eng.new_ctrl()
    rank = eng.zeros(P, dtype=torch.int32) if restart else None
    # plain PSO: scratch that lets sp_pso_run chain its generations (no last-CTA tail per generation)
    chain_rows = None if restart else eng.zeros(int(L.load().sp_pso_chain_scalars(ld)))

    st = L.PsoState()
    st.dtype, st.objective, st.constraint = eng.sp_dt, (obj if obj is not None else L.SP_OBJ_HOST), cons
    st.P, st.N, st.maxiter, st.ld = P, N, int(maxiter), ld
    st.w, st.c1, st.c2 = float(inertia), float(cognitivity), float(sociability)
    st.xtol, st.ftol = float(xtol), float(ftol)
    st.gamma = float(gamma) if restart else -1.0
    st.chain_rows = None if chain_rows is None else chain_rows.data_ptr()
    # swarm radius threshold, _cpso.py:216
    st.delta = float(np.log(1.0 + 0.003 * P) / np.max((0.2, np.log(0.01 * maxiter)))) if restart else 0.0
    st.seed = fresh_seed(seed)
    st.X, st.V, st.pbest = X.data_ptr(), V.data_ptr(), pbest.data_ptr()
    st.pbestfit, st.pfit, st.gbest = pbestfit.data_ptr(), pfit.data_ptr(), gbest.data_ptr()
    st.lower, st.upper = d_lower.data_ptr(), d_upper.data_ptr()
    st.ctrl, st.scratch = ctrl.data_ptr(), scratch.data_ptr()

    if x0 is not None:
        eng.upload_rows(x0, out=X)
    elif stream is not None:
        jitter, perm = stream.lhs(P, N)
        d_j, d_p = eng.upload_rows(jitter), torch.from_numpy(perm).to(eng.device)
        L.call("sp_lhs_init", eng.sp_dt, X.data_ptr(), P, N, ld, st.lower, st.upper, 0, d_j.data_ptr(), d_p.data_ptr(),
               eng.stream)
    else:
        L.call("sp_lhs_init", eng.sp_dt, X.data_ptr(), P, N, ld, st.lower, st.upper, st.seed, None, None, eng.stream)
    pbest.copy_(X)  # _cpso.py:221

    eng.evaluate(fun, args, obj, X, P, N, pbestfit)  # _cpso.py:224
    pfit.copy_(pbestfit)
    L.call("sp_best_init", eng.sp_dt, X.data_ptr(), pbestfit.data_ptr(), P, N, ld, gbest.data_ptr(), ctrl.data_ptr(),
           scratch.data_ptr(), eng.stream)

    hist = History(return_all, maxiter, P, N, verbosity)
    observe = hist.enabled or callback is not None

    def snapshot(it):
        c = eng.read_ctrl(ctrl)
        if not observe:
            return c
        xbest = gbest[:N].to("cpu").numpy().astype(np.float64)
        Xh = eng.download_rows(X, P, N)
        hist.put(it, Xh, pfit.to("cpu").numpy().astype(np.float64), xbest if it == 1 else None, c.gfit)
        if callback is not None:
            res = OptimizeResult(x=xbest, fun=c.gfit, nfev=it * P, nit=it)
            hist.into(res, it)
            callback(Xh, res)
        return c

    c = snapshot(1)
    it = 1
    last = max(int(maxiter), 2)
    fast = obj is not None and stream is None and not observe
    streamer = HistoryStreamer.maybe(eng, hist, callback, P, N) if obj is not None and stream is None else None
    keep = None
    rank_ptr = rank.data_ptr() if restart else None
    eager_left, last_restart = 0, -(1 << 30)
    loop_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    loop_ev[0].record()
    while c.status == L.SP_RUNNING:
        if fast:
            n = min(64 if it < 64 else 256, last - it)
            if restart and eager_left <= 0:
                # restart out of the common path: a firing restart parks the chunk, the host resumes it.
                # Chunks are short: what is enqueued behind a parked generation still drains as no-ops.
                L.call("sp_pso_run_lazy", C.byref(st), it + 1, min(n, 32), eng.stream)
                c = eng.read_ctrl(ctrl)
                if c.status == L.SP_STATUS_RESTART_PENDING:
                    L.call("sp_cpso_restart_resume", C.byref(st), c.nit, rank_ptr, eng.stream)
                    if c.nit - last_restart < 16:  # restarts come in runs (at large popsize: every generation of the
                        eager_left = 32            # first part of a run): the gated in-chunk kernels are then cheaper
                    last_restart = c.nit
                    c.status = L.SP_RUNNING
                it = c.nit
                continue
            n = min(n, 32)
            L.call("sp_pso_run", C.byref(st), it + 1, n, rank_ptr, eng.stream)
            eager_left -= n
            c = eng.read_ctrl(ctrl)
            it = c.nit
            if c.flag > 0:  # the last generation of the window still restarted: stay with the eager sequence
                eager_left = max(eager_left, 32)
                last_restart = c.nit
            continue
        if streamer is not None:  # return_all: snapshots leave through a side stream, no per-generation sync
            for _ in range(min(64, last - it)):
                it += 1
                L.call("sp_pso_generation", C.byref(st), it, eng.stream)
                streamer.push(it, X, pfit)  # the population before the competitive restart, _cpso.py:281-307
                if restart:
                    L.call("sp_cpso_restart", C.byref(st), it, rank_ptr, eng.stream)
            c = eng.read_ctrl(ctrl)
            it = c.nit
            continue
        it += 1
        if stream is not None:
            r1, r2 = stream.pso(P, N)
            keep = (eng.upload_rows(r1), eng.upload_rows(r2))
            st.r1, st.r2 = keep[0].data_ptr(), keep[1].data_ptr()
        if obj is not None:
            L.call("sp_pso_generation", C.byref(st), it, eng.stream)
        else:
            L.call("sp_pso_propose", C.byref(st), it, eng.stream)
            eng.evaluate(fun, args, None, X, P, N, pfit)
            L.call("sp_select_sync", eng.sp_dt, it, int(maxiter), float(xtol), float(ftol), X.data_ptr(),
                   pfit.data_ptr(), pbest.data_ptr(), pbestfit.data_ptr(), P, N, ld, 1, gbest.data_ptr(),
                   ctrl.data_ptr(), scratch.data_ptr(), eng.stream)
        c = snapshot(it)
        if c.status == L.SP_RUNNING and restart:  # _cpso.py:304-307
            if stream is None:
                L.call("sp_cpso_restart", C.byref(st), it, rank_ptr, eng.stream)
            else:
                L.call("sp_cpso_restart_plan", C.byref(st), it, rank_ptr, eng.stream)
                nw = eng.read_ctrl(ctrl).flag
                if nw > 0:
                    fresh = eng.upload_rows(stream.restart(nw, N, lower, upper))
                    L.call("sp_cpso_restart_apply", C.byref(st), it, rank_ptr, fresh.data_ptr(), eng.stream)
                    eng.sync()

    it = c.nit
    loop_ev[1].record()
    loop_ev[1].synchronize()
    global LAST_LOOP  # (device time of the generation loop in ms, generations): what bench.py compares the sharded swarm with
    LAST_LOOP = (float(loop_ev[0].elapsed_time(loop_ev[1])), int(it - 1))
    if streamer is not None:
        streamer.finish(hist, it)
    res = OptimizeResult(
        x=gbest[:N].to("cpu").numpy().astype(np.float64),
        success=c.status >= 0,
        status=int(c.status),
        message=messages[int(c.status)],
        fun=float(c.gfit),
        nfev=it * P,
        nit=it,
    )
    hist.into(res, it)
    return res


LAST_LOOP = (0.0, 0)

register("cpso", minimize)
