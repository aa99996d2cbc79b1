"""Neighbourhood Algorithm front-end and generation driver (CUDA backend).

Mirrors stochopy/optimize/na/_na.py: ``minimize`` keeps the reference's keyword
signature, defaults and validation (:11-131, including the ``callback=True``
default that makes a direct call raise ValueError); the generation loop (:134-262)
keeps the growing archive on the device (transposed) and resamples inside the
Voronoi cells of the ``nr`` best models with sp_na_resample.
"""
import numpy as np
import torch

from .. import _lib as L
from ._common import Engine, History, NumpyStream, device_objective, fresh_seed, messages, validate_common, device_scope
from ._helpers import OptimizeResult, register

__all__ = ["minimize"]


@device_scope
def minimize(
    fun,
    bounds,
    x0=None,
    args=(),
    maxiter=100,
    popsize=10,
    nrperc=0.5,
    seed=None,
    xtol=1.0e-8,
    ftol=1.0e-8,
    workers=1,
    backend=None,
    return_all=False,
    verbosity=1.0,
    callback=True,
    dtype="float64",
    device=None,
    rng="philox",
):
    """Neighborhood Algorithm on the GPU; arguments as stochopy.optimize.na.minimize (_na.py:11-27)."""
    validate_common(fun, bounds, None)
    if x0 is not None:
        if np.ndim(x0) != 2 or np.shape(x0)[1] != len(bounds):
            raise ValueError()
    if popsize < 2:
        raise ValueError()
    if x0 is not None and len(x0) != popsize:
        raise ValueError()
    if not 0.0 < nrperc <= 1.0:
        raise ValueError()
    if callback is not None and not hasattr(callback, "__call__"):
        raise ValueError()
    if rng not in {"philox", "numpy"}:
        raise ValueError()

    eng = Engine(dtype, device, backend)
    bounds = np.asarray(bounds, dtype=np.float64)
    N, P = len(bounds), int(popsize)
    lower, upper = bounds[:, 0].copy(), bounds[:, 1].copy()
    span = upper - lower  # _na.py:156-161
    span_mask = span > 0.0
    span = np.where(span_mask, span, 1.0)
    unnorm = lambda x: np.where(span_mask, x * span + lower, upper)
    obj = device_objective(fun, args)
    stream = NumpyStream(seed) if rng == "numpy" else None
    seed64 = fresh_seed(seed)
    nr = max(1, int(nrperc * P))
    ld = eng.ld(N)
    cap = P * max(int(maxiter), 2)

    X, pbest = eng.rows(P, N), eng.rows(P, N)
    pbestfit, pfit = eng.empty(P), eng.empty(P)
    gbest = eng.zeros(ld)
    d_lower, d_upper = eng.upload_vec(lower, ld), eng.upload_vec(upper, ld)
    # unnormalize as an affine map for the objective kernel: x * scale + shift (zero-span axes -> upper)
    d_scale = eng.upload_vec(np.where(span_mask, span, 0.0), ld)
    d_shift = eng.upload_vec(np.where(span_mask, lower, upper), ld)
    d_mask = torch.from_numpy(span_mask.astype(np.int32)).to(eng.device)
    ctrl, scratch = eng.new_ctrl()
    # the per-walker distance scratch is P x cap scalars and the transposed archive N x cap (cap = popsize * maxiter):
    # quadratic in popsize -- fail early and clearly instead of running the device out of memory
    need = (P * cap + (N + 2) * cap) * eng.np_dt.itemsize
    free, _total = torch.cuda.mem_get_info(eng.device)
    if need > 0.9 * free:
        raise MemoryError(f"na: popsize={P}, maxiter={maxiter} needs {need / 2**30:.1f} GiB of device scratch "
                          f"(popsize^2 * maxiter scalars); {free / 2**30:.1f} GiB are free -- use a smaller popsize or maxiter")
    archT = eng.zeros(N, cap)
    archfit = eng.zeros(cap)
    rank = eng.zeros(cap, dtype=torch.int32)
    cells = eng.zeros(nr, dtype=torch.int32)
    work = eng.zeros(P * cap)

    if x0 is not None:
        eng.upload_rows(x0, out=X)
    elif stream is not None:
        jitter, perm = stream.lhs(P, N)
        d_j, d_p = eng.upload_rows(jitter), torch.from_numpy(perm).to(eng.device)
        L.call("sp_lhs_init", eng.sp_dt, X.data_ptr(), P, N, ld, d_lower.data_ptr(), d_upper.data_ptr(), 0,
               d_j.data_ptr(), d_p.data_ptr(), eng.stream)
    else:
        L.call("sp_lhs_init", eng.sp_dt, X.data_ptr(), P, N, ld, d_lower.data_ptr(), d_upper.data_ptr(), seed64, None,
               None, eng.stream)
    # normalize to the unit cube, _na.py:160,169 (device buffers, elementwise plumbing)
    t_mask = torch.from_numpy(span_mask).to(eng.device)
    t_span = torch.from_numpy(span.astype(eng.np_dt)).to(eng.device)
    X[:, :N] = torch.where(t_mask, (X[:, :N] - d_lower[:N]) / t_span, d_upper[:N])
    pbest.copy_(X)

    def evaluate(dst):
        eng.evaluate(fun, args, obj, X, P, N, dst, d_scale, d_shift, to_user=unnorm)

    evaluate(pbestfit)
    pfit.copy_(pbestfit)
    L.call("sp_best_init", eng.sp_dt, X.data_ptr(), pbestfit.data_ptr(), P, N, ld, gbest.data_ptr(), ctrl.data_ptr(),
           scratch.data_ptr(), eng.stream)
    M = 0
    L.call("sp_na_append", eng.sp_dt, archT.data_ptr(), cap, M, X.data_ptr(), P, N, ld, eng.stream)
    archfit[M:M + P].copy_(pfit)
    M += P

    hist = History(return_all, maxiter, P, N, verbosity)
    observe = hist.enabled or callback is not None

    def snapshot(it):
        c = eng.read_ctrl(ctrl)
        if not observe:
            return c
        xbest = gbest[:N].to("cpu").numpy().astype(np.float64)
        Xh = eng.download_rows(X, P, N)
        # quirk kept: the first xall entry is the normalised population (_na.py:186-187)
        hist.put(it, Xh if it == 1 else unnorm(Xh), pfit.to("cpu").numpy().astype(np.float64),
                 xbest if it == 1 else None, c.gfit)
        if callback is not None:
            res = OptimizeResult(x=unnorm(xbest), fun=c.gfit, nfev=it * P, nit=it)
            hist.into(res, it)
            callback(unnorm(Xh), res)
        return c

    c = snapshot(1)
    it = 1
    while c.status == L.SP_RUNNING:
        it += 1
        L.call("sp_fitness_rank", eng.sp_dt, archfit.data_ptr(), M, rank.data_ptr(), eng.stream)
        L.call("sp_na_cells", rank.data_ptr(), M, nr, cells.data_ptr(), eng.stream)
        d_u = None
        if stream is not None:  # one uniform per active coordinate, individual-major (_na.py:299)
            u = np.zeros((P, N))
            u[:, span_mask] = stream.rs.random_sample((P, int(span_mask.sum())))
            d_u = eng.upload_rows(u)
        L.call("sp_na_resample", eng.sp_dt, archT.data_ptr(), cap, M, cells.data_ptr(), nr, X.data_ptr(), P, N, ld,
               d_mask.data_ptr(), None if d_u is None else d_u.data_ptr(), seed64, it, work.data_ptr(), eng.stream)
        evaluate(pfit)
        L.call("sp_select_sync", eng.sp_dt, it, int(maxiter), float(xtol), float(ftol), X.data_ptr(), pfit.data_ptr(),
               pbest.data_ptr(), pbestfit.data_ptr(), P, N, ld, 1, gbest.data_ptr(), ctrl.data_ptr(),
               scratch.data_ptr(), eng.stream)
        if M + P <= cap:
            L.call("sp_na_append", eng.sp_dt, archT.data_ptr(), cap, M, X.data_ptr(), P, N, ld, eng.stream)
            archfit[M:M + P].copy_(pfit)
            M += P
        c = snapshot(it)

    it = c.nit
    res = OptimizeResult(
        x=unnorm(gbest[:N].to("cpu").numpy().astype(np.float64)),
        success=c.status >= 0,
        status=int(c.status),
        message=messages[int(c.status)],
        fun=float(c.gfit),
        nfev=it * P,
        nit=it,
    )
    hist.into(res, it)
    return res


register("na", minimize)
