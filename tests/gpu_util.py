"""Helpers for the GPU parity tests: drive the C ABI directly with torch buffers."""
import ctypes as C

import numpy as np
import torch

from stochopy_b200 import _lib as L
from stochopy_b200.optimize._common import Engine


def tol_for(dtype):
    """(rtol on positions, fitness tolerance factor on sum|terms|) per dtype.
    fp64: the north-star's 1e-6 relative is the contract; we hold 1e-11.
    fp32: |df| <= 2e-6 * sum|terms| (24-bit arithmetic, ~N accumulated roundings)."""
    return (1e-12, 1e-13) if np.dtype(dtype) == np.float64 else (2e-6, 4e-6)


class DeRig:
    """One DE population on the device, steppable through sp_de_generation."""

    def __init__(self, X, pbestfit, gbest, objective, strategy, constraint, F, CR, lower, upper, seed=0,
                 maxiter=1000, xtol=1e-8, ftol=1e-8, dtype="float64", it0=2):
        self.eng = eng = Engine(dtype)
        P, N = X.shape
        self.P, self.N, self.ld = P, N, eng.ld(N)
        self.X = [eng.rows(P, N), eng.rows(P, N)]
        eng.upload_rows(X, out=self.X[it0 & 1])
        self.pbestfit = eng.upload_vec(pbestfit)
        self.pfit = eng.empty(P)
        self.gbest = eng.upload_vec(gbest, self.ld)
        self.lower, self.upper = eng.upload_vec(lower, self.ld), eng.upload_vec(upper, self.ld)
        self.ctrl, self.scratch = eng.new_ctrl()
        st = self.st = L.DeState()
        st.dtype, st.objective = eng.sp_dt, L.OBJECTIVES[objective]
        st.strategy, st.constraint = L.DE_STRATEGIES[strategy], (L.CONS_RANDOM if constraint == "Random" else L.CONS_NONE)
        st.P, st.N, st.maxiter, st.ld = P, N, maxiter, self.ld
        st.F, st.CR, st.xtol, st.ftol, st.seed = F, CR, xtol, ftol, seed
        st.X[0], st.X[1] = self.X[0].data_ptr(), self.X[1].data_ptr()
        st.pbestfit, st.pfit, st.gbest = self.pbestfit.data_ptr(), self.pfit.data_ptr(), self.gbest.data_ptr()
        st.lower, st.upper = self.lower.data_ptr(), self.upper.data_ptr()
        st.ctrl, st.scratch = self.ctrl.data_ptr(), self.scratch.data_ptr()
        self._keep = None

    def draws(self, r1, donors, irand, repair):
        eng = self.eng
        self._keep = (eng.upload_rows(r1), torch.from_numpy(np.ascontiguousarray(donors, dtype=np.int64)).to(eng.device),
                      torch.from_numpy(np.ascontiguousarray(irand, dtype=np.int64)).to(eng.device),
                      None if repair is None else eng.upload_rows(repair))
        st = self.st
        st.r1, st.donors, st.irand = (t.data_ptr() for t in self._keep[:3])
        st.repair = None if repair is None else self._keep[3].data_ptr()

    def step(self, it):
        L.call("sp_de_generation", C.byref(self.st), it, self.eng.stream)
        self.eng.sync()

    def get(self, it):
        """State after generation `it`."""
        e = self.eng
        return dict(X=e.download_rows(self.X[(it & 1) ^ 1], self.P, self.N),
                    pbestfit=self.pbestfit.cpu().numpy().astype(np.float64),
                    pfit=self.pfit.cpu().numpy().astype(np.float64),
                    gbest=self.gbest[: self.N].cpu().numpy().astype(np.float64), ctrl=e.read_ctrl(self.ctrl))


class PsoRig:
    def __init__(self, X, V, pbest, pbestfit, gbest, objective, constraint, w, c1, c2, lower, upper, seed=0,
                 maxiter=1000, xtol=1e-8, ftol=1e-8, dtype="float64", gamma=None, delta=0.0):
        self.eng = eng = Engine(dtype)
        P, N = X.shape
        self.P, self.N, self.ld = P, N, eng.ld(N)
        self.X, self.V, self.pbest = eng.upload_rows(X), eng.upload_rows(V), eng.upload_rows(pbest)
        self.pbestfit = eng.upload_vec(pbestfit)
        self.pfit = eng.empty(P)
        self.gbest = eng.upload_vec(gbest, self.ld)
        self.lower, self.upper = eng.upload_vec(lower, self.ld), eng.upload_vec(upper, self.ld)
        self.ctrl, self.scratch = eng.new_ctrl()
        self.rank = eng.zeros(P, dtype=torch.int32)
        st = self.st = L.PsoState()
        st.dtype, st.objective = eng.sp_dt, L.OBJECTIVES[objective]
        st.constraint = L.CONS_SHRINK if constraint == "Shrink" else L.CONS_NONE
        st.P, st.N, st.maxiter, st.ld = P, N, maxiter, self.ld
        st.w, st.c1, st.c2, st.xtol, st.ftol, st.seed = w, c1, c2, xtol, ftol, seed
        st.gamma, st.delta = (-1.0 if gamma is None else gamma), delta
        st.X, st.V, st.pbest = self.X.data_ptr(), self.V.data_ptr(), self.pbest.data_ptr()
        st.pbestfit, st.pfit, st.gbest = self.pbestfit.data_ptr(), self.pfit.data_ptr(), self.gbest.data_ptr()
        st.lower, st.upper = self.lower.data_ptr(), self.upper.data_ptr()
        st.ctrl, st.scratch = self.ctrl.data_ptr(), self.scratch.data_ptr()
        self._keep = None

    def draws(self, r1, r2):
        self._keep = (self.eng.upload_rows(r1), self.eng.upload_rows(r2))
        self.st.r1, self.st.r2 = self._keep[0].data_ptr(), self._keep[1].data_ptr()

    def step(self, it):
        L.call("sp_pso_generation", C.byref(self.st), it, self.eng.stream)
        self.eng.sync()

    def restart(self, it, fresh=None):
        e = self.eng
        L.call("sp_cpso_restart_plan", C.byref(self.st), it, self.rank.data_ptr(), e.stream)
        nw = e.read_ctrl(self.ctrl).flag
        f = None if fresh is None else e.upload_rows(fresh(nw))
        L.call("sp_cpso_restart_apply", C.byref(self.st), it, self.rank.data_ptr(),
               None if f is None else f.data_ptr(), e.stream)
        e.sync()
        return nw

    def get(self):
        e = self.eng
        g = lambda t: e.download_rows(t, self.P, self.N)
        return dict(X=g(self.X), V=g(self.V), pbest=g(self.pbest),
                    pbestfit=self.pbestfit.cpu().numpy().astype(np.float64),
                    pfit=self.pfit.cpu().numpy().astype(np.float64),
                    gbest=self.gbest[: self.N].cpu().numpy().astype(np.float64), ctrl=e.read_ctrl(self.ctrl))


def device_eval(name, X, dtype="float64", scale=None, shift=None):
    eng = Engine(dtype)
    P, N = X.shape
    dX = eng.upload_rows(X)
    out = eng.empty(P)
    ds = None if scale is None else eng.upload_vec(scale, eng.ld(N))
    dh = None if shift is None else eng.upload_vec(shift, eng.ld(N))
    L.call("sp_eval", L.OBJECTIVES[name], eng.sp_dt, dX.data_ptr(), P, N, dX.shape[1],
           None if ds is None else ds.data_ptr(), None if dh is None else dh.data_ptr(), out.data_ptr(), eng.stream)
    return out.cpu().numpy().astype(np.float64)


def device_lhs(P, N, bounds, seed, dtype="float64", jitter=None, perm=None):
    eng = Engine(dtype)
    ld = eng.ld(N)
    X = eng.rows(P, N)
    lo, hi = eng.upload_vec(bounds[:, 0], ld), eng.upload_vec(bounds[:, 1], ld)
    if jitter is None:
        L.call("sp_lhs_init", eng.sp_dt, X.data_ptr(), P, N, ld, lo.data_ptr(), hi.data_ptr(), seed, None, None, eng.stream)
    else:
        dj = eng.upload_rows(jitter)
        dp = torch.from_numpy(np.ascontiguousarray(perm.T, dtype=np.int64)).to(eng.device)
        L.call("sp_lhs_init", eng.sp_dt, X.data_ptr(), P, N, ld, lo.data_ptr(), hi.data_ptr(), 0, dj.data_ptr(),
               dp.data_ptr(), eng.stream)
    return eng.download_rows(X, P, N)
