"""Host-side plumbing shared by the method front-ends.

Mirrors the role of the reference's ``@optimizer`` decorator
(stochopy/optimize/_common.py:27-106): it turns the per-individual ``fun(x)``
contract into a population evaluation -- here either a device kernel (when
``fun`` is one of the factory objectives) or a device->host->device round trip
that calls the user's Python callable row by row (the exact reference contract).
PyTorch is used for device buffers and streams only.
"""
import ctypes as C
import functools
import threading
import os

import numpy as np
import torch

from .. import _lib as L

# stochopy/optimize/_common.py:12-24
messages = {
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

_TORCH_DT = {"float32": torch.float32, "float64": torch.float64}
_PINNED = {}  # one pinned landing pad for the control block per (device, host thread): concurrent runs do not share it


def resolve_dtype(dtype):
    name = np.dtype(dtype).name
    if name not in _TORCH_DT:
        raise ValueError()
    return np.dtype(name), _TORCH_DT[name], (L.SP_F32 if name == "float32" else L.SP_F64)


def device_objective(fun, args):
    """Objective id if ``fun`` is a factory objective (ours or the reference's), else None."""
    if args:
        return None
    oid = getattr(fun, "_sp_objective", None)
    if oid is not None:
        return oid
    mod = getattr(fun, "__module__", "") or ""
    name = getattr(fun, "__name__", "")
    if mod == "stochopy.factory.benchmark" and name in L.OBJECTIVES:
        return L.OBJECTIVES[name]
    return None


def device_scope(fn):
    """Run a front-end with its ``device=`` option as the current CUDA device.  The C library
    launches on the current device (it only ever calls cudaGetDevice), so a run on "cuda:1" while
    device 0 is current must switch for its whole duration; the previous device is restored on
    return.  Without the option (or without CUDA: the engine then raises) nothing happens."""

    @functools.wraps(fn)
    def wrapper(*args, **kw):
        dev = kw.get("device")
        if dev is None or not torch.cuda.is_available():
            return fn(*args, **kw)
        d = torch.device("cuda", dev) if isinstance(dev, int) else torch.device(dev)
        if d.type != "cuda":
            raise ValueError()
        idx = torch.cuda.current_device() if d.index is None else d.index
        kw["device"] = torch.device("cuda", idx)
        with torch.cuda.device(idx):
            return fn(*args, **kw)

    return wrapper


def fresh_seed(seed):
    if seed is None:
        return int.from_bytes(os.urandom(8), "little")
    return int(seed) & 0xFFFFFFFFFFFFFFFF


class Engine:
    """Device context of one optimiser run: device, dtype, stream, buffers."""

    def __init__(self, dtype="float64", device=None, backend=None):
        L.load()
        self.backend = backend
        if not torch.cuda.is_available():
            raise L.EngineError("stochopy_b200 needs a CUDA device (no CPU fallback)")
        self.np_dt, self.t_dt, self.sp_dt = resolve_dtype(dtype)
        dev = torch.device("cuda") if device is None else (torch.device("cuda", device) if isinstance(device, int)
                                                           else torch.device(device))
        if dev.type != "cuda":
            raise ValueError()
        # always an explicit index ("cuda" = the current device); the C library launches on the current
        # device, so the front-ends run under `device_scope` which makes this device current
        self.device = torch.device("cuda", torch.cuda.current_device() if dev.index is None else dev.index)
        self.vec = 16 // self.np_dt.itemsize
        key = (self.device.type, self.device.index, threading.get_ident())
        if key not in _PINNED:
            _PINNED[key] = torch.empty(512, dtype=torch.uint8, pin_memory=True)
        self._ctrl_host = _PINNED[key]

    # -- buffers ----------------------------------------------------------------
    def ld(self, n):
        return (n + self.vec - 1) // self.vec * self.vec

    def empty(self, *shape, dtype=None):
        return torch.empty(shape, dtype=dtype or self.t_dt, device=self.device)

    def zeros(self, *shape, dtype=None):
        return torch.zeros(shape, dtype=dtype or self.t_dt, device=self.device)

    def rows(self, p, n):
        """Zero-filled (p, ld) row buffer."""
        return self.zeros(p, self.ld(n))

    def rows_scratch(self, p, n):
        """(p, ld) row buffer that is about to be overwritten: only the padding columns (if any)
        have to be zero, so an unpadded buffer is left uninitialised."""
        return self.empty(p, n) if self.ld(n) == n else self.zeros(p, self.ld(n))

    def upload_rows(self, host, out=None):
        """Host (p, n) array -> padded device rows (copy; the caller's array is never
        aliased).  fp32/fp64 sources travel as they are and are cast on the device.  A source in
        page-locked memory (e.g. the numpy view of a pinned torch tensor) is copied asynchronously by
        DMA straight into the row buffer; the caller keeps it alive until the run returns."""
        host = np.asarray(host)
        if host.dtype not in (np.float32, np.float64):
            host = host.astype(np.float64)
        host = np.ascontiguousarray(host)
        p, n = host.shape
        out = self.rows_scratch(p, n) if out is None else out
        src = torch.from_numpy(host)
        pinned = src.is_pinned()
        if out.shape[1] == n:  # unpadded rows: one copy (and the dtype cast, if any, on the device side of it)
            out.copy_(src, non_blocking=pinned)
        else:
            out[:, :n].copy_(src.to(self.device, non_blocking=pinned))
        return out

    def upload_vec(self, host, pad_to=None, dtype=None):
        host = np.ascontiguousarray(host, dtype=dtype or self.np_dt)
        n = host.shape[0]
        out = torch.zeros(pad_to or n, dtype=torch.from_numpy(host).dtype, device=self.device)
        out[:n].copy_(torch.from_numpy(host))
        return out

    def download_rows(self, dev, p, n):
        return dev[:p, :n].to("cpu").numpy().astype(np.float64)

    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def sync(self):
        torch.cuda.current_stream(self.device).synchronize()

    # -- control block ------------------------------------------------------------
    def new_ctrl(self):
        host = L.Ctrl()
        host.status = L.SP_RUNNING
        raw = np.frombuffer(bytes(host), dtype=np.uint8).copy()
        ctrl = torch.from_numpy(raw).to(self.device)
        scratch = torch.zeros(int(L.load().sp_scratch_bytes()), dtype=torch.uint8, device=self.device)
        return ctrl, scratch

    def read_ctrl(self, ctrl, cls=L.Ctrl):
        """Device control block -> host mirror (one small D2H through pinned memory)."""
        n = C.sizeof(cls)
        self._ctrl_host[:n].copy_(ctrl[:n], non_blocking=True)
        self.sync()
        return cls.from_buffer_copy(self._ctrl_host[:n].numpy().tobytes())

    def new_struct(self, host):
        """Upload a ctypes struct as a device byte buffer."""
        raw = np.frombuffer(bytes(host), dtype=np.uint8).copy()
        return torch.from_numpy(raw).to(self.device)

    # -- population evaluation (the reference's `fun(X)` wrapper) -----------------------
    def evaluate(self, fun, args, obj, X, p, n, out, scale=None, shift=None, to_user=None, clip=False):
        """out[:p] = fun(X[i]).  Device kernel for factory objectives and for objectives
        compiled from CUDA source (jit.py); otherwise the reference's per-individual
        Python contract via a host round trip.  scale/shift (device) and clip describe on
        the device what ``to_user`` does on the host: x -> clip(x, -1, 1) * scale + shift."""
        if obj is not None:
            L.call("sp_eval", obj, self.sp_dt, X.data_ptr(), p, n, X.shape[1],
                   None if scale is None else scale.data_ptr(), None if shift is None else shift.data_ptr(),
                   out.data_ptr(), self.stream)
            return
        if getattr(fun, "_sp_jit", False):
            if args:
                raise ValueError()
            fun.evaluate_rows(self, X.clamp(-1.0, 1.0) if clip else X, p, n, out, scale, shift)
            return
        host = self.download_rows(X, p, n)
        if to_user is not None:
            host = to_user(host)
        if self.backend == "mpi":  # the reference's rank-strided evaluation, over torch.distributed
            from ..parallel import evaluate_split

            f = evaluate_split(fun, args, host)
        else:
            f = np.array([fun(row, *args) for row in host], dtype=np.float64)
        out[:p].copy_(torch.from_numpy(f.astype(self.np_dt)))


class NumpyStream:
    """rng="numpy": draw from numpy's legacy MT19937 generator in the reference's
    order so a fixed seed follows the reference's trajectory.
    Host-side draws are uploaded; the kernels consume them instead of Philox."""

    def __init__(self, seed):
        self.rs = np.random.RandomState(seed)

    def lhs(self, P, N):  # _common.py:111,113
        jitter = self.rs.uniform(size=(P, N))
        perm = np.stack([self.rs.permutation(P) for _ in range(N)], axis=0)  # (N, P)
        return jitter, perm

    def de(self, P, N, k, lower, upper, repair):  # _de.py:250, 306/311, 340; de/_constraints.py:24
        r1 = self.rs.rand(P, N)
        donors = np.empty((k, P), dtype=np.int64)
        base = np.arange(P)
        for i in range(P):
            donors[:, i] = self.rs.permutation(np.delete(base, i))[:k]
        irand = self.rs.randint(N, size=P).astype(np.int64)
        rep = self.rs.uniform(lower, upper, (P, N)) if repair else None
        return r1, donors, irand, rep

    def pso(self, P, N):  # _cpso.py:262-263
        return self.rs.rand(P, N), self.rs.rand(P, N)

    def restart(self, nw, N, lower, upper):  # _cpso.py:422
        return self.rs.uniform(lower, upper, (nw, N))

    def mean0(self, N):  # _cmaes.py:180, _vdcma.py:181
        return self.rs.uniform(-1.0, 1.0, N)

    def normal(self, *shape):  # _cmaes.py:234, _vdcma.py:208,239,246
        return self.rs.randn(*shape)


class History:
    """xall / funall of `return_all` (_de.py:221-234, 270-278)."""

    def __init__(self, enabled, maxiter, P, N, verbosity):
        self.enabled = bool(enabled)
        if self.enabled:
            self.nout = int(np.ceil(verbosity * P))
            w = max(1, self.nout)
            self.xall = np.empty((maxiter, w, N))
            self.funall = np.empty((maxiter, w))

    def put(self, it, X, pfit, gbest=None, gfit=None):
        if not self.enabled:
            return
        if self.nout > 0:
            self.xall[it - 1] = X[: self.nout]
            self.funall[it - 1] = pfit[: self.nout]
        elif gbest is not None:
            self.xall[it - 1] = gbest
            self.funall[it - 1] = gfit
        else:
            b = int(np.argmin(pfit))
            self.xall[it - 1] = X[b]
            self.funall[it - 1] = pfit[b]

    def into(self, res, it):
        if self.enabled:
            res.update({"xall": self.xall[:it], "funall": self.funall[:it]})


class HistoryStreamer:
    """`return_all` without stalling the generation loop (reference
    `_de.py:221-234,270-278`, `_cpso.py:231-244`): after every generation the first `nout`
    rows and their fitness are snapshotted device-to-device into a small ring on the
    launching stream, and a side stream moves the ring slots to pinned host memory while
    the next generations run.  An event per slot keeps a slot from being overwritten
    before its copy has left the device.  The pinned host side is a ring too (a window of
    `win` generations, at most PIN_LIMIT bytes): when it wraps, the host waits for the copy
    event of the generation it is about to overwrite -- `win` generations in the past, so
    the device queue stays full -- and moves that part of the window into the History
    arrays.  Nothing here synchronises with the host per generation."""

    SLOTS = 4
    PIN_LIMIT = 256 << 20  # bytes of pinned window

    @classmethod
    def maybe(cls, eng, hist, callback, P, N):
        if not hist.enabled or callback is not None or hist.nout <= 0:
            return None
        return cls(eng, hist, N)

    def __init__(self, eng, hist, N):
        self.eng, self.N, self.nout, self.hist = eng, N, hist.nout, hist
        self.main = torch.cuda.current_stream(eng.device)
        self.side = torch.cuda.Stream(device=eng.device)
        maxiter = hist.xall.shape[0]
        per_gen = self.nout * (N + 1) * eng.np_dt.itemsize
        self.win = int(max(self.SLOTS, min(maxiter, self.PIN_LIMIT // max(1, per_gen))))
        self.slotX = [eng.empty(self.nout, N) for _ in range(self.SLOTS)]
        self.slotF = [eng.empty(self.nout) for _ in range(self.SLOTS)]
        self.hX = torch.empty((self.win, self.nout, N), dtype=eng.t_dt, pin_memory=True)
        self.hF = torch.empty((self.win, self.nout), dtype=eng.t_dt, pin_memory=True)
        self.free = [None] * self.SLOTS
        self.done = {}  # generation -> event of its D2H copy (generations still in the window)
        self.first = None
        self.last = None

    def _drain(self, upto):
        """Move generations <= upto from the pinned window into the History arrays."""
        for it in sorted(g for g in self.done if g <= upto):
            self.done.pop(it).synchronize()
            h = (it - 1) % self.win
            self.hist.xall[it - 1] = self.hX[h].numpy()
            self.hist.funall[it - 1] = self.hF[h].numpy()

    def push(self, it, X, fit):
        """Enqueue the snapshot of generation `it` (X: device rows, fit: device vector)."""
        k = it % self.SLOTS
        if it - self.win in self.done:  # the window wraps onto a generation that has not been moved out yet
            self._drain(it - self.win + max(0, self.win // 2 - 1))  # half a window at a time
        if self.free[k] is not None:
            self.main.wait_event(self.free[k])
        self.slotX[k].copy_(X[: self.nout, : self.N])
        self.slotF[k].copy_(fit[: self.nout])
        ready = torch.cuda.Event()
        ready.record(self.main)
        h = (it - 1) % self.win
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            self.hX[h].copy_(self.slotX[k], non_blocking=True)
            self.hF[h].copy_(self.slotF[k], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.side)
        self.free[k] = done
        self.done[it] = done
        if self.first is None:
            self.first = it
        self.last = it

    def finish(self, hist, nit, transform=None):
        """Wait for the copies and hand generations first..nit to the History arrays
        (`transform`: host map applied to the float64 rows, e.g. un-standardisation)."""
        self.side.synchronize()
        if self.last is not None:
            self._drain(self.last)
        if transform is not None and self.first is not None and nit >= self.first:
            a, b = self.first - 1, nit
            hist.xall[a:b] = transform(hist.xall[a:b].copy())


def validate_common(fun, bounds, callback):
    """Checks shared by every front-end (same exception types as the reference)."""
    if not hasattr(fun, "__call__"):
        raise TypeError()
    if np.ndim(bounds) != 2:
        raise ValueError()
    if callback is not None and not hasattr(callback, "__call__"):
        raise ValueError()
