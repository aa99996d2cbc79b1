"""User objectives as CUDA C source, compiled at run time with NVRTC.

The reference's contract is ``fun(x, *args) -> float`` on a 1-D array, called once per
individual (stochopy/optimize/_common.py:27-106).  An arbitrary Python callable keeps
that contract through a device -> host -> device round trip per generation.  An
objective written as a CUDA device function keeps the whole generation on the GPU:

    rosen = stochopy_b200.jit_objective('''
    __device__ real objective(const real* x, int n) {
      real s = 0;
      for (int i = 0; i + 1 < n; ++i) {
        const real a = x[i + 1] - x[i] * x[i], b = 1 - x[i];
        s += 100 * a * a + b * b;
      }
      return s;
    }''')
    res = stochopy_b200.optimize.minimize(rosen, bounds, method="de", options={...})

``real`` is ``float`` or ``double`` according to the run's ``dtype``.  The object is
still a callable with the reference's signature (``rosen(x)`` evaluates one point on
the device), so ``callback`` code and post-processing that call ``fun(res.x)`` keep working.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L

__all__ = ["JitObjective", "jit_objective"]


class JitObjective:
    _sp_jit = True

    def __init__(self, source):
        if not isinstance(source, str) or "objective" not in source:
            raise ValueError("source must define `__device__ real objective(const real* x, int n)`")
        self.source = source
        self._handles = {}  # (sp_dtype, device index) -> handle

    def check(self, dtype="float64"):
        """Compile only (works without a GPU); raises EngineError with the NVRTC log on errors."""
        n = C.c_int64()
        L.call("sp_jit_check", self.source.encode(), L.SP_F32 if np.dtype(dtype) == np.float32 else L.SP_F64, C.byref(n))
        return int(n.value)

    def handle(self, sp_dt, device_index):
        key = (int(sp_dt), int(device_index))
        h = self._handles.get(key)
        if h is None:
            out = C.c_void_p()
            with torch.cuda.device(device_index):
                L.call("sp_jit_compile", self.source.encode(), int(sp_dt), C.byref(out))
            h = self._handles[key] = out.value
        return h

    def evaluate_rows(self, eng, X, p, n, out, scale=None, shift=None):
        """out[:p] = objective(X[i] * scale + shift) on the device (X: padded device rows)."""
        L.call("sp_jit_eval", C.c_void_p(self.handle(eng.sp_dt, eng.device.index)), eng.sp_dt, X.data_ptr(), p, n,
               X.shape[1], None if scale is None else scale.data_ptr(), None if shift is None else shift.data_ptr(),
               out.data_ptr(), eng.stream)

    def __call__(self, x, *args):
        if args:
            raise ValueError("a JIT objective takes no extra arguments (bake them into the source)")
        from .optimize._common import Engine

        eng = Engine("float64")
        x = np.asarray(x, dtype=np.float64).reshape(1, -1)
        X = eng.upload_rows(x)
        out = eng.empty(1)
        self.evaluate_rows(eng, X, 1, x.shape[1], out)
        return float(out.item())

    def __del__(self):
        try:
            for h in self._handles.values():
                L.load().sp_jit_free(C.c_void_p(h))
        except Exception:
            pass


def jit_objective(source):
    """CUDA C source defining ``__device__ real objective(const real* x, int n)`` -> objective usable
    with every method of ``stochopy_b200.optimize.minimize`` (evaluated on the device)."""
    return JitObjective(source)
