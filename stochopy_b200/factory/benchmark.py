"""The seven benchmark objectives of stochopy.factory (benchmark.py:1-156).

Each is a callable ``f(x) -> float`` for a 1-D ``x`` -- the reference's
contract -- evaluated by the CUDA kernel ``sp_eval`` (one row).  Passed as
``fun`` to ``optimize.minimize`` they are recognised (``_sp_objective``) and the
whole population is evaluated inside the fused generation kernels instead.
"""
import numpy as np

from .. import _lib as L

__all__ = ["ackley", "griewank", "quartic", "rastrigin", "rosenbrock", "sphere", "styblinski_tang"]


class DeviceObjective:
    def __init__(self, name):
        self.__name__ = name
        self.__qualname__ = name
        self._sp_objective = L.OBJECTIVES[name]
        self.__doc__ = f"The {name.replace('_', '-').title()} function (evaluated on the GPU)."

    def __call__(self, x):
        return float(self.batch(np.asarray(x, dtype=np.float64).reshape(1, -1))[0])

    def batch(self, X, dtype="float64"):
        """f[i] = fun(X[i]) for a host (P, N) array, on the device."""
        import torch

        from ..optimize._common import Engine

        eng = Engine(dtype)
        X = np.asarray(X)
        p, n = X.shape
        dX = eng.upload_rows(X)
        out = eng.empty(p)
        L.call("sp_eval", self._sp_objective, eng.sp_dt, dX.data_ptr(), p, n, dX.shape[1], None, None,
               out.data_ptr(), eng.stream)
        return out.to("cpu").numpy().astype(np.float64)

    def __repr__(self):
        return f"<stochopy_b200.factory.{self.__name__}>"


ackley = DeviceObjective("ackley")
griewank = DeviceObjective("griewank")
quartic = DeviceObjective("quartic")
rastrigin = DeviceObjective("rastrigin")
rosenbrock = DeviceObjective("rosenbrock")
sphere = DeviceObjective("sphere")
styblinski_tang = DeviceObjective("styblinski_tang")
