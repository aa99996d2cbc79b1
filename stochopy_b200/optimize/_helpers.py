"""API dispatch: same surface as stochopy/optimize/_helpers.py:8-94."""

__all__ = ["minimize", "OptimizeResult", "register"]

_optimizer_map = {}


class OptimizeResult(dict):
    """Optimization result: a dict with attribute access (stochopy/_common.py:1-35).

    Fields: x, success, status, message, fun, nfev, nit (+ xall, funall with
    ``return_all``); the repr hides xall/funall like the reference's."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__

    def __repr__(self):
        if not self:
            return self.__class__.__name__ + "()"
        width = max(len(k) for k in self) + 1
        shown = sorted((k, v) for k, v in self.items() if k not in ("xall", "funall"))
        return "\n".join(k.rjust(width) + ": " + repr(v) for k, v in shown)

    def __dir__(self):
        return list(self.keys())


def register(name, minimize):
    """Register an optimizer under a method string (_helpers.py:39-41)."""
    _optimizer_map[name] = minimize


def minimize(fun, bounds, x0=None, args=(), method="de", options=None, callback=None):
    """Minimize ``fun`` with a population method on the GPU.

    Same signature, method strings ('cmaes', 'cpso', 'de', 'na', 'pso', 'vdcma'),
    option names, result fields and exception types as
    ``stochopy.optimize.minimize`` (_helpers.py:44-94).  Extra options accepted
    by every method: ``dtype`` ('float64' | 'float32'), ``device``, ``rng``
    ('philox': counter-based draws in the kernels; 'numpy': the reference's
    MT19937 draw order from the host, for fixed-seed trajectory parity).
    An unknown method raises KeyError, like the reference."""
    options = dict(options) if options else {}
    fn = _optimizer_map[method]
    if options.get("backend") == "mpi" and options.get("seed") is None:
        # the reference's mpi backend broadcasts rank 0's population every generation; here all ranks run
        # the same optimiser, so they must share one random stream: rank 0's seed
        from ..parallel import shared_seed

        options["seed"] = shared_seed(None)
    from .._lib import nvtx_range

    with nvtx_range(f"stochopy_b200.minimize[{method}]"):
        return fn(fun=fun, bounds=bounds, x0=x0, args=args, callback=callback, **options)
