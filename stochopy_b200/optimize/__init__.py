"""stochopy.optimize-compatible surface (stochopy/optimize/__init__.py:1-18)."""
from ._helpers import OptimizeResult, minimize, register
from ._cmaes import minimize as cmaes
from ._cpso import minimize as cpso
from ._de import minimize as de
from ._na import minimize as na
from ._pso import minimize as pso
from ._vdcma import minimize as vdcma

__all__ = ["OptimizeResult", "minimize", "register", "cmaes", "cpso", "de", "na", "pso", "vdcma"]
