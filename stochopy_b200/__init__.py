"""stochopy_b200: the population hot path of stochopy.optimize.minimize() on B200.

Drop-in for ``stochopy.optimize`` / ``stochopy.factory`` (same call signatures,
method strings, result fields and errors); every generation runs in hand-written
sm_100a CUDA kernels behind the C ABI of include/stochopy_b200.h.
"""
from . import factory, optimize
from .jit import JitObjective, jit_objective

__version__ = "0.1.0"
__all__ = ["factory", "optimize", "jit_objective", "JitObjective", "__version__"]
