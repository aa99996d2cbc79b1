"""stochopy_b200: the population hot path of stochopy.optimize.minimize() on B200.

Drop-in for ``stochopy.optimize`` / ``stochopy.factory`` (same call signatures,
method strings, result fields and errors); every generation runs in hand-written
sm_100a CUDA kernels behind the C ABI of include/stochopy_b200.h.
"""
from . import factory, optimize

__version__ = "0.1.0"
__all__ = ["factory", "optimize", "__version__"]
