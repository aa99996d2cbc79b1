from .benchmark import *  # noqa: F401,F403
from .benchmark import __all__  # noqa: F401
