"""nmrfit_b200 - B200-native objective-evaluation hot path of pnnl/nmrfit behind
nmrfit's own API (``fit(data, lower, upper)`` -> ``FitUtility`` ->
``generate_result(scale)``).

    import nmrfit_b200 as nmrfit

Importing the package does not touch CUDA; the shared library is loaded on first
use and there is no CPU fallback.
"""
from . import containers
from . import utils
from . import equations
from . import proc_autophase
from .core import *

__version__ = '0.1.0'
