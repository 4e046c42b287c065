"""drjacoby_b200 -- B200-native (sm_100a) implementation of the drjacoby MCMC sampling core."""
from .engine import Model, RunSpec, Session, run_raw, builtin_models, device_count, fp64_peak_tflops  # noqa: F401
from ._lib import DrjacobyError  # noqa: F401

__version__ = "0.1.0"
