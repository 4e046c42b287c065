"""drjacoby_b200 -- B200-native (sm_100a) implementation of the drjacoby MCMC sampling core."""
from ._lib import DrjacobyError  # noqa: F401
from .engine import Model, RunSpec, Session, run_raw, builtin_models, device_count, fp64_peak_tflops, release_cached_memory  # noqa: F401
from .api import (CudaSource, DrjacobyOutput, check_params, cuda_template, define_params, run_mcmc,  # noqa: F401
                  sample_chains, source_cuda, use_builtin)
from .diagnostics import gelman_rubin  # noqa: F401
from .rcompat import RStream  # noqa: F401

__version__ = "0.1.0"
