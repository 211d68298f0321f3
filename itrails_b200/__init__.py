"""itrails_b200 — B200-native (sm_100a) coalescent-HMM hot path of iTRAILS.

Mirrors the reference's Python API for the path (same function names, argument
meaning and return types):

    trans_emiss_calc            get_trans_emiss.py:8
    loglik_wrapper(_par)        optimizer.py:40,93
    post_prob_wrapper           optimizer.py:241
    viterbi_wrapper             optimizer.py:357
    maf_parser, parse_coordinates, get_obs_state_dct, get_idx_state   read_data.py
    optimizer, optimization_wrapper                                   optimizer.py:396,586

Everything numeric runs in hand-written CUDA behind a C ABI
(include/itrails_b200.h, itrails_b200/lib/libitrails_b200.so).  There is no CPU
fallback: importing the package is cheap, but any compute call raises
``ItrailsCudaError`` if the library or a B200-class GPU is missing.
"""
import os as _os

# ~20 concurrent CUDA streams (recursions, posterior length groups, copies): more hardware
# work queues than the default 8, or they alias and serialise.  Must be set before the CUDA
# context exists; libitrails_b200.so sets it too when it is loaded first.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

__version__ = "0.1.0"

from ._lib import ItrailsCudaError, ItrailsError  # noqa: F401
from .engine import Engine  # noqa: F401
from .read_data import (get_idx_state, get_obs_state_dct, maf_parser,  # noqa: F401
                        parse_coordinates)
from .cutpoints import cutpoints_AB, cutpoints_ABC, get_times  # noqa: F401
# (the function `optimizer` is not re-exported: it would shadow the module of the same name,
# which — as in the reference — is imported as `from itrails_b200.optimizer import optimizer`)
from .optimizer import (loglik_wrapper, loglik_wrapper_par, post_prob_wrapper,  # noqa: F401
                        viterbi_wrapper, optimization_wrapper)
from .get_trans_emiss import trans_emiss_calc  # noqa: F401
