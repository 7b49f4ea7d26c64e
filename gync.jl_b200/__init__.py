"""gync.jl_b200 — B200-native hot path of GynC.jl (batched GynCycle likelihood + EM prior update).

Host-side mirror of the reference's Julia API for this path (api.py) over the C-ABI library
(csrc/, include/gync_b200.h).  Import as `gync_b200` (the directory name is not a Python identifier).
"""
from . import lib  # noqa: F401
from . import eb  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import (DEFAULT_ATOL, DEFAULT_RTOL, Config, Lausanne, Patient, Pfizer, Sampling, WeightedChain,  # noqa: F401
                  emiteration, emiteration_bang, emiterations, forwardsol, gync, llh, loglik_matrix, logprior,
                  new_sampling, priory0, referencesolution, sample, sample_bang, sample_chains)
