"""ctypes binding of the C-ABI shared library (include/gync_b200.h).

This is the same ABI a Julia `ccall` binds: column-major arrays, Int64/Cint
widths, in-place mutation.  There is NO CPU fallback: if the library is missing or no CUDA
device is present, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GYNC_B200_LIB") or os.path.join(_HERE, "csrc", "libgync_b200.so")   # GYNC_B200_LIB: A/B and profile builds

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class SolverOpts(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("h0", C.c_double), ("hmax", C.c_double),
                ("max_steps", C.c_int32), ("reserved", C.c_int32)]


class Prior(C.Structure):
    _fields_ = [("gmm_means", C.c_double * (31 * 33)), ("gmm_stds", C.c_double * 33),
                ("upper", C.c_double * 82), ("mu_T", C.c_double), ("sd_T", C.c_double)]


class AmmTune(C.Structure):
    _fields_ = [("beta", C.c_double), ("scale", C.c_double), ("m", c_int64_p), ("Mv", c_double_p), ("Mvv", c_double_p),
                ("SigmaLm", c_double_p)]


class GyncError(RuntimeError):
    pass


#: every symbol include/gync_b200.h declares (tests check that the .so exports all of them)
EXPORTS = [
    "gync_init", "gync_shutdown", "gync_last_error", "gync_abi_version", "gync_dims",
    "gync_solve", "gync_forwardsol", "gync_loglik_batch", "gync_loglik_batch_dev",
    "gync_logprior_batch", "gync_mcmc_run", "gync_em_iterate", "gync_em_iterate_dev",
    "gync_em_norms_dev", "gync_em_step_dev", "gync_launch_count", "gync_wc_normalize", "gync_fp64_peak",
    "gync_flops_per_step", "gync_method_order", "gync_em_comm_create", "gync_em_comm_connect", "gync_em_comm_destroy",
    "gync_em_iterate_fused_dev", "gync_em_comm_error", "gync_wc_normalize_dev", "gync_init_multi", "gync_ndev",
    "gync_em_iterate_multi", "gync_mcmc_run_amm", "gync_wc_colstats_dev", "gync_wc_expsum_dev", "gync_wc_rows_dev",
    "gync_wc_scale_dev", "gync_collapse_duplicates", "gync_mcmc_run_ids", "gync_mcmc_run_amm_ids", "gync_mcmc_last_stats", "gync_mcmc_last_cancelled", "gync_mcmc_last_early", "gync_mcmc_last_histogram", "gync_em_create", "gync_em_iterate_handle", "gync_em_destroy",
    "gync_likelihoodmat", "gync_likelihoodmat_dev", "gync_phi_batch", "gync_projectsimplex", "gync_ebmodel_create",
    "gync_ebmodel_from_matrices", "gync_ebmodel_destroy", "gync_ebmodel_dims", "gync_ebmodel_matrices",
    "gync_ebmodel_device_ptrs", "gync_ebmodel_logl", "gync_ebmodel_dlogl", "gync_ebmodel_hz", "gync_ebmodel_dhz",
    "gync_ebmodel_em", "gync_ebmodel_mple", "gync_ebmodel_dhz_dev", "gync_ebmodel_dlogl_dev", "gync_ebmodel_hz_dev",
    "gync_ebmodel_ascent_dev", "gync_ebmodel_nan_flag", "gync_ebmodel_create_zshard", "gync_ebmodel_zshard_finish",
]

_lib = None


def load():
    """Load the shared library (no device needed for loading)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GyncError(f"{LIB_PATH} not found — run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.gync_last_error.restype = C.c_char_p
    lib.gync_launch_count.restype = C.c_int64
    lib.gync_init.argtypes = [C.c_int]
    lib.gync_init_multi.argtypes = [C.c_int32, c_int32_p]
    lib.gync_em_iterate_multi.argtypes = [c_double_p, c_double_p, C.c_int64, C.c_int32, C.c_int32]
    lib.gync_dims.argtypes = [c_int32_p] * 5
    lib.gync_solve.argtypes = [c_double_p, c_double_p, C.c_int64, c_double_p, C.c_int32, C.POINTER(SolverOpts),
                               c_double_p, c_int32_p]
    lib.gync_forwardsol.argtypes = [c_double_p, C.c_int64, c_double_p, C.c_int32, C.POINTER(SolverOpts),
                                    c_double_p, c_int32_p]
    lib.gync_loglik_batch.argtypes = [c_double_p, C.c_int64, c_double_p, C.c_int32, c_double_p,
                                      C.POINTER(SolverOpts), c_double_p, c_double_p, c_int32_p, c_int32_p]
    lib.gync_loglik_batch_dev.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p,
                                          C.POINTER(SolverOpts), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p]
    lib.gync_logprior_batch.argtypes = [c_double_p, C.c_int64, C.POINTER(Prior), c_double_p]
    lib.gync_mcmc_run.argtypes = [c_double_p, c_double_p, C.c_int64, c_double_p, C.c_double, c_double_p, C.c_int32,
                                  c_double_p, c_int32_p, C.POINTER(Prior), C.POINTER(SolverOpts), C.c_int64,
                                  C.c_int32, C.c_uint64, C.c_uint64, c_double_p, c_double_p, c_double_p, c_int64_p]
    lib.gync_mcmc_run_amm.argtypes = [c_double_p, c_double_p, C.c_int64, c_double_p, C.c_double, c_double_p, C.c_int32,
                                      c_double_p, c_int32_p, C.POINTER(Prior), C.POINTER(SolverOpts), C.c_int64,
                                      C.c_int32, C.c_uint64, C.c_uint64, c_double_p, c_double_p, c_double_p, c_double_p,
                                      c_int64_p, C.POINTER(AmmTune)]
    c_uint64_p = C.POINTER(C.c_uint64)
    lib.gync_mcmc_run_ids.argtypes = [c_double_p, c_double_p, C.c_int64, c_double_p, C.c_double, c_double_p, C.c_int32,
                                      c_double_p, c_int32_p, C.POINTER(Prior), C.POINTER(SolverOpts), C.c_int64,
                                      C.c_int32, C.c_uint64, C.c_uint64, c_uint64_p, c_double_p, c_double_p, c_double_p,
                                      c_int64_p]
    lib.gync_mcmc_run_amm_ids.argtypes = [c_double_p, c_double_p, C.c_int64, c_double_p, C.c_double, c_double_p, C.c_int32,
                                          c_double_p, c_int32_p, C.POINTER(Prior), C.POINTER(SolverOpts), C.c_int64,
                                          C.c_int32, C.c_uint64, C.c_uint64, c_uint64_p, c_double_p, c_double_p, c_double_p,
                                          c_double_p, c_int64_p, C.POINTER(AmmTune)]
    lib.gync_mcmc_last_stats.argtypes = [c_uint64_p, c_uint64_p, c_int32_p]
    lib.gync_mcmc_last_stats.restype = None
    lib.gync_mcmc_last_cancelled.restype = C.c_uint64
    lib.gync_mcmc_last_early.restype = C.c_uint64
    lib.gync_mcmc_last_histogram.argtypes = [c_uint64_p, c_uint64_p]
    lib.gync_mcmc_last_histogram.restype = None
    lib.gync_em_iterate.argtypes = [c_double_p, c_double_p, C.c_int64, C.c_int32, C.c_int32, c_double_p]
    lib.gync_em_create.argtypes = [c_double_p, C.c_int64, C.c_int32, C.POINTER(C.c_void_p)]
    lib.gync_em_iterate_handle.argtypes = [C.c_void_p, c_double_p, C.c_int32, c_double_p]
    lib.gync_em_destroy.argtypes = [C.c_void_p]
    lib.gync_em_destroy.restype = None
    lib.gync_em_iterate_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.gync_em_norms_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
    lib.gync_em_step_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gync_wc_normalize.argtypes = [c_double_p, c_double_p, c_int64_p, C.c_int64, C.c_int32, c_double_p,
                                      c_double_p, c_double_p]
    lib.gync_fp64_peak.argtypes = [c_double_p]
    lib.gync_wc_normalize_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gync_wc_colstats_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
    lib.gync_wc_expsum_dev.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gync_wc_rows_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
    lib.gync_wc_scale_dev.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.gync_collapse_duplicates.argtypes = [c_double_p, C.c_int64, c_double_p, c_int64_p, c_int64_p]
    lib.gync_flops_per_step.restype = C.c_double
    lib.gync_em_comm_create.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.gync_em_comm_connect.argtypes = [C.c_void_p, C.c_void_p]
    lib.gync_em_comm_destroy.argtypes = [C.c_void_p]
    lib.gync_em_comm_destroy.restype = None
    lib.gync_em_iterate_fused_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
    lib.gync_em_comm_error.argtypes = [C.c_void_p]
    lib.gync_likelihoodmat.argtypes = [c_double_p, C.c_int64, c_double_p, C.c_int64, C.c_int32, c_double_p, C.c_int32, c_double_p]
    lib.gync_likelihoodmat_dev.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                           C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.gync_phi_batch.argtypes = [c_double_p, C.c_int64, C.POINTER(SolverOpts), c_double_p, c_int32_p]
    lib.gync_projectsimplex.argtypes = [c_double_p, C.c_int64]
    lib.gync_ebmodel_create.argtypes = [c_double_p, C.c_int64, c_double_p, C.c_int64, c_double_p, C.c_int64, C.c_int32,
                                        c_double_p, c_double_p, C.POINTER(C.c_void_p)]
    lib.gync_ebmodel_from_matrices.argtypes = [c_double_p, C.c_int64, C.c_int64, c_double_p, C.c_int64, C.POINTER(C.c_void_p)]
    lib.gync_ebmodel_destroy.argtypes = [C.c_void_p]
    lib.gync_ebmodel_destroy.restype = None
    lib.gync_ebmodel_dims.argtypes = [C.c_void_p, c_int64_p, c_int64_p, c_int64_p]
    lib.gync_ebmodel_matrices.argtypes = [C.c_void_p, c_double_p, c_double_p]
    lib.gync_ebmodel_device_ptrs.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.gync_ebmodel_logl.argtypes = [C.c_void_p, c_double_p, c_double_p]
    lib.gync_ebmodel_dlogl.argtypes = [C.c_void_p, c_double_p, c_double_p]
    lib.gync_ebmodel_hz.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p]
    lib.gync_ebmodel_dhz.argtypes = [C.c_void_p, c_double_p, C.c_int32, c_double_p]
    lib.gync_ebmodel_em.argtypes = [C.c_void_p, c_double_p, C.c_int32, c_double_p]
    lib.gync_ebmodel_mple.argtypes = [C.c_void_p, c_double_p, C.c_double, C.c_double, C.c_int32, C.c_int32, c_double_p]
    lib.gync_ebmodel_dhz_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.gync_ebmodel_dlogl_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gync_ebmodel_hz_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gync_ebmodel_ascent_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_double, C.c_double,
                                            C.c_void_p]
    lib.gync_ebmodel_nan_flag.argtypes = [C.c_void_p, C.c_void_p]
    lib.gync_ebmodel_create_zshard.argtypes = [c_double_p, C.c_int64, c_double_p, C.c_int64, c_double_p, C.c_int64, C.c_int64,
                                               C.c_int64, C.c_int32, c_double_p, c_double_p, c_double_p, C.POINTER(C.c_void_p)]
    lib.gync_ebmodel_zshard_finish.argtypes = [C.c_void_p, c_double_p]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise GyncError(f"gync_b200 error {rc}: {load().gync_last_error().decode()}")


def dptr(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def fcol(a, dtype=np.float64):
    """Column-major copy/view (Julia layout)."""
    return np.asfortranarray(a, dtype=dtype)


def opts(rtol=None, atol=None, h0=0.0, hmax=0.0, max_steps=0):
    return SolverOpts(rtol or 0.0, atol or 0.0, h0, hmax, max_steps, 0)
