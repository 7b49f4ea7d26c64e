"""ctypes loader of the C oracle (oracle/gync_oracle.c) — test infrastructure only."""
import ctypes as C
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgync_oracle.so")
dp = C.POINTER(C.c_double)


class Stats(C.Structure):
    _fields_ = [("nfe", C.c_long), ("nje", C.c_long), ("nlu", C.c_long), ("nsteps", C.c_long), ("nrej", C.c_long)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            import subprocess
            subprocess.check_call(["make", "-s", "-C", _HERE])
        L = C.CDLL(_SO)
        L.orc_llh.restype = C.c_double
        L.orc_llh_from_traj.restype = C.c_double
        with open(os.path.join(_HERE, "..", "gync.jl_b200", "data", "reference_constants.json")) as f:
            ref = np.array(json.load(f)["refallparms"], dtype=np.float64)
        L.orc_set_reference(ref.ctypes.data_as(dp))
        _lib = L
    return _lib


def P(a):
    return a.ctypes.data_as(dp)


def rhs(y, p):
    f = np.empty(33)
    lib().orc_rhs(P(np.ascontiguousarray(y, dtype=np.float64)), P(np.ascontiguousarray(p, dtype=np.float64)), P(f))
    return f


def llh_batch(X, datas, sigmas, mode=1, rtol=1e-12, atol=1e-12, redundant=False, want_traj=False, want_work=False):
    """X (N,116); datas list of 31x4; -> LL (N,M) [, traj (N,62,4)] [, work (nfe,nje,nlu,nsteps)]."""
    X = np.asfortranarray(np.atleast_2d(X), dtype=np.float64)
    N = X.shape[0]
    M = len(datas)
    D = np.empty((31, 4, M), order="F")
    for m, d in enumerate(datas):
        D[:, :, m] = d
    sig = np.ascontiguousarray(np.broadcast_to(np.asarray(sigmas, dtype=np.float64), (M,)))
    LL = np.empty((N, M), order="F")
    tr = np.empty((N, 62, 4)) if want_traj else None
    work = (C.c_long * 4)()
    lib().orc_llh_batch(P(X), C.c_int64(N), P(D), M, P(sig), mode, C.c_double(rtol), C.c_double(atol),
                        int(redundant), P(LL), None if tr is None else P(tr), work)
    out = [LL]
    if want_traj:
        out.append(tr)
    if want_work:
        out.append(tuple(work))
    return out[0] if len(out) == 1 else tuple(out)


def gync(y0, p, t, mode=1, rtol=1e-12, atol=1e-12):
    t = np.ascontiguousarray(t, dtype=np.float64)
    Y = np.empty((t.size, 33))
    lib().orc_gync(P(np.ascontiguousarray(y0, dtype=np.float64)), P(np.ascontiguousarray(p, dtype=np.float64)), P(t),
                   t.size, mode, C.c_double(rtol), C.c_double(atol), P(Y))
    return Y


def emiteration(w, L, mt=False):
    w = np.array(w, dtype=np.float64)
    L = np.asfortranarray(L, dtype=np.float64)
    K, M = L.shape
    norms = np.empty(M)
    (lib().orc_emiteration_mt if mt else lib().orc_emiteration)(P(w), P(L), C.c_int64(K), M, P(norms))
    return w


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    """Pin the OpenMP thread count of the batch calls (independent of OMP_NUM_THREADS in the environment)."""
    lib().orc_set_num_threads(int(n))
    return num_threads()


def host_cores():
    """Cores this process may run on (affinity mask if the platform has one)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1
