#!/usr/bin/env python
"""The path a Julia host would drive: ONE process, all GPUs of the box through gync_init_multi, PAGEABLE host arrays
(a Julia Array is pageable) straight into the blocking C-ABI calls.  Prints one JSON object.
    python tools/julia_shaped_bench.py            (on a box with >= 1 GPU; use gpurun --gpus N)
  loglik : gync_loglik_batch, 1e5 near-set samples per device, 1 patient; repeated calls reuse the staging pool
  em     : gync_em_iterate, K = 1e6 x 40, 100 iterations (host arrays in, weights out; L is uploaded every call)
  mcmc   : gync_mcmc_run_ids, 512 chains per device x 600 iterations, chains sharded over the devices by the library;
           the sharded chains are compared bit for bit with a single-device run of the same chains"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402  (device count only)
import gync_b200 as g  # noqa: E402
from gync_b200 import lib as L  # noqa: E402

lib = L.load()
nd = torch.cuda.device_count()
out = {"devices": nd, "host_buffers": "pageable numpy arrays (no pinning)", "process": "single"}
rng = np.random.Generator(np.random.PCG64(5))
x0 = np.concatenate([g.api.refparms, g.api.refy0, [28.0]])
data = g.Lausanne(1).data
L.check(lib.gync_init_multi(nd, None))

# ---- likelihood batch
N = 100_000 * nd
X = np.asfortranarray(np.exp(np.log(x0)[None, :] + 0.0998 * rng.standard_normal((N, 116))))
LL = np.empty((N, 1), order="F")
st = np.zeros(N, dtype=np.int32)
Dh = np.asfortranarray(data)
sig = np.array([0.1])
op = L.opts(g.api.DEFAULT_RTOL, g.api.DEFAULT_ATOL)
ts = []
for _ in range(4):
    t0 = time.perf_counter()
    L.check(lib.gync_loglik_batch(L.dptr(X), N, L.dptr(Dh), 1, L.dptr(sig), C.byref(op), L.dptr(LL), None,
                                  st.ctypes.data_as(L.c_int32_p), None))
    ts.append(time.perf_counter() - t0)
out["loglik"] = {"samples": N, "seconds_each_call": ts, "evals_per_s_best": N / min(ts), "evals_per_s_median": N / float(np.median(ts)),
                 "failed": int((st != 0).sum()), "h2d_bytes": int(X.nbytes), "note": "first call allocates the staging pool"}

# ---- EM
K, M, niter = 1_000_000, 40, 100
Lm = np.asfortranarray(rng.random((K, M)))
Lm /= Lm.sum(0, keepdims=True)
w = np.full(K, 1.0 / K)
ts = []
for _ in range(3):
    w[:] = 1.0 / K
    t0 = time.perf_counter()
    g.emiteration_bang(w, Lm, niter=niter)
    ts.append(time.perf_counter() - t0)
out["em"] = {"K": K, "M": M, "niter": niter, "seconds_each_call": ts, "iters_per_s_incl_upload": niter / min(ts),
             "sum_w": float(w.sum()), "h2d_bytes": int(Lm.nbytes + w.nbytes)}

# ---- Metropolis: chains sharded over the devices by the library
Cn, iters = 512 * nd, int(os.environ.get("ITERS", 600))
la = g.api.alldatas()
c0 = g.Config()
c0.prior()


def chains():
    ss = []
    for c in range(Cn):
        cf = g.Config(g.Patient(la[c % 40], f"p{c % 40}"), thin=10)
        cf.priory0 = c0.priory0
        ss.append(g.new_sampling(cf, seed=31, chain_id=c))
    return ss


g.sample_chains(chains()[:32 * nd], 10)
ss = chains()
t0 = time.perf_counter()
g.sample_chains(ss, iters)
dt = time.perf_counter() - t0
solves, steps, kk = C.c_uint64(), C.c_uint64(), C.c_int32()
lib.gync_mcmc_last_stats(C.byref(solves), C.byref(steps), C.byref(kk))
out["mcmc"] = {"chains": Cn, "iters": iters, "seconds": dt, "chain_iters_per_s": Cn * iters / dt,
               "accept_rate": float(np.mean([s.variate.naccept for s in ss])) / iters, "likelihood_solves": int(solves.value),
               "iterations_per_solve": Cn * iters / max(1, solves.value), "slots_per_round": int(kk.value)}
if nd > 1:      # the same chains on ONE device: chains must not depend on how many devices ran them
    L.check(lib.gync_init(0))
    sub = [0, 1, Cn // 2, Cn - 1]
    one = [chains()[i] for i in sub]
    g.sample_chains(one, iters)
    out["mcmc"]["sharded_equals_single_device"] = bool(all(np.array_equal(one[j].samples, ss[i].samples) for j, i in enumerate(sub)))
lib.gync_shutdown()
print(json.dumps(out))
