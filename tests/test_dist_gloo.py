"""N>1 host logic on CPU: world_size-2 gloo run of the sharded EM driver (gync.jl_b200/dist.py) with the
oracle's arithmetic standing in for the CUDA kernels; result must equal the unsharded oracle."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle"))
from gync_b200.dist import shard_bounds, em_iterate_sharded
import gync_oracle as o
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(11)
K, M, niter = 1003, 7, 25
L = rng.random((K, M)); L /= L.sum(0, keepdims=True)
w0 = rng.random(K); w0 /= w0.sum()
lo, hi = shard_bounds(K, world, rank)
Ll, wl = L[lo:hi].copy(), w0[lo:hi].copy()
def norms():
    return torch.from_numpy(Ll.T @ wl)
def step(n):
    n = n.numpy()
    s = (Ll / n[None, :]).sum(1)
    wl[:] = wl / M * s                      # em.jl:33-39 on the local rows
    return torch.from_numpy(Ll.T @ wl)
calls = [0]
def allreduce(x):
    calls[0] += 1
    dist.all_reduce(x)
em_iterate_sharded(niter, norms, step, allreduce)
assert calls[0] == niter + 1                 # ONE collective per iteration (+1 to start)
ref = w0.copy()
for _ in range(niter):
    ref = o.emiteration(ref, L)
err = np.max(np.abs(wl - ref[lo:hi]) / ref[lo:hi])
assert err < 1e-12, err
tot = torch.tensor([wl.sum()]); dist.all_reduce(tot)
assert abs(tot.item() - 1) < 1e-12
print("rank", rank, "ok", err)
dist.destroy_process_group()
'''


def test_sharded_em_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29561", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


WORKER_WC = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle"))
from gync_b200.dist import shard_bounds, wc_normalize_sharded
import gync_oracle as o
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(12)
K, M = 501, 5
LL = -50.0 * rng.random((K, M)); lp = -3.0 * rng.random(K); cn = rng.integers(1, 9, K)
lo, hi = shard_bounds(K, world, rank)
Ll, lpl, cnl = LL[lo:hi].copy(), lp[lo:hi], cn[lo:hi]
w = np.empty(hi - lo); d = np.empty(hi - lo)
def colstats():
    return torch.from_numpy(np.concatenate([Ll.max(0), [lpl.max()], [float(cnl.sum())]]))
def expsum(st):
    Ll[:] = np.exp(Ll - st.numpy()[None, :M])
    return torch.from_numpy(Ll.sum(0))
def rows(st):
    s = st.numpy()
    Ll[:] = Ll / s[None, :M]
    d[:] = Ll.mean(1) * np.exp(lpl - s[M]); w[:] = cnl / s[M + 1]
    return torch.tensor([d.sum()])
def scale(t):
    d[:] = d / t.item()
ncoll = [0]
def amax(x): ncoll[0] += 1; dist.all_reduce(x, op=dist.ReduceOp.MAX)
def asum(x): ncoll[0] += 1; dist.all_reduce(x)
wc_normalize_sharded(colstats, expsum, rows, scale, amax, asum)
assert ncoll[0] == 4
Lr, wr, dr = o.weightedchain_normalize(lp, LL, cn)
assert np.allclose(Ll, Lr[lo:hi], rtol=1e-13) and np.allclose(w, wr[lo:hi], rtol=1e-13) and np.allclose(d, dr[lo:hi], rtol=1e-12)
print("rank", rank, "ok")
dist.destroy_process_group()
'''


def test_sharded_wc_normalize_gloo_world2(tmp_path):
    """row-sharded WeightedChain normalisation — allreduce(max), allreduce(sum) — equals the unsharded oracle."""
    script = tmp_path / "worker_wc.py"
    script.write_text(WORKER_WC.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29563", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


WORKER_EB = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle"))
from gync_b200.dist import shard_bounds, mple_sharded
import eb_oracle as eo
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(13)
K, M, zm, n = 60, 5, 2, 8
Ld = rng.random((K, M)) + 0.05
Lz = rng.random((K * zm, K)) + 0.05
lo, hi = shard_bounds(K * zm, world, rank)
Lzl = Lz[lo:hi]
w = np.full(K, 1.0 / K)
reg = 0.5
h = eo.mple_stepsize(1.0, reg, M)
ncoll = [0]
def dhz_partial():
    rho = Lzl @ w                                        # local rows only
    wz = np.tile(w, zm)[lo:hi] / zm
    d = -(Lzl.T @ (wz / rho))
    for z in range(lo, hi):                              # the log term of the rows held here
        d[z % K] -= np.log(rho[z - lo]) / zm
    return torch.from_numpy(d)
def dlogl():
    return torch.from_numpy(eo.dlogl(w, Ld.T))
def step(ga, gb):
    g = reg * ga.numpy() + (1 - reg) * gb.numpy()
    w[:] = eo.movefromboundary(w, eo.projectsimplex(w + h * g))
def asum(x):
    ncoll[0] += 1
    dist.all_reduce(x)
mple_sharded(n - 1, reg, dhz_partial, dlogl, step, asum)
assert ncoll[0] == n - 1                                  # ONE collective per iteration
ref = eo.gradientascent(eo.dmple_obj(Ld.T, Lz, reg), np.full(K, 1.0 / K), n, h)[-1]
assert np.max(np.abs(w - ref)) < 1e-14, np.max(np.abs(w - ref))
print("rank", rank, "ok")
dist.destroy_process_group()
'''


def test_sharded_mple_gloo_world2(tmp_path):
    """z-sharded projected gradient ascent (dist.mple_sharded): one allreduce(sum) of the K partial gradients per iteration
    reproduces the unsharded oracle iterates."""
    script = tmp_path / "worker_eb.py"
    script.write_text(WORKER_EB.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29565", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


WORKER_CHAINS = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle"))
from gync_b200.dist import shard_bounds, sample_chains_sharded
from gync_b200 import api
import gync_oracle as o
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
g = np.load(os.path.join({root!r}, "tests", "golden", "solution_golden.npz"))
means, stds = g["gmm_means"], g["gmm_stds"]

def advance(ss, iters):
    """Stand-in for the device sampler: the oracle's Metropolis step on a cheap surrogate of the target (log-prior +
    Jacobian term, no solve), randomness addressed by (seed, chain_id, iteration) like the device's Philox streams."""
    for s in ss:
        v = s.variate
        logf = lambda xl: o.logprior(np.exp(xl), means, stds) + xl.sum()
        cur = logf(v.state) if np.isnan(v.logf) else v.logf
        rows = []
        for it in range(v.iteration, v.iteration + iters):
            r = np.random.Generator(np.random.Philox(key=[v.seed, v.chain_id], counter=it))
            v.state, cur, a = o.metropolis_step(v.state, cur, v.SigmaL, r.standard_normal(116), r.random(), logf)
            v.naccept += int(a)
            rows.append(np.exp(v.state))
        v.logf, v.iteration = cur, v.iteration + iters
        s.samples = np.vstack([s.samples, np.array(rows)])

mk = lambda: [api.new_sampling(api.Config(propvar=api.uniformpropvar(0.03)), seed=77) for _ in range(7)]   # 2 ranks: blocks of 4 and 3
allc = sample_chains_sharded(mk(), 15, advance=advance)
assert len(allc) == 7 and [s.variate.chain_id for s in allc] == list(range(7))
lo, hi = shard_bounds(7, world, rank)
mine = sample_chains_sharded(mk(), 15, advance=advance, gather=False)
assert len(mine) == hi - lo
ref = mk()
for i, s in enumerate(ref):
    s.variate.chain_id = i
advance(ref, 15)                                                              # the unsharded run
for a, b in zip(allc, ref):
    assert np.array_equal(a.samples, b.samples) and a.variate.naccept == b.variate.naccept
moved = [s for s in allc if s.variate.naccept > 0]
assert len(moved) >= 2 and len({{s.samples.tobytes() for s in moved}}) == len(moved)     # same seed, distinct streams
print("rank", rank, "ok")
dist.destroy_process_group()
'''


def test_sharded_chains_gloo_world2(tmp_path):
    """Chain sharding: contiguous blocks, no data-path collective, results independent of the world
    size because every chain carries its own stream id."""
    script = tmp_path / "worker_chains.py"
    script.write_text(WORKER_CHAINS.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29564", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2
