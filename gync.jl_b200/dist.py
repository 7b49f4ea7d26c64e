"""Multi-GPU host logic: one process per GPU (torchrun), `torch.distributed` for the plumbing.

What shards (plus the z-matrix of the §8f-3 regulariser):
  * likelihood batches / Metropolis chains — contiguous blocks of samples per rank, NO collective;
  * the EM update — rows of L and w are sharded, L never moves; the only exchange is ONE
    allreduce(sum) of the M per-patient denominators sum_k L[k,m] w[k] per iteration.
The drivers take the local kernels as callables, so the same sequencing is exercised on CPU
(gloo, world_size 2) in tests/test_dist_gloo.py with the oracle's arithmetic standing in for
the CUDA kernels.
"""
from __future__ import annotations

import ctypes as C


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous block partition [lo, hi) of n rows; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sample_chains_sharded(samplings, iters, advance=None, rank=None, world=None, gather=True):
    """sample!(s, iters) for MANY chains over the ranks (BASELINE configs[3]: 4096 chains over 8 GPUs; replaces the Slurm
    fan-out of batch.jl:3-22).  Chains are independent, so each rank advances a contiguous block with its own persistent
    kernel and there is NO data-path collective; the only communication is the optional gather of the finished
    chains at the end (every rank then holds all of them, e.g. to build the WeightedChain).
    Every chain gets its Philox stream id from its GLOBAL position before the split, so the chains do not depend on how
    many ranks ran them.  `advance(local_samplings, iters)` defaults to api.sample_chains (the C ABI)."""
    import torch.distributed as dist
    if world is None:
        world, rank = dist.get_world_size(), dist.get_rank()
    for i, s in enumerate(samplings):
        if s.variate.chain_id is None:
            s.variate.chain_id = i
    lo, hi = shard_bounds(len(samplings), world, rank)
    if advance is None:
        from . import api
        advance = api.sample_chains
    local = list(samplings[lo:hi])
    if local:
        advance(local, iters)
    if not gather or world == 1:
        return local
    parts = [None] * world
    dist.all_gather_object(parts, local)
    return [s for part in parts for s in part]


def em_iterate_sharded(niter, norms_fn, step_fn, allreduce):
    """EM on row shards.  norms_fn() -> local partial denominators (M);  step_fn(global_norms) ->
    local partial denominators of the NEXT iteration after updating the local weights in place;
    allreduce(x) sums x over ranks in place.  One collective per iteration (+ one to start)."""
    norms = norms_fn()
    allreduce(norms)
    for _ in range(niter):
        nxt = step_fn(norms)
        allreduce(nxt)
        norms = nxt
    return norms


def wc_normalize_sharded(colstats_fn, expsum_fn, rows_fn, scale_fn, allreduce_max, allreduce_sum):
    """WeightedChain normalisation (weightedchain.jl:59-67) on row shards.  The local phases are callables:
      colstats_fn() -> stats (M+2): column maxima of LL, max log-prior, count total (local)
      expsum_fn(stats) -> colsum (M): L = exp(LL - colmax) in place, local column sums
      rows_fn(stats)  -> total (1):  stats = [colsum | lpmax | sumcounts]; local weights, unnormalised density, its total
      scale_fn(total):               density /= total
    Collectives: one max (M+1 values), two sums (1 and M values), one sum (1 value) — each once."""
    stats = colstats_fn()
    M = stats.shape[0] - 2
    allreduce_max(stats[:M + 1])
    allreduce_sum(stats[M + 1:])
    colsum = expsum_fn(stats)
    allreduce_sum(colsum)
    stats[:M] = colsum
    total = rows_fn(stats)
    allreduce_sum(total)
    scale_fn(total)
    return stats


def mple_sharded(niter, reg, dhz_partial_fn, dlogl_fn, step_fn, allreduce_sum):
    """Projected gradient ascent on reg*hz + (1-reg)*logl (likelihoodmodel.jl:33-64, gradientascent.jl:1-5) with the z-matrix
    of hz row-sharded over the ranks (it is the memory hog: Z x K doubles) and everything else replicated:
      dhz_partial_fn() -> this rank's partial dhz (K), summed over ranks by ONE allreduce per iteration;
      dlogl_fn() -> dlogl (K, replicated: the K x M data matrix is small);  step_fn(ga, gb): the projected step, identical
      on every rank because its inputs are.  The weights live inside the callables (device-resident)."""
    for _ in range(niter):
        ga = gb = None
        if reg != 0:
            ga = dhz_partial_fn()
            allreduce_sum(ga)
        if reg != 1:
            gb = dlogl_fn()
        step_fn(ga, gb)


class ShardedEB:
    """LikelihoodModel with the z-matrix row-sharded over the ranks (one process per GPU); kernels through the C ABI.
    ys: D x K (all samples, replicated), datas: D x M, zs_local: this rank's columns [z_off, z_off + Z_local) of the D x Z_total
    z samples.  `allreduce_min/max/sum` default to torch.distributed."""

    def __init__(self, ys, datas, zs_local, z_off, Z_total, sig_meas, sig_z=None, device=None,
                 allreduce_min=None, allreduce_max=None, allreduce_sum=None):
        import numpy as np
        import torch
        from . import lib as _L
        self.torch, self._L, self.lib, self.np = torch, _L, _L.load(), np
        if allreduce_sum is None:
            import torch.distributed as dist
            allreduce_min = lambda x: dist.all_reduce(x, op=dist.ReduceOp.MIN)
            allreduce_max = lambda x: dist.all_reduce(x, op=dist.ReduceOp.MAX)
            allreduce_sum = lambda x: dist.all_reduce(x)
        self.allreduce_sum = allreduce_sum
        ys, datas, zs_local = _L.fcol(ys), _L.fcol(datas), _L.fcol(zs_local)
        self.K, self.M = ys.shape[1], datas.shape[1]
        sm = np.ascontiguousarray(sig_meas, dtype=np.float64)
        sz = None if sig_z is None else np.ascontiguousarray(sig_z, dtype=np.float64)
        scal = np.empty(2)
        self.h = C.c_void_p()
        _L.check(self.lib.gync_ebmodel_create_zshard(_L.dptr(ys), self.K, _L.dptr(datas), self.M, _L.dptr(zs_local), zs_local.shape[1],
                                                     int(z_off), int(Z_total), ys.shape[0], _L.dptr(sm), _L.dptr(sz), _L.dptr(scal),
                                                     C.byref(self.h)))
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        t = torch.from_numpy(scal).to(dev)
        allreduce_min(t[0:1])                        # global min of the squared distances (likelihoodmat.jl:53)
        allreduce_max(t[1:2])                        # global max of the normalisation (likelihoodmat.jl:54)
        scal = np.ascontiguousarray(t.cpu().numpy())
        _L.check(self.lib.gync_ebmodel_zshard_finish(self.h, _L.dptr(scal)))
        self.w = torch.full((self.K,), 1.0 / self.K, dtype=torch.float64, device=dev)
        self.ga, self.gb = torch.empty_like(self.w), torch.empty_like(self.w)
        self.hzv = torch.empty(1, dtype=torch.float64, device=dev)

    def _s(self):
        return self.torch.cuda.current_stream().cuda_stream

    def dhz_partial(self, cont=False):
        self._L.check(self.lib.gync_ebmodel_dhz_dev(self.h, self.w.data_ptr(), 1 if cont else 0, self.ga.data_ptr(), self._s()))
        return self.ga

    def dhz(self, cont=False):
        g = self.dhz_partial(cont)
        self.allreduce_sum(g)
        return g

    def dlogl(self):
        self._L.check(self.lib.gync_ebmodel_dlogl_dev(self.h, self.w.data_ptr(), self.gb.data_ptr(), self._s()))
        return self.gb

    def hz(self):
        self._L.check(self.lib.gync_ebmodel_hz_dev(self.h, self.w.data_ptr(), self.hzv.data_ptr(), self._s()))
        self.allreduce_sum(self.hzv)
        return float(self.hzv.item())

    def mple(self, reg, h, nsteps, ndata=None, autoh=True):
        """nsteps projected steps on the resident weights (likelihoodmodel.jl:33-41); raises on NaN in dhz (:77)."""
        if autoh:
            h = h / ((1 - reg) * ((self.M if ndata is None else ndata) / (1 / 100)) + reg / (1 / 1000))
        ca, cb = (1.0 if reg == 1 else reg), (1.0 if reg == 0 else 1.0 - reg)

        def step(ga, gb):
            self._L.check(self.lib.gync_ebmodel_ascent_dev(self.h, self.w.data_ptr(), None if ga is None else ga.data_ptr(), ca,
                                                           None if gb is None else gb.data_ptr(), cb, float(h), self._s()))
        mple_sharded(nsteps, reg, self.dhz_partial, self.dlogl, step, self.allreduce_sum)
        if self.lib.gync_ebmodel_nan_flag(self.h, self._s()) == 1:
            raise FloatingPointError("encountered NaN in dhz")
        return self.w

    def close(self):
        if self.h:
            self.torch.cuda.synchronize()
            self.lib.gync_ebmodel_destroy(self.h)
            self.h = None


class DeviceWC:
    """Row shard of (LL, logprior, counts) on this rank's GPU for wc_normalize_sharded; kernels through the C ABI."""

    def __init__(self, LL_local_colmajor, logprior_local, counts_local):
        import torch
        from . import lib as _L
        self.torch, self._L, self.lib = torch, _L, _L.load()
        self.LL, self.lp, self.cn = LL_local_colmajor, logprior_local, counts_local    # (M, K_local) row-major == K_local x M col-major
        self.M, self.K = self.LL.shape
        dev = self.LL.device
        self.stats = torch.empty(self.M + 2, dtype=torch.float64, device=dev)
        self.colsum = torch.empty(self.M, dtype=torch.float64, device=dev)
        self.total = torch.empty(1, dtype=torch.float64, device=dev)
        self.weights = torch.empty(self.K, dtype=torch.float64, device=dev)
        self.density = torch.empty(self.K, dtype=torch.float64, device=dev)

    def _s(self):
        return self.torch.cuda.current_stream().cuda_stream

    def colstats(self):
        self._L.check(self.lib.gync_wc_colstats_dev(self.LL.data_ptr(), self.lp.data_ptr(), self.cn.data_ptr(), self.K, self.M,
                                                    self.stats.data_ptr(), self._s()))
        return self.stats

    def expsum(self, stats):
        self._L.check(self.lib.gync_wc_expsum_dev(self.LL.data_ptr(), self.K, self.M, stats.data_ptr(), self.colsum.data_ptr(), self._s()))
        return self.colsum

    def rows(self, stats):
        self._L.check(self.lib.gync_wc_rows_dev(self.LL.data_ptr(), self.lp.data_ptr(), self.cn.data_ptr(), self.K, self.M,
                                                stats.data_ptr(), self.weights.data_ptr(), self.density.data_ptr(),
                                                self.total.data_ptr(), self._s()))
        return self.total

    def scale(self, total):
        self._L.check(self.lib.gync_wc_scale_dev(self.density.data_ptr(), self.K, total.data_ptr(), self._s()))

    def normalize(self, allreduce_max=None, allreduce_sum=None):
        """LL -> L in place; returns (weights, density) of the local rows."""
        if allreduce_max is None:
            import torch.distributed as dist
            allreduce_max = lambda x: dist.all_reduce(x, op=dist.ReduceOp.MAX)
            allreduce_sum = lambda x: dist.all_reduce(x)
        wc_normalize_sharded(self.colstats, self.expsum, self.rows, self.scale, allreduce_max, allreduce_sum)
        return self.weights, self.density


class DeviceEM:
    """Row shard of (w, L) resident on this rank's GPU; kernels through the C ABI (device pointers)."""

    def __init__(self, w_local, L_local_colmajor):
        import torch
        from . import lib as _L
        self.torch, self._L, self.lib = torch, _L, _L.load()
        self.w = w_local                      # torch float64 (K_local,)
        self.L = L_local_colmajor             # torch float64 (M, K_local) row-major == K_local x M column-major
        self.M, self.K = self.L.shape
        self.bufs = [torch.empty(self.M, dtype=torch.float64, device=w_local.device) for _ in range(2)]
        self.flip = 0

    def _stream(self):
        return self.torch.cuda.current_stream().cuda_stream

    def norms(self):
        out = self.bufs[self.flip]
        self._L.check(self.lib.gync_em_norms_dev(self.w.data_ptr(), self.L.data_ptr(), self.K, self.M, out.data_ptr(),
                                                 self._stream()))
        return out

    def step(self, norms):
        self.flip ^= 1
        out = self.bufs[self.flip]
        self._L.check(self.lib.gync_em_step_dev(self.w.data_ptr(), self.L.data_ptr(), self.K, self.M, norms.data_ptr(),
                                                out.data_ptr(), self._stream()))
        return out

    def iterate(self, niter):
        import torch.distributed as dist
        return em_iterate_sharded(niter, self.norms, self.step, lambda x: dist.all_reduce(x))


class FusedEM:
    """Row shard of (w, L) on this rank's GPU, iterated by the FUSED kernel: EM step + allreduce of the M
    denominators in one persistent cooperative launch per rank, partial sums exchanged through
    cudaIpc-mapped peer mailboxes over NVLink (csrc/gync_kernels.cuh: gync_em_kernel with peers).
    `torch.distributed` is used once, to exchange the 64-byte IPC handles."""

    def __init__(self, w_local, L_local_colmajor, world=1, rank=0, all_gather_bytes=None):
        import torch
        from . import lib as _L
        self.torch, self._L, self.lib = torch, _L, _L.load()
        self.w, self.L = w_local, L_local_colmajor
        self.M, self.K = self.L.shape
        self.world, self.rank = world, rank
        handle = (C.c_char * 64)()
        self.comm = C.c_void_p()
        _L.check(self.lib.gync_em_comm_create(world, rank, handle, C.byref(self.comm)))
        if world > 1:
            if all_gather_bytes is None:
                import torch.distributed as dist

                def all_gather_bytes(b):
                    t = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(w_local.device)
                    out = [torch.empty_like(t) for _ in range(world)]
                    dist.all_gather(out, t)
                    return b"".join(bytes(o.cpu().numpy().tobytes()) for o in out)
            allh = all_gather_bytes(bytes(handle))
            assert len(allh) == 64 * world
            self._handles = C.create_string_buffer(allh, 64 * world)
            _L.check(self.lib.gync_em_comm_connect(self.comm, self._handles))

    def iterate(self, niter):
        """All ranks must call this with the same niter."""
        self._L.check(self.lib.gync_em_iterate_fused_dev(self.comm, self.w.data_ptr(), self.L.data_ptr(), self.K, self.M,
                                                         niter, self.torch.cuda.current_stream().cuda_stream))

    def error(self):
        return self.lib.gync_em_comm_error(self.comm)

    def close(self):
        if self.comm:
            self.torch.cuda.synchronize()
            self.lib.gync_em_comm_destroy(self.comm)
            self.comm = None
