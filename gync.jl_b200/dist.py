"""Multi-GPU host logic: one process per GPU (torchrun), `torch.distributed` for the plumbing.

What shards:
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
