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
