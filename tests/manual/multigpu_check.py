"""Multi-GPU parity check (run under torchrun on a box with >= 2 GPUs):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/manual/multigpu_check.py
Sharded EM through (a) step kernels + NCCL allreduce and (b) the fused peer-memory kernel must both equal the
unsharded oracle to 1e-10; sharded likelihood batches must equal the single-GPU result bit for bit."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import gync_b200 as g
from gync_b200 import lib as L
from gync_b200.dist import DeviceEM, FusedEM, shard_bounds
import gync_oracle as o
L.check(L.load().gync_init(local))
rng = np.random.default_rng(21)
K, M, niter = 200_003, 40, 12
Lm = rng.random((K, M)); Lm /= Lm.sum(0, keepdims=True)
w0 = rng.random(K); w0 /= w0.sum()
ref = w0.copy()
for _ in range(niter):
    ref = o.emiteration(ref, Lm)
lo, hi = shard_bounds(K, world, rank)
dev = torch.device("cuda", local)
def shard():
    return torch.from_numpy(w0[lo:hi].copy()).to(dev), torch.from_numpy(np.ascontiguousarray(Lm[lo:hi].T)).to(dev)
w, Ll = shard()
a = DeviceEM(w, Ll); a.iterate(niter); torch.cuda.synchronize()
e1 = np.max(np.abs(a.w.cpu().numpy() - ref[lo:hi]) / ref[lo:hi])
w, Ll = shard()
f = FusedEM(w, Ll, world, rank); dist.barrier()
f.iterate(5); f.iterate(niter - 5); torch.cuda.synchronize()
assert f.error() == 0, "peer wait timed out"
e2 = np.max(np.abs(f.w.cpu().numpy() - ref[lo:hi]) / ref[lo:hi])
# all ranks must hold bit-identical denominators -> identical weights where shards are re-computed: compare sums
tot = torch.tensor([float(f.w.sum().item())], dtype=torch.float64, device=dev); dist.all_reduce(tot)
print(f"rank {rank}: NCCL-path err {e1:.2e}  fused-path err {e2:.2e}  sum(w) {tot.item():.15f}", flush=True)
assert e1 < 1e-10 and e2 < 1e-10 and abs(tot.item() - 1) < 1e-12
# timing of both paths at cfg-5 size
K5 = 1_000_000 // world
gen = torch.Generator(device=dev); gen.manual_seed(5 + rank)
Lb = torch.rand(M, K5, dtype=torch.float64, device=dev, generator=gen)
cs = Lb.sum(1); dist.all_reduce(cs); Lb /= cs[:, None]
for name, mk in (("nccl", lambda w_: DeviceEM(w_, Lb)), ("fused", lambda w_: FusedEM(w_, Lb, world, rank))):
    wb = torch.full((K5,), 1.0 / (K5 * world), dtype=torch.float64, device=dev)
    em = mk(wb); dist.barrier(); em.iterate(10); torch.cuda.synchronize(); dist.barrier()
    e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); em.iterate(100); e1_.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1_)], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"EM K=1e6 over {world} GPUs, {name}: {t.item() * 10:.2f} us/iter  ({1e5 / t.item():.0f} iters/s)", flush=True)
    if name == "fused":
        assert em.error() == 0
        em.close()
f.close()
dist.barrier(); dist.destroy_process_group()
print(f"rank {rank} ok")
