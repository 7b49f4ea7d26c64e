"""Row-sharded EM (cfg 5: K = 10^6 rows x 40 over the ranks) through the fused compute+exchange kernel, with the partly
shared-memory-resident shard on and off (GYNC_EM_NO_HYBRID), and a parity check of the sharded result against the oracle.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tests/manual/em_sharded_time.py"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from gync_b200 import lib as L
from gync_b200.dist import FusedEM, shard_bounds
import gync_oracle as o
L.check(L.load().gync_init(local))
libc = C.CDLL(None)
dev = torch.device("cuda", local)
M = 40
def mode(m):
    if m == "streaming": libc.setenv(b"GYNC_EM_NO_HYBRID", b"1", 1)
    else: libc.unsetenv(b"GYNC_EM_NO_HYBRID")
# parity at the shard size where the hybrid is active (K/world = 125 000 at 8 ranks)
K, niter = 125_000 * world + 3, 8
rng = np.random.default_rng(77)
Lm = rng.random((K, M)); Lm /= Lm.sum(0, keepdims=True)
w0 = rng.random(K); w0 /= w0.sum()
ref = w0.copy()
if rank == 0:
    for _ in range(niter): ref = o.emiteration(ref, Lm)
ref_t = torch.from_numpy(ref).to(dev); dist.broadcast(ref_t, 0)
lo, hi = shard_bounds(K, world, rank)
for m in ("hybrid", "streaming"):
    mode(m)
    w = torch.from_numpy(w0[lo:hi].copy()).to(dev); Ll = torch.from_numpy(np.ascontiguousarray(Lm[lo:hi].T)).to(dev)
    f = FusedEM(w, Ll, world, rank); dist.barrier(); f.iterate(niter); torch.cuda.synchronize()
    e = torch.tensor([float(((f.w - ref_t[lo:hi]).abs() / ref_t[lo:hi]).max())], dtype=torch.float64, device=dev)
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    assert f.error() == 0 and e.item() < 1e-10, e.item()
    if rank == 0: print(f"parity {m}: max rel err vs oracle over all ranks {e.item():.2e}", flush=True)
    f.close()
K5 = 1_000_000 // world
gen = torch.Generator(device=dev); gen.manual_seed(5 + rank)
Lb = torch.rand(M, K5, dtype=torch.float64, device=dev, generator=gen)
cs = Lb.sum(1); dist.all_reduce(cs); Lb /= cs[:, None]
for rep in range(3):
    for m in ("hybrid", "streaming"):
        mode(m)
        wb = torch.full((K5,), 1.0 / (K5 * world), dtype=torch.float64, device=dev)
        em = FusedEM(wb, Lb, world, rank); dist.barrier(); em.iterate(10); torch.cuda.synchronize(); dist.barrier()
        best = 1e9
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); em.iterate(100); e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, t.item())
        if rank == 0:
            print(f"EM K=1e6 over {world} GPUs ({K5} rows per GPU), {m}: {best * 10:.2f} us/iter  ({1e5 / best:.0f} iters/s)", flush=True)
        assert em.error() == 0
        em.close()
dist.barrier(); dist.destroy_process_group()
