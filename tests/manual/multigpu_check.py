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
# ---- z-sharded regulariser (dist.ShardedEB) and row-sharded WeightedChain normalisation (dist.DeviceWC) vs the oracle
import eb_oracle as eo
from gync_b200.dist import ShardedEB, DeviceWC
rs = np.random.default_rng(31)                      # same stream on every rank: replicated inputs
Ke, Me, zm = 600, 11, 2
base = np.array([40.0, 5.0, 150.0, 6.0]); sg = np.tile(np.array([12.0, 1.0, 40.0, 1.5]), (31, 1)) * 0.3
ys = [base * (1 + 0.03 * rs.standard_normal((31, 4))) for _ in range(Ke)]
dts = []
for _ in range(Me):
    d_ = base * (1 + 0.05 * rs.standard_normal((31, 4))); d_[rs.random((31, 4)) < 0.4] = np.nan; dts.append(d_)
zs = [y + rs.standard_normal((31, 4)) * sg for _ in range(zm) for y in ys]
col = lambda ms: np.asfortranarray(np.stack([m_.ravel(order="F") for m_ in ms], axis=1))
zlo, zhi = shard_bounds(Ke * zm, world, rank)
sh = ShardedEB(col(ys), col(dts), col(zs)[:, zlo:zhi], zlo, Ke * zm, sg.ravel(order="F"), device=dev)
Ld_o, Lz_o = eo.likelihoodmat_nanfast(ys, dts, sg), eo.likelihoodmat_nanfast(zs, ys, sg)
w_e = np.full(Ke, 1.0 / Ke)
e_dhz = float(np.max(np.abs(sh.dhz().cpu().numpy() / eo.dhzdiscr(w_e, Lz_o) - 1)))
e_hz = abs(sh.hz() / eo.hz(w_e, Lz_o, eo.repeatweights(w_e, Ke * zm)) - 1)
n_e, reg_e = 8, 0.5
w_dev = sh.mple(reg_e, 1.0, n_e - 1).cpu().numpy()
ref_e = eo.gradientascent(eo.dmple_obj(Ld_o.T, Lz_o, reg_e), w_e, n_e, eo.mple_stepsize(1.0, reg_e, Me))[-1]
e_mple = float(np.max(np.abs(w_dev - ref_e)))
print(f"rank {rank}: z-sharded dhz err {e_dhz:.2e}  hz err {e_hz:.2e}  mple iterate err {e_mple:.2e}", flush=True)
assert e_dhz < 1e-8 and e_hz < 1e-9 and e_mple < 1e-12          # dhz/hz inherit the oracle's binomial-formula cancellation in Lz
sh.close()
Kw = 40_003
LLw = -rs.random((Kw, M)) * 600; lpw = -rs.random(Kw) * 50; cnw = rs.integers(1, 9, Kw).astype(np.int64)
wlo, whi = shard_bounds(Kw, world, rank)
wc = DeviceWC(torch.from_numpy(np.ascontiguousarray(LLw[wlo:whi].T)).to(dev), torch.from_numpy(lpw[wlo:whi].copy()).to(dev),
              torch.from_numpy(cnw[wlo:whi].copy()).to(dev))
ww, dd = wc.normalize()
l0, w0_, d0 = o.weightedchain_normalize(lpw, LLw, cnw)
e_w = float(np.max(np.abs(ww.cpu().numpy() / w0_[wlo:whi] - 1))); e_d = float(np.max(np.abs(dd.cpu().numpy() / d0[wlo:whi] - 1)))
e_l = float(np.max(np.abs(wc.LL.cpu().numpy().T - l0[wlo:whi])))
print(f"rank {rank}: sharded WeightedChain normalisation  weights err {e_w:.2e}  density err {e_d:.2e}  L abs err {e_l:.2e}", flush=True)
assert e_w < 1e-13 and e_d < 1e-11 and e_l < 1e-15
dist.barrier(); dist.destroy_process_group()
print(f"rank {rank} ok")
