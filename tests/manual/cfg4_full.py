"""BASELINE configs[3] in full: 4096 Metropolis chains x 10^4 iterations on synthetic patient data (chain c on synthetic
patient c mod 40, reference start, default proposal, thin 10), sharded over the ranks by dist.sample_chains_sharded.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tests/manual/cfg4_full.py
Rank 0 prints one JSON object (wall clock around the blocking call, max over ranks)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import gync_b200 as g
from gync_b200 import lib as L
from gync_b200.dist import sample_chains_sharded
import bench
L.check(L.load().gync_init(local))
NC, IT = int(os.environ.get("NC", 4096)), int(os.environ.get("IT", 10000))
Xh = bench.near_set(4096, bench.SEED)
D = bench.synthetic_patients(Xh, 40)
pats = [g.Patient(np.ascontiguousarray(D[:, :, m]), f"syn{m + 1}") for m in range(40)]
c0 = g.Config(); c0.prior()
ss = []
for c in range(NC):
    cf = g.Config(pats[c % 40], thin=10)
    cf.priory0 = c0.priory0
    ss.append(g.new_sampling(cf, seed=4))
sample_chains_sharded([s for s in ss[:8 * world]], 20, gather=False)          # warm-up (first call: kernel attributes)
for s in ss:
    s.variate.chain_id = None
ss = [g.new_sampling(s.config, seed=4) for s in ss]
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
mine = sample_chains_sharded(ss, IT, gather=False)
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
dist.all_reduce(dt, op=dist.ReduceOp.MAX)
acc = torch.tensor([float(sum(s.variate.naccept for s in mine)), float(len(mine))], dtype=torch.float64, device="cuda")
dist.all_reduce(acc)
ok = all(s.samples.shape == (IT // 10, 116) and np.all(np.isfinite(s.samples)) for s in mine)
if rank == 0:
    print(json.dumps({"config": "4096 chains x 1e4 iterations, thin 10, synthetic patients, sharded over %d GPUs" % world, "chains": NC, "iters": IT,
                      "seconds": float(dt.item()), "chain_iters_per_s": NC * IT / float(dt.item()),
                      "accept_rate": float(acc[0].item()) / (float(acc[1].item()) * IT), "all_rows_finite": bool(ok),
                      "chains_per_rank": len(mine)}))
dist.destroy_process_group()
