#!/usr/bin/env python
"""Per-phase SM-clock breakdown of the row-sharded EM iteration (fused compute + exchange kernel) under torchrun, from a
-DGYNC_EM_PHASES build of the library (GYNC_B200_LIB points at it):
   nvcc ... -DGYNC_EM_PHASES gync_b200.cu -o tools/ab/lib_em_phases.so
   GYNC_B200_LIB=$PWD/tools/ab/lib_em_phases.so python -m torch.distributed.run --nproc-per-node N tools/em_phase_breakdown.py
K = 1e6 x 40 rows sharded over the ranks, 200 iterations; rank 0 prints one JSON object."""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from gync_b200 import lib as L
from gync_b200.dist import FusedEM
lib = L.load(); L.check(lib.gync_init(local))
dev = torch.device("cuda", local)
K5, M, niter = 1_000_000 // world, 40, 200
gen = torch.Generator(device=dev); gen.manual_seed(5 + rank)
Lb = torch.rand(M, K5, dtype=torch.float64, device=dev, generator=gen)
cs = Lb.sum(1); dist.all_reduce(cs); Lb /= cs[:, None]
wb = torch.full((K5,), 1.0 / (K5 * world), dtype=torch.float64, device=dev)
em = FusedEM(wb, Lb, world, rank); dist.barrier(); em.iterate(10); torch.cuda.synchronize(); dist.barrier()
out = (C.c_uint64 * 8)()
lib.gync_em_phase_read.argtypes = [C.POINTER(C.c_uint64)]
L.check(lib.gync_em_phase_read(out))            # clear
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); em.iterate(niter); e1.record(); torch.cuda.synchronize()
L.check(lib.gync_em_phase_read(out))
us = e0.elapsed_time(e1) * 1e3 / niter
mhz = torch.cuda.clock_rate(local) if hasattr(torch.cuda, "clock_rate") else 1965
nbi, ns = max(1, out[5]), max(1, out[6])
res = {"world": world, "K_per_gpu": K5, "M": M, "us_per_iter": us, "sm_clock_mhz": mhz,
       "per_block_iteration_us": {"local_pass": out[0] / nbi / mhz, "arrival_to_all_contributions_read": out[1] / nbi / mhz,
                                  "rank_order_sum": out[4] / nbi / mhz},
       "last_arriving_block_us": {"gather_block_partials": out[2] / ns / mhz, "mailbox_stores": out[3] / ns / mhz},
       "nvlink_bytes_per_iteration_per_rank": (world - 1) * M * 16,
       "note": "ticks summed over blocks and iterations / counts; the wait of a block includes waiting for the slowest block of "
               "its own GPU, the gather and the NVLink flight of the slowest rank"}
assert em.error() == 0
t = torch.tensor([us], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    res["us_per_iter_max_over_ranks"] = float(t.item())
    print(json.dumps(res))
em.close()
dist.destroy_process_group()
