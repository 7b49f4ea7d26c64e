"""A/B of the learned longest-first sample ordering (csrc/gync_sched.cuh) in one process: same inputs, results must be
bit-identical, only the kernel time changes."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from gync_b200 import lib as L, api
lib = L.load(); L.check(lib.gync_init(0))
N = int(os.environ.get("N", 100000))
X = bench.near_set(N)
dX = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()
dD = torch.from_numpy(np.ascontiguousarray(api.Lausanne(1).data.T)).cuda()
dS = torch.tensor([0.1], dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
op = L.opts(7e-10, 1e-11)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
outs = {}
for name, env in (("sched", None), ("no-sched", "1"), ("sched again", None)):
    if env: os.environ["GYNC_NO_SCHED"] = env
    else: os.environ.pop("GYNC_NO_SCHED", None)
    LL = torch.empty(N, dtype=torch.float64, device="cuda")
    ns = torch.empty(2 * N, dtype=torch.int32, device="cuda")
    ts = []
    for r in range(5):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.gync_loglik_batch_dev(dX.data_ptr(), N, dD.data_ptr(), 1, dS.data_ptr(), C.byref(op), LL.data_ptr(), None, None, ns.data_ptr(), st))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    outs[name] = (LL.clone(), ns.clone())
    print(f"{name:12s} ms per batch: " + ", ".join(f"{t:.1f}" for t in ts) + f"   best {min(ts):.1f}  ({N / min(ts) * 1e3:.0f} evals/s)", flush=True)
a, b = outs["sched"], outs["no-sched"]
print("bit-identical log-likelihoods:", bool(torch.equal(a[0], b[0])), " identical step counts:", bool(torch.equal(a[1], b[1])))
