import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from gync_b200 import lib as L, api
lib = L.load(); L.check(lib.gync_init(0))
N = 100000
X = np.asfortranarray(bench.near_set(N))
Xp = torch.from_numpy(np.ascontiguousarray(X.T)).pin_memory()
Xnp = Xp.numpy().T
data = np.asfortranarray(api.Lausanne(1).data); sig = np.array([0.1])
LLh = np.empty((N, 1), order="F"); sth = np.zeros(N, dtype=np.int32); nsh = np.zeros((N, 2), dtype=np.int32, order="F")
op = L.opts(1e-8, 1e-10)
def host_call(want_ns):
    t0 = time.perf_counter()
    L.check(lib.gync_loglik_batch(L.dptr(Xnp), N, L.dptr(data), 1, L.dptr(sig), C.byref(op), L.dptr(LLh), None,
                                  sth.ctypes.data_as(L.c_int32_p), nsh.ctypes.data_as(L.c_int32_p) if want_ns else None))
    return (time.perf_counter() - t0) * 1e3
print("host path, nsteps requested :", ["%.0f" % host_call(True) for _ in range(6)])
print("host path, nsteps NOT requested:", ["%.0f" % host_call(False) for _ in range(6)])
os.environ["GYNC_NO_SCHED"] = "1"
print("host path, no sched:", ["%.0f" % host_call(False) for _ in range(3)])
# same work with pre-allocated device buffers (torch H2D from pinned + _dev entry + D2H), to separate allocation effects
os.environ.pop("GYNC_NO_SCHED", None)
dev = torch.device("cuda", 0)
dX = torch.empty(116, N, dtype=torch.float64, device=dev)
dD = torch.from_numpy(np.ascontiguousarray(api.Lausanne(1).data.T)).to(dev); dS = torch.tensor([0.1], dtype=torch.float64, device=dev)
dLL = torch.empty(N, dtype=torch.float64, device=dev); dSt = torch.empty(N, dtype=torch.int32, device=dev)
hLL = torch.empty(N, dtype=torch.float64).pin_memory()
st = torch.cuda.current_stream().cuda_stream
def dev_call():
    t0 = time.perf_counter()
    dX.copy_(Xp, non_blocking=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(lib.gync_loglik_batch_dev(dX.data_ptr(), N, dD.data_ptr(), 1, dS.data_ptr(), C.byref(op), dLL.data_ptr(), None, dSt.data_ptr(), None, st))
    e1.record()
    hLL.copy_(dLL, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, e0.elapsed_time(e1)
r = [dev_call() for _ in range(10)]
print("pre-allocated: wall", ["%.0f" % a for a, _ in r])
print("pre-allocated: kernel events", ["%.0f" % b for _, b in r])
