"""A/B timing of several builds of libgync_b200.so in ONE process on ONE box (boxes differ by ~10 %).
   GYNC_LIBS=a.so,b.so,...  [GYNC_TOLS=rtol:atol,rtol:atol,...]  N=32768 REPS=3 python tools/ab_bench.py"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gync_b200 import lib as L, api
libs = os.environ["GYNC_LIBS"].split(",")
N = int(os.environ.get("N", 32768)); reps = int(os.environ.get("REPS", 3))
rtol = float(os.environ.get("RTOL", 1e-8)); atol = float(os.environ.get("ATOL", 1e-10))
rng = np.random.Generator(np.random.PCG64(20160913))
x0 = np.concatenate([api.refparms, api.refy0, [28.0]])
X = np.exp(np.log(x0)[None, :] + 0.0998 * rng.standard_normal((N, 116)))
data = api.Lausanne(1).data
dX = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()
dD = torch.from_numpy(np.ascontiguousarray(data.T)).cuda()
dS = torch.tensor([0.1], dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
tols = [tuple(float(v) for v in t.split(":")) for t in os.environ["GYNC_TOLS"].split(",")] if os.environ.get("GYNC_TOLS") else [(rtol, atol)] * len(libs)
ops = [L.opts(rt, at) for rt, at in tols]
handles = []
for p in libs:
    h = C.CDLL(os.path.join(ROOT, p) if not os.path.isabs(p) else p)
    h.gync_loglik_batch_dev.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    assert h.gync_init(0) == 0
    handles.append(h)
outs = [torch.empty(N, dtype=torch.float64, device="cuda") for _ in libs]
res = {p: [] for p in libs}
for r in range(reps + 1):
    for p, h, o, op in zip(libs, handles, outs, ops):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = h.gync_loglik_batch_dev(dX.data_ptr(), N, dD.data_ptr(), 1, dS.data_ptr(), C.byref(op), o.data_ptr(), None, None, None, st)
        e1.record(); torch.cuda.synchronize()
        assert rc == 0
        if r > 0: res[p].append(e0.elapsed_time(e1))
ref = outs[0]
for p, o in zip(libs, outs):
    ms = min(res[p])
    print("%-28s best %.1f ms  %.0f solves/s   (all: %s)  max|dLL| vs first: %.2e" % (p, ms, N / ms * 1e3, ", ".join("%.1f" % v for v in res[p]), float((o - ref).abs().max())))
