"""GPU exploration script (not part of the product): correctness spot checks + timings of tuning variants."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import gync_b200 as g
from gync_b200 import lib as L
import gync_oracle as o

lib = L.load()
L.check(lib.gync_init(0))
print(torch.cuda.get_device_name(0))
tf = C.c_double(); L.check(lib.gync_fp64_peak(C.byref(tf))); print("fp64 FMA peak TFLOP/s: %.2f" % tf.value)

N = int(os.environ.get("N", 16384))
rng = np.random.default_rng(20160911 + 2)
x0 = o.init_x()
X = np.exp(np.log(x0)[None, :] + 0.0998 * rng.standard_normal((N, 116)))
X[0] = x0
c = g.Config()
rtol = float(os.environ.get("RTOL", 1e-8)); atol = float(os.environ.get("ATOL", 1e-8))

# correctness vs host shim (same algorithm compiled for the CPU) on a few samples
hs = C.CDLL(os.path.join(ROOT, "tests", "_build", "host_shim.so"))
LL, st, Ym, ns = g.loglik_matrix(X[:256], [c.data], [0.1], rtol, atol, want_traj=True, want_steps=True)
print("llh(ref) = %.12f  (golden -247.759224067864) status0=%d steps=%s" % (LL[0, 0], st[0], ns[0]))
dp = C.POINTER(C.c_double)
mx = 0
for i in range(8):
    tr = np.empty((62, 4)); na = C.c_int(); nr = C.c_int()
    s = hs.hs_meas_traj(X[i].ctypes.data_as(dp), C.c_double(rtol), C.c_double(atol), 200000, tr.ctypes.data_as(dp), C.byref(na), C.byref(nr))
    mx = max(mx, np.nanmax(np.abs(tr - Ym[i]) / (np.abs(tr) + 1e-300)))
    print(i, s, st[i], na.value, ns[i, 0], "llh gpu %.10f host %.10f" % (LL[i, 0], o.llh_from_traj(tr, c.data, 0.1)))
print("max rel traj diff gpu vs host build:", mx, " failures:", int((st != 0).sum()), " steps mean/max:", ns[:, 0].mean(), ns[:, 0].max())

# timing of variants (device-resident inputs, CUDA events)
dX = torch.from_numpy(np.asfortranarray(X).T.copy()).cuda()        # (116,N) row-major == (N,116) col-major
dD = torch.from_numpy(np.asfortranarray(c.data).T.copy()).cuda()
dS = torch.tensor([0.1], dtype=torch.float64, device="cuda")
dLL = torch.empty(N, dtype=torch.float64, device="cuda")
dSt = torch.empty(N, dtype=torch.int32, device="cuda")
dNs = torch.empty(2 * N, dtype=torch.int32, device="cuda")
op = L.opts(rtol, atol)
stream = torch.cuda.current_stream().cuda_stream
for var in [int(v) for v in os.environ.get("VARIANTS", "0,1,2,3,4,5,6").split(",")]:
    os.environ["GYNC_LLK_VARIANT"] = str(var)
    L.check(lib.gync_init(0))
    def run():
        L.check(lib.gync_loglik_batch_dev(dX.data_ptr(), N, dD.data_ptr(), 1, dS.data_ptr(), C.byref(op), dLL.data_ptr(), None, dSt.data_ptr(), dNs.data_ptr(), stream))
    run(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    steps = dNs[:N].double().sum().item() + dNs[N:].double().sum().item()
    print("variant %d: %.1f ms  -> %.0f solves/s   (%.3g step attempts/s, mean steps %.0f)" % (var, ms, N / ms * 1e3, steps / ms * 1e3, steps / N))
os.environ["GYNC_LLK_VARIANT"] = "0"; L.check(lib.gync_init(0))

# EM
for K in (100_000, 1_000_000):
    M = 40
    Lm = torch.rand(M, K, dtype=torch.float64, device="cuda"); Lm /= Lm.sum(1, keepdim=True)   # (M,K) row-major == K x M col-major
    w = torch.full((K,), 1.0 / K, dtype=torch.float64, device="cuda")
    L.check(lib.gync_em_iterate_dev(w.data_ptr(), Lm.data_ptr(), K, M, 10, None, stream)); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); L.check(lib.gync_em_iterate_dev(w.data_ptr(), Lm.data_ptr(), K, M, 100, None, stream)); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 100
    by = 8 * K * M + 16 * K + 16 * M
    print("EM K=%d: %.2f us/iter  %.0f iters/s  %.0f GB/s algorithmic" % (K, ms * 1e3, 1e3 / ms, by / ms / 1e6))
    # parity vs oracle on K=1e5
    if K == 100_000:
        w2 = np.full(K, 1.0 / K); Lh = Lm.T.cpu().numpy()
        wo = w2.copy()
        for _ in range(5): wo = o.emiteration(wo, Lh)
        g.emiteration_bang(w2, Lh, niter=5)
        print("EM parity (5 iters) max rel:", np.max(np.abs(w2 - wo) / wo))
