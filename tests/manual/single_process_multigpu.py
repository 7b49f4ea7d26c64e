"""Single-process multi-GPU (gync_init_multi): what a Julia host, which is one process, would call.
Run on a box with >= 2 GPUs:  python tests/manual/single_process_multigpu.py
Checks: sharded likelihood batch == single-device result bit for bit; sharded EM == oracle to 1e-10; timings."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import gync_b200 as g
from gync_b200 import lib as L
import gync_oracle as o
lib = L.load()
nd = torch.cuda.device_count()
print("devices:", nd)
rng = np.random.default_rng(31)
N = 40_000
X = np.exp(np.log(o.init_x())[None, :] + 0.0998 * rng.standard_normal((N, 116)))
la = o.patients()[0]
datas = [la[0], la[1], la[4], la[5], la[6], la[7]]
L.check(lib.gync_init(0))
t0 = time.time(); LL1, st1, Y1 = g.loglik_matrix(X, datas, [0.1] * 6, want_traj=True); t1 = time.time() - t0
L.check(lib.gync_init_multi(nd, None)); assert lib.gync_ndev() == nd
t0 = time.time(); LLm, stm, Ym = g.loglik_matrix(X, datas, [0.1] * 6, want_traj=True); tm = time.time() - t0
assert np.array_equal(LL1, LLm) and np.array_equal(st1, stm) and np.array_equal(Y1, Ym, equal_nan=True)
print(f"loglik {N} x 6: 1 GPU {t1:.2f} s, {nd} GPUs {tm:.2f} s (host-buffer API incl. copies)  identical: True")
K, M, niter = 1_000_000, 40, 50
Lm = np.asfortranarray(rng.random((K, M))); Lm /= Lm.sum(0, keepdims=True)
w0 = rng.random(K); w0 /= w0.sum()
w = w0.copy()
t0 = time.time(); g.emiteration_bang(w, Lm, niter=niter); tm = time.time() - t0
L.check(lib.gync_init(0))
w1 = w0.copy()
t0 = time.time(); g.emiteration_bang(w1, Lm, niter=niter); t1 = time.time() - t0
ref = w0.copy()
for _ in range(3):
    ref = o.emiteration(ref, Lm)
w3 = w0.copy(); L.check(lib.gync_init_multi(nd, None)); g.emiteration_bang(w3, Lm, niter=3)
print(f"EM K=1e6 x 40, {niter} iters (host-buffer API incl. copies): 1 GPU {t1:.2f} s, {nd} GPUs {tm:.2f} s;"
      f" multi vs single max rel {np.max(np.abs(w - w1) / w1):.1e}; multi vs oracle (3 iters) {np.max(np.abs(w3 - ref) / ref):.1e}")
assert np.max(np.abs(w - w1) / w1) < 1e-10 and np.max(np.abs(w3 - ref) / ref) < 1e-10
lib.gync_shutdown()
print("ok")
