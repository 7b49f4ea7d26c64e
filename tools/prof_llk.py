"""Small single-launch driver for ncu: one loglik launch (and optionally one EM launch)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gync_b200 as g
from gync_b200 import lib as L
lib = L.load(); L.check(lib.gync_init(0))
N = int(os.environ.get("N", 9472)); rtol = float(os.environ.get("RTOL", 1e-8)); atol = float(os.environ.get("ATOL", 1e-8))
rng = np.random.default_rng(20160911 + 2)
x0 = np.concatenate([g.api.refparms, g.api.refy0, [28.0]])
X = np.exp(np.log(x0)[None, :] + 0.0998 * rng.standard_normal((N, 116)))
c = g.Config()
dX = torch.from_numpy(np.asfortranarray(X).T.copy()).cuda()
dD = torch.from_numpy(np.asfortranarray(c.data).T.copy()).cuda()
dS = torch.tensor([0.1], dtype=torch.float64, device="cuda")
dLL = torch.empty(N, dtype=torch.float64, device="cuda")
op = L.opts(rtol, atol)
st = torch.cuda.current_stream().cuda_stream
for _ in range(int(os.environ.get("REPS", 2))):
    L.check(lib.gync_loglik_batch_dev(dX.data_ptr(), N, dD.data_ptr(), 1, dS.data_ptr(), C.byref(op), dLL.data_ptr(), None, None, None, st))
torch.cuda.synchronize()
print("llh mean", dLL[torch.isfinite(dLL)].mean().item())
if os.environ.get("EM"):
    K = int(os.environ["EM"]); M = 40
    Lm = torch.rand(M, K, dtype=torch.float64, device="cuda"); Lm /= Lm.sum(1, keepdim=True)
    w = torch.full((K,), 1.0 / K, dtype=torch.float64, device="cuda")
    for _ in range(2):
        L.check(lib.gync_em_iterate_dev(w.data_ptr(), Lm.data_ptr(), K, M, 20, None, st))
    torch.cuda.synchronize(); print("em sum", w.sum().item())
