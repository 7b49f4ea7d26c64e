"""Small driver for ncu: a few launches of each empirical-Bayes kernel (likelihood matrix, one-pass dhz, the two-pass
matvec pair, dlogl, the fused projected-ascent step) at the bench sizes."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gync_b200 import lib as L
lib = L.load(); L.check(lib.gync_init(0))
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
K, M, D = int(os.environ.get("K", 10000)), 40, 124
gen = torch.Generator(device=dev); gen.manual_seed(99)
sig = torch.tensor(np.tile(np.array([12.0, 1.0, 40.0, 1.5]), (31, 1)).ravel(order="F"), device=dev)
Y = torch.randn(D, K, dtype=torch.float64, device=dev, generator=gen) * 0.3
Zs = Y + torch.randn(D, K, dtype=torch.float64, device=dev, generator=gen) * 0.1
Lz = torch.empty(K, K, dtype=torch.float64, device=dev)
for _ in range(2):
    L.check(lib.gync_likelihoodmat_dev(Zs.data_ptr(), K, K, 1, Y.data_ptr(), K, K, 1, D, sig.data_ptr(), 1, Lz.data_ptr(), st))
B = torch.randn(M, D, dtype=torch.float64, device=dev, generator=gen)
B[torch.rand(M, D, device=dev, generator=gen) < 0.4] = float("nan")
Ld = torch.empty(M, K, dtype=torch.float64, device=dev)
L.check(lib.gync_likelihoodmat_dev(Y.data_ptr(), K, K, 1, B.data_ptr(), M, 1, D, D, sig.data_ptr(), 1, Ld.data_ptr(), st))
torch.cuda.synchronize()
h = C.c_void_p()
L.check(lib.gync_ebmodel_from_matrices(L.dptr(np.asfortranarray(Ld.cpu().numpy().T)), K, M,
                                       L.dptr(np.asfortranarray(Lz.cpu().numpy().T)), K, C.byref(h)))
w = torch.full((K,), 1.0 / K, dtype=torch.float64, device=dev)
ga, gb = torch.empty_like(w), torch.empty_like(w)


def grads(which, reps):
    for _ in range(reps):
        if which & 1:
            L.check(lib.gync_ebmodel_dhz_dev(h, w.data_ptr(), 0, ga.data_ptr(), st))
        if which & 2:
            L.check(lib.gync_ebmodel_dlogl_dev(h, w.data_ptr(), gb.data_ptr(), st))


grads(3, 3)          # one-pass dhz + dlogl
os.environ["GYNC_DHZ_TWO_PASS"] = "1"
grads(1, 2)          # two-pass dhz
del os.environ["GYNC_DHZ_TWO_PASS"]
torch.cuda.synchronize()
wh = np.full(K, 1.0 / K)
L.check(lib.gync_ebmodel_mple(h, L.dptr(wh), 0.5, 1e-6, 3, 0, None))
print("sum w", wh.sum())
lib.gync_ebmodel_destroy(h)
