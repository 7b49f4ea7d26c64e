"""A/B timing of the EM kernels of several builds in one process.  GYNC_LIBS=a.so,b.so python tools/ab_em.py"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
libs = os.environ["GYNC_LIBS"].split(",")
st = torch.cuda.current_stream().cuda_stream
hs = []
for p in libs:
    h = C.CDLL(os.path.join(ROOT, p)); assert h.gync_init(0) == 0
    h.gync_em_iterate_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    hs.append(h)
for K in [int(k) for k in os.environ.get("KS", "100000,1000000").split(",")]:
    M = 40
    Lm = torch.rand(M, K, dtype=torch.float64, device="cuda"); Lm /= Lm.sum(1, keepdim=True)
    ws = []
    for p, h in zip(libs, hs):
        best = 1e9
        w = torch.full((K,), 1.0 / K, dtype=torch.float64, device="cuda")
        h.gync_em_iterate_dev(w.data_ptr(), Lm.data_ptr(), K, M, 10, None, st); torch.cuda.synchronize()
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); assert h.gync_em_iterate_dev(w.data_ptr(), Lm.data_ptr(), K, M, 100, None, st) == 0; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 100)
        ws.append(w)
        print("K=%d %-24s %.2f us/iter  (%.0f iters/s)  max rel diff vs first %.1e" % (K, p, best * 1e3, 1e3 / best, float(((w - ws[0]).abs() / ws[0]).max())))
