"""EM kernel with a partly resident shard (rows that fit stay in shared memory, the rest streams) against the all-streaming
kernel, same library, switched per launch through GYNC_EM_NO_HYBRID.  KS=125000,140000 python tools/em_hybrid_ab.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
libc = C.CDLL(None)
h = C.CDLL(os.path.join(ROOT, os.environ.get("GYNC_LIB", "gync.jl_b200/csrc/libgync_b200.so"))); assert h.gync_init(0) == 0
h.gync_em_iterate_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
st = torch.cuda.current_stream().cuda_stream
M = 40
for K in [int(k) for k in os.environ.get("KS", "100000,115000,125000,140000,160000,1000000").split(",")]:
    torch.manual_seed(K)
    Lm = torch.rand(M, K, dtype=torch.float64, device="cuda"); Lm /= Lm.sum(1, keepdim=True)
    ws = []
    for mode in ("hybrid", "streaming"):
        if mode == "streaming": libc.setenv(b"GYNC_EM_NO_HYBRID", b"1", 1)
        else: libc.unsetenv(b"GYNC_EM_NO_HYBRID")
        w = torch.full((K,), 1.0 / K, dtype=torch.float64, device="cuda")
        hist = torch.zeros(7, K, dtype=torch.float64, device="cuda")
        assert h.gync_em_iterate_dev(w.data_ptr(), Lm.data_ptr(), K, M, 7, hist.data_ptr(), st) == 0; torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); assert h.gync_em_iterate_dev(w.data_ptr(), Lm.data_ptr(), K, M, 100, None, st) == 0; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 100)
        ws.append((w, hist))
        d = float(((w - ws[0][0]).abs() / ws[0][0]).max()); dh = float(((hist - ws[0][1]).abs() / ws[0][1]).max())
        print("K=%-8d %-10s %.2f us/iter  max rel diff of w / history vs hybrid: %.1e / %.1e  sum(w)-1 = %.1e"
              % (K, mode, best * 1e3, d, dh, float(w.sum()) - 1))
