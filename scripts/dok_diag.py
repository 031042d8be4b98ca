import ctypes as C, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sparse_matrix_b200 as S
from sparse_matrix_b200 import generators as G
a = G.uniform_random(1_000_000, 4_000_000, 8, seed=5, dtype=np.int64, int_range=1 << 15)
tr, tc, tv = G.triplets_with_rewrites(a, seed=5)
n = len(tv)
dev = torch.device("cuda", 0)
h = S.Handle(0); L = h.L
d_r = torch.from_numpy(tr.view(np.int64)).to(dev); d_c = torch.from_numpy(tc.view(np.int64)).to(dev); d_v = torch.from_numpy(tv).to(dev)
torch.cuda.synchronize()
prev = None
for it in range(12):
    out = C.c_void_p()
    t0 = time.perf_counter()
    st = L.spam_dok_to_csr_dev(h.h, 3, a[0], a[1], n, C.c_void_p(d_r.data_ptr()), C.c_void_p(d_c.data_ptr()), C.c_void_p(d_v.data_ptr()), C.byref(out))
    t1 = time.perf_counter()
    L.spam_cuda_synchronize(h.h)
    t2 = time.perf_counter()
    if prev: L.spam_dcsr_free(h.h, prev)
    prev = out
    t3 = time.perf_counter()
    print(f"it{it}: call {1e3*(t1-t0):.2f} ms  sync {1e3*(t2-t1):.2f} ms  free {1e3*(t3-t2):.3f} ms rc={st}")
