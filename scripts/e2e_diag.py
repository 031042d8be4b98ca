"""Where does the host-buffer (e2e) path spend its time?  Times the two C-ABI phases separately."""
import ctypes as C, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sparse_matrix_b200 as S
from sparse_matrix_b200 import generators as G
from bench import pinned_array

MODE = sys.argv[1] if len(sys.argv) > 1 else "plain"
if MODE != "plain":
    import torch
    torch.cuda.init(); torch.zeros(1, device="cuda")
mat = G.poisson2d(2048)
rows, cols = mat[0], mat[1]
nnz_a = len(mat[3])
h = S.Handle(0); L = h.L
if MODE == "torch_stream":
    h.set_stream(torch.cuda.current_stream().cuda_stream)
print("MODE", MODE)
keep = []
p_ptr = pinned_array(L, rows + 1, np.uint64, keep); p_ptr[:] = mat[2]
p_idx = pinned_array(L, nnz_a, np.uint64, keep); p_idx[:] = mat[3]
p_val = pinned_array(L, nnz_a, np.float64, keep); p_val[:] = mat[4]
nnz_c = 54_484_996
c_ptr = pinned_array(L, rows + 1, np.uint64, keep)
c_idx = pinned_array(L, nnz_c, np.uint64, keep)
c_val = pinned_array(L, nnz_c, np.float64, keep)
P = S._lib.ptr
for it in range(4):
    nz = C.c_uint64()
    t0 = time.perf_counter()
    st = L.spam_spgemm_symbolic(h.h, 1, rows, cols, P(p_ptr), P(p_idx), P(p_val), rows, cols, P(p_ptr), P(p_idx), P(p_val), P(c_ptr), C.byref(nz))
    t1 = time.perf_counter()
    st2 = L.spam_spgemm_numeric(h.h, P(c_idx), P(c_val), 1)
    t2 = time.perf_counter()
    print(f"it{it}: symbolic rc={st} {1e3*(t1-t0):7.2f} ms   numeric rc={st2} {1e3*(t2-t1):7.2f} ms  nnz={nz.value}")
# pieces
for it in range(2):
    out = C.c_void_p()
    t0 = time.perf_counter()
    L.spam_csr_upload(h.h, 1, rows, cols, nnz_a, P(p_ptr), P(p_idx), P(p_val), C.byref(out)); L.spam_cuda_synchronize(h.h)
    t1 = time.perf_counter()
    c = C.c_void_p()
    L.spam_spgemm_dev(h.h, out, out, C.byref(c)); L.spam_cuda_synchronize(h.h)
    t2 = time.perf_counter()
    L.spam_dcsr_download(h.h, c, P(c_ptr), P(c_idx), P(c_val))
    t3 = time.perf_counter()
    L.spam_dcsr_free(h.h, c); L.spam_dcsr_free(h.h, out); L.spam_cuda_synchronize(h.h)
    t4 = time.perf_counter()
    print(f"pieces it{it}: upload {1e3*(t1-t0):.2f}  spgemm_dev {1e3*(t2-t1):.2f}  download {1e3*(t3-t2):.2f}  free {1e3*(t4-t3):.2f} ms")
