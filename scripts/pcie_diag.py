"""PCIe copy-rate diagnostic: torch pinned tensors vs spam_host_alloc buffers through the library."""
import ctypes as C, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sparse_matrix_b200 as S

n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, f in (("torch pinned H2D", lambda: d.copy_(h, non_blocking=True)), ("torch pinned D2H", lambda: h.copy_(d, non_blocking=True))):
    f(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5): f()
    torch.cuda.synchronize()
    print(f"{name}: {5 * n / (time.perf_counter() - t) / 1e9:.1f} GB/s")
hp = torch.empty(n, dtype=torch.uint8)
t = time.perf_counter(); d.copy_(hp); torch.cuda.synchronize(); print(f"torch pageable H2D: {n / (time.perf_counter() - t) / 1e9:.1f} GB/s")
hd = S.Handle(0); L = hd.L
p = C.c_void_p(); assert L.spam_host_alloc(C.byref(p), n) == 0
rt = C.CDLL("libcudart.so.12") if False else None
# copy via the library: upload a fake CSR of nnz entries (u64 idx + f64 val = 16 B/entry)
nnz = n // 16
buf = np.frombuffer((C.c_char * n).from_address(p.value), dtype=np.uint64, count=nnz * 2)
buf[:] = 0
ptr = np.zeros(2, dtype=np.uint64); ptr[1] = nnz
out = C.c_void_p()
for rep in range(3):
    t = time.perf_counter()
    st = L.spam_csr_upload(hd.h, 1, 1, 1 << 20, nnz, ptr.ctypes.data, buf[:nnz].ctypes.data, buf[nnz:].ctypes.data, C.byref(out))
    L.spam_cuda_synchronize(hd.h)
    dt = time.perf_counter() - t
    print(f"spam_csr_upload rep{rep} rc={st}: {nnz * 16 / dt / 1e9:.1f} GB/s ({dt*1e3:.1f} ms)")
    o_idx = buf[:nnz]; o_val = buf[nnz:]
    t = time.perf_counter()
    st = L.spam_dcsr_download(hd.h, out, None, o_idx.ctypes.data, o_val.ctypes.data)
    dt = time.perf_counter() - t
    print(f"spam_dcsr_download rep{rep} rc={st}: {nnz * 16 / dt / 1e9:.1f} GB/s ({dt*1e3:.1f} ms)")
    L.spam_dcsr_free(hd.h, out)
