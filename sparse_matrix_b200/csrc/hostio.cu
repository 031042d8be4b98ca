// hostio.cu — host <-> device copies of the host-buffer entry points.
//
// The reference's caller owns ordinary heap memory (Rust Vecs: spam_csr/src/lib.rs:25-32 fields, and the result
// Vecs of mul_hash.rs:119).  cudaMemcpyAsync from or to PAGEABLE memory goes through the driver's own small bounce
// buffer at 10-15 GB/s; page-locked memory moves at the PCIe rate (~55 GB/s on this pool).  Measured through the
// two-phase C ABI on Poisson 2048^2: 23.7 ms with pinned buffers, 88 ms with pageable ones.  So:
//   * a buffer that is already page-locked (cudaHostAlloc / cudaHostRegister / spam_host_alloc) is copied directly;
//   * a pageable buffer is moved through the handle's own ring of pinned slots: host threads memcpy a slot while
//     the DMA engine moves the previous one (H2D), or drain a slot while the next ones arrive (D2H).
// Registering the caller's buffer on the fly (cudaHostRegister) costs more than the copy for a one-shot call.
//
// Index width: col_idx crosses the bus as u64 (the reference's usize) and is narrowed / widened on the device.
// Converting on the host instead would ship 24% fewer bytes, but a 16-thread host widens 54 M indices in ~25 ms
// (12 bytes of memory traffic per index at 26 GB/s, measured), against 8 ms to ship the zero bytes.
#include <cstring>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

constexpr int HS_SLOTS = 4;
constexpr size_t HS_SLOT_BYTES = (size_t)32 << 20;

bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

void parallel_memcpy(void* dst, const void* src, size_t bytes, int nthreads) {
  if (bytes < ((size_t)4 << 20) || nthreads <= 1) { memcpy(dst, src, bytes); return; }
  std::vector<std::thread> th;
  th.reserve(nthreads);
  for (int t = 0; t < nthreads; ++t) {
    const size_t a = (bytes * t / nthreads) & ~(size_t)63, b = t + 1 == nthreads ? bytes : (bytes * (t + 1) / nthreads) & ~(size_t)63;
    th.emplace_back([=] { memcpy((char*)dst + a, (const char*)src + a, b - a); });
  }
  for (auto& x : th) x.join();
}

int ensure_stage(spam_handle* h) {
  if (h->stage) return SPAM_OK;
  HostStage* s = new HostStage();
  s->buf = nullptr;
  for (auto& e : s->ev) e = nullptr;
  unsigned hw = std::thread::hardware_concurrency();
  const char* env = getenv("SPAM_HOST_THREADS");
  int nt = env ? atoi(env) : (int)(hw > 16 ? 16 : hw);
  s->nthreads = nt < 1 ? 1 : nt;
  cudaError_t e = cudaHostAlloc((void**)&s->buf, HS_SLOTS * HS_SLOT_BYTES, cudaHostAllocDefault);
  for (int i = 0; i < HS_SLOTS && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&s->ev[i], cudaEventDisableTiming);
  h->stage = s;
  if (e != cudaSuccess) { host_stage_free(h); cudaGetLastError(); return spam_fail(h, SPAM_ENOMEM, "pinned staging ring", e); }
  return SPAM_OK;
}

}  // namespace

void host_stage_free(spam_handle* h) {
  HostStage* s = h->stage;
  if (!s) return;
  for (auto& e : s->ev) if (e) cudaEventDestroy(e);
  if (s->buf) cudaFreeHost(s->buf);
  delete s;
  h->stage = nullptr;
}

// asynchronous on h->stream when src is page-locked; for pageable src the call returns once the last slot has been
// handed to the DMA engine (the source may be reused, like a pageable cudaMemcpyAsync)
int host_to_dev(spam_handle* h, void* d_dst, const void* h_src, size_t bytes) {
  if (!bytes) return SPAM_OK;
  if (is_pinned(h_src) || bytes < ((size_t)1 << 20)) {
    CK(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, h->stream));
    return SPAM_OK;
  }
  CKS(ensure_stage(h));
  HostStage* s = h->stage;
  size_t off = 0;
  for (int c = 0; off < bytes; ++c, off += HS_SLOT_BYTES) {
    const int slot = c % HS_SLOTS;
    const size_t n = bytes - off < HS_SLOT_BYTES ? bytes - off : HS_SLOT_BYTES;
    unsigned char* sb = s->buf + (size_t)slot * HS_SLOT_BYTES;
    CK(cudaEventSynchronize(s->ev[slot]));  // the DMA that last read this slot is done (no-op the first time)
    parallel_memcpy(sb, (const char*)h_src + off, n, s->nthreads);
    CK(cudaMemcpyAsync((char*)d_dst + off, sb, n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaEventRecord(s->ev[slot], h->stream));
  }
  return SPAM_OK;
}

// page-locked dst: asynchronous on h->stream (the caller synchronises); pageable dst: returns when the data is there
int dev_to_host(spam_handle* h, void* h_dst, const void* d_src, size_t bytes) {
  if (!bytes) return SPAM_OK;
  if (is_pinned(h_dst) || bytes < ((size_t)1 << 20)) {
    CK(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, h->stream));
    return SPAM_OK;
  }
  CKS(ensure_stage(h));
  HostStage* s = h->stage;
  const size_t nchunks = (bytes + HS_SLOT_BYTES - 1) / HS_SLOT_BYTES;
  auto issue = [&](size_t c) -> cudaError_t {
    const size_t off = c * HS_SLOT_BYTES, n = bytes - off < HS_SLOT_BYTES ? bytes - off : HS_SLOT_BYTES;
    const int slot = (int)(c % HS_SLOTS);
    cudaError_t e = cudaMemcpyAsync(s->buf + (size_t)slot * HS_SLOT_BYTES, (const char*)d_src + off, n, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaEventRecord(s->ev[slot], h->stream);
    return e;
  };
  for (size_t c = 0; c < nchunks && c < (size_t)HS_SLOTS; ++c) CK(issue(c));
  for (size_t c = 0; c < nchunks; ++c) {
    const size_t off = c * HS_SLOT_BYTES, n = bytes - off < HS_SLOT_BYTES ? bytes - off : HS_SLOT_BYTES;
    const int slot = (int)(c % HS_SLOTS);
    CK(cudaEventSynchronize(s->ev[slot]));
    parallel_memcpy((char*)h_dst + off, s->buf + (size_t)slot * HS_SLOT_BYTES, n, s->nthreads);
    if (c + HS_SLOTS < nchunks) CK(issue(c + HS_SLOTS));
  }
  return SPAM_OK;
}
