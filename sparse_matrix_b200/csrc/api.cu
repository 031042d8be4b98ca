// api.cu — the extern "C" boundary of libspam_cuda.so (see include/spam_cuda.h for what each
// entry point replaces in the reference).  Plain pointers and sizes only; no exceptions cross it.
#include <cstdio>
#include <cstring>
#include <new>

#include "common.cuh"

int spam_fail(spam_handle* h, int status, const char* what, cudaError_t ce) {
  if (h) {
    h->err = what ? what : "";
    if (ce != cudaSuccess) {
      h->err += ": ";
      h->err += cudaGetErrorString(ce);
    }
  }
  return status;
}

int dev_alloc(spam_handle* h, void** p, size_t bytes) {
  *p = nullptr;
  if (bytes == 0) bytes = 16;
  cudaError_t e = h->pool ? cudaMallocFromPoolAsync(p, bytes, h->pool, h->stream) : cudaMallocAsync(p, bytes, h->stream);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    return spam_fail(h, SPAM_ENOMEM, "device allocation failed", e);
  }
  if (e != cudaSuccess) return spam_fail(h, SPAM_ECUDA, "cudaMallocAsync", e);
  return SPAM_OK;
}

int dev_free(spam_handle* h, void* p) {
  if (!p) return SPAM_OK;
  cudaError_t e = cudaFreeAsync(p, h->stream);
  if (e != cudaSuccess) return spam_fail(h, SPAM_ECUDA, "cudaFreeAsync", e);
  return SPAM_OK;
}

struct SpgemmHostState {  // what lives between spam_spgemm_symbolic and spam_spgemm_numeric
  spam_dcsr* a;
  spam_dcsr* b;  // may alias a
  SpgemmPending* pend;
};
struct DokPending {
  spam_dcsr* c;
};

namespace {

__global__ void k_partition_points(const u64* __restrict__ ps, u64 m, u32 parts, u64* __restrict__ out) {
  // mul_hash.rs:52-62: avg = ceil(total / tnum); rows_offset[t] = partition_point(ps <= avg*t) - 1
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > parts) return;
  if (t == 0) { out[0] = 0; return; }
  if (t == parts) { out[parts] = m; return; }
  const u64 total = ps[m];
  const u64 avg = (total + parts - 1) / parts;
  const u64 bound = avg * t;
  u64 lo = 0, hi = m + 1;  // first index with ps[idx] > bound
  while (lo < hi) {
    const u64 mid = (lo + hi) >> 1;
    if (ps[mid] <= bound) lo = mid + 1; else hi = mid;
  }
  out[t] = lo - 1;  // ps[0] = 0 <= bound so lo >= 1
}

// Device-time cost of a row of C as a function of its intermediate-product count f: f x a per-product
// weight in 1/16 units, fitted to the measured times of 48 row blocks of R-MAT scale 22 on B200
// (profiles/r01_rmat22_v3_balance.txt): every shared-memory bin costs about 30 ps per product (the merge
// bin slightly more: random gathers of short B rows), rows of the global-table bin about 120 ps (L2
// atomics).  Shard times predicted by this table are within 3% (rms) of the measured ones.
__device__ __forceinline__ u32 row_cost_q(u32 f, u64 alen) {
  u32 w;
  if (f <= 128) w = 18;           // merge / tiny
  else if (f <= 8192) w = 16;     // one-warp and team hash bins
  else w = 64;                    // global-table bin
  (void)alen;
  const u64 c = (u64)f * w;
  return c > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)c;
}

__global__ void __launch_bounds__(256) k_flop_to_cost(u32* __restrict__ flop, const u64* __restrict__ a_ptr, u64 m) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) flop[i] = row_cost_q(flop[i], a_ptr[i + 1] - a_ptr[i]);
}

// Column range of A (= the rows of B the product can touch): max into work_b, min as max(~col) into work_a,
// both zeroed by the caller.
__global__ void __launch_bounds__(256) k_col_range(const u32* __restrict__ idx, u64 nnz, Counters* cnt) {
  u32 mx = 0, nmn = 0;
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) {
    const u32 c = idx[i];
    mx = max(mx, c);
    nmn = max(nmn, ~c);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    nmn = max(nmn, __shfl_xor_sync(0xffffffffu, nmn, d));
  }
  if ((threadIdx.x & 31) == 0) { atomicMax(&cnt->work_b, mx); atomicMax(&cnt->work_a, nmn); }
}

// row_ptr of a matrix of which only rows [r0, r1) were uploaded (entries r0..r1 of the host array sit at
// ptr[r0..r1], still absolute): rebase them and make every other row empty.
__global__ void __launch_bounds__(256) k_window_ptr(u64* __restrict__ ptr, u64 n_entries, u64 r0, u64 r1, u64 base,
                                                    u64 nnz_w) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_entries; i += stride)
    ptr[i] = i < r0 ? 0 : (i > r1 ? nnz_w : ptr[i] - base);
}

__global__ void __launch_bounds__(256) k_rebase_ptr(const u64* __restrict__ in, u64* __restrict__ out, u64 n) {
  const u64 base = in[0];
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i] - base;
}

// rows of a matrix picked by index: lengths, then (after the scan) one warp per picked row copies it
__global__ void __launch_bounds__(256) k_select_len(const u64* __restrict__ ptr, const u64* __restrict__ rows, u64 n,
                                                    u64 m, u32* __restrict__ len, Counters* cnt) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 r = rows[i];
  if (r >= m) { atomicOr(&cnt->error, 2u); len[i] = 0; return; }
  const u64 l = ptr[r + 1] - ptr[r];
  len[i] = l > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)l;
}
template <class W>
__global__ void __launch_bounds__(256) k_select_copy(const u64* __restrict__ ptr, const u32* __restrict__ idx,
                                                     const W* __restrict__ val, const u64* __restrict__ rows, u64 n,
                                                     const u64* __restrict__ optr, u32* __restrict__ oidx,
                                                     W* __restrict__ oval) {
  const u64 w = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const u64 r = rows[w];
  const u64 lo = ptr[r], hi = ptr[r + 1], o = optr[w];
  for (u64 e = lo + lane; e < hi; e += 32) { oidx[o + (e - lo)] = idx[e]; oval[o + (e - lo)] = val[e]; }
}

int set_device(spam_handle* h) {
  CK(cudaSetDevice(h->device));
  return SPAM_OK;
}

void free_dcsr(spam_handle* h, spam_dcsr* m) {
  if (!m) return;
  if (m->sorted_copy) { free_dcsr(h, m->sorted_copy); m->sorted_copy = nullptr; }
  if (m->owning) {
    dev_free(h, m->ptr);
    dev_free(h, m->idx);
    dev_free(h, m->val);
  }
  delete m;
}

void drop_spgemm_state(spam_handle* h) {
  SpgemmHostState* s = static_cast<SpgemmHostState*>(h->pending);
  if (!s) return;
  if (s->pend) spgemm_pending_free(h, s->pend);
  if (s->b && s->b != s->a) free_dcsr(h, s->b);
  free_dcsr(h, s->a);
  delete s;
  h->pending = nullptr;
}

void drop_dok_state(spam_handle* h) {
  if (!h->dok_pending) return;
  free_dcsr(h, h->dok_pending->c);
  delete h->dok_pending;
  h->dok_pending = nullptr;
}

}  // namespace
void free_dcsr_tree(spam_handle* h, spam_dcsr* m) { free_dcsr(h, m); }
namespace {
int valid_dtype(int dt) { return dt == SPAM_F32 || dt == SPAM_F64 || dt == SPAM_I32 || dt == SPAM_I64; }

}  // namespace

void timing_begin_product(spam_handle* h) {
  if (!h->timing) return;
  const int other = h->ev_cur ^ 1;
  if (h->ev_pending[other]) {  // never harvested (no sync covered it yet): wait for it rather than lose it
    if (cudaEventSynchronize(h->evs[other][4]) == cudaSuccess) timing_harvest(h, other);
    h->ev_pending[other] = false;
  }
  h->ev_cur = other;
  h->ev = h->evs[other];
}

void timing_harvest(spam_handle* h, int set) {
  if (!h->ev_pending[set]) return;
  h->ev_pending[set] = false;
  cudaEvent_t* e = h->evs[set];
  float t[5];
  if (cudaEventElapsedTime(&t[0], e[0], e[1]) != cudaSuccess || cudaEventElapsedTime(&t[1], e[1], e[2]) != cudaSuccess ||
      cudaEventElapsedTime(&t[2], e[2], e[3]) != cudaSuccess || cudaEventElapsedTime(&t[3], e[3], e[4]) != cudaSuccess ||
      cudaEventElapsedTime(&t[4], e[0], e[4]) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  h->stats.ms_flop = t[0]; h->stats.ms_symbolic = t[1]; h->stats.ms_scan = t[2]; h->stats.ms_numeric = t[3];
  h->stats.ms_total = t[4];
  for (int i = 0; i < 5; ++i) h->acc_ms[i] += t[i];
  h->acc_n += 1;
}

void finish_timing(spam_handle* h) {
  if (!h->timing || !h->ev_pending[h->ev_cur]) return;
  if (cudaEventSynchronize(h->ev[4]) != cudaSuccess) return;
  timing_harvest(h, h->ev_cur);
}

extern "C" {

int spam_cuda_abi_version(void) { return 6; }  // 2: 16-entry bin arrays; 3: spam_rows_to_parts_cost; 4: transpose, phase totals; 5: ewise, spam_mm_parse; 6: stats.fallbacks/paths, sorted = 0, spam_comm_*, pageable host path

const char* spam_strerror(int s) {
  switch (s) {
    case SPAM_OK: return "ok";
    case SPAM_EINVAL: return "invalid argument";
    case SPAM_EDIM: return "dimension mismatch: A.cols != B.rows";
    case SPAM_ECOLS: return "dimension >= 2^32-1 not representable (u32::MAX is the empty-slot sentinel)";
    case SPAM_ENOMEM: return "out of memory";
    case SPAM_ECUDA: return "CUDA error";
    case SPAM_ESTATE: return "call out of sequence (no pending symbolic/build phase)";
    case SPAM_EOVERFLOW: return "flop or nnz count overflow";
    case SPAM_EINDEX: return "index out of range";
    case SPAM_EDTYPE: return "operand dtypes differ";
    default: return "unknown status";
  }
}

const char* spam_last_error(const spam_handle* h) { return h ? h->err.c_str() : "null handle"; }

int spam_cuda_create(spam_handle** out, int device) {
  if (!out) return SPAM_EINVAL;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return SPAM_ECUDA; }  // no CPU fallback
  if (device < 0 || device >= ndev) return SPAM_EINVAL;
  spam_handle* h = new (std::nothrow) spam_handle();
  if (!h) return SPAM_ENOMEM;
  h->device = device; h->timing = false; h->pending = nullptr; h->dok_pending = nullptr;
  h->scan_ws = nullptr; h->scan_ws_cap = 0;
  h->pool = nullptr;
  h->comm = nullptr;
  h->stage = nullptr;
  {
    const char* e = getenv("SPAM_LANES");  // read once: 0 keeps every row bin on the main stream
    h->use_lanes = !(e && e[0] == '0');
    e = getenv("SPAM_SORT_B");
    h->sort_b = !(e && e[0] == '0');
    e = getenv("SPAM_MERGE_WIN");
    h->merge_win = e ? atoi(e) & 3 : 0;
    e = getenv("SPAM_EWISE_TMA");
    h->ewise_tma = !(e && e[0] == '0');
    e = getenv("SPAM_DOK_BUCKET");
    h->dok_bucket = !(e && e[0] == '0');
    e = getenv("SPAM_MERGE_PERSIST");
    h->merge_persist = e ? atoi(e) : 0;
    e = getenv("SPAM_ONEPASS");
    h->onepass = e && e[0] == '1';
    e = getenv("SPAM_SPMV_TMA");
    h->spmv_tma = e && e[0] == '1';
    e = getenv("SPAM_MERGE_PF");
    h->merge_pf = e ? atoi(e) & 7 : -1;
    e = getenv("SPAM_L2_PERSIST");
    h->l2_persist = e ? atoi(e) : 0;
    h->l2_persist_max = 0; h->l2_window_max = 0;
    e = getenv("SPAM_ESC");
    h->use_esc = e ? (e[0] == '3' ? 3 : (e[0] == '2' ? 2 : (e[0] == '1' ? 1 : 0))) : 0;
  }
  h->d_cnt = nullptr; h->h_cnt = nullptr; h->own_stream = nullptr; h->stream = nullptr;
  h->stats = spam_stats{};
  for (auto& set : h->evs) for (auto& e : set) e = nullptr;
  h->ev = h->evs[0]; h->ev_cur = 0; h->ev_pending[0] = h->ev_pending[1] = false;
  for (auto& a : h->acc_ms) a = 0.0;
  h->acc_n = 0;
  for (auto& s : h->lane) s = nullptr;
  for (auto& e : h->lane_ev) e = nullptr;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
  h->stream = h->own_stream;
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_cnt, sizeof(Counters));
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&h->h_cnt, sizeof(Counters), cudaHostAllocDefault);
  for (int i = 0; i < 12 && e == cudaSuccess; ++i) e = cudaEventCreate(&h->evs[i / 6][i % 6]);
  for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&h->lane[i], cudaStreamNonBlocking);
  for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&h->lane_ev[i], cudaEventDisableTiming);
  cudaDeviceProp prop;
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
  if (e == cudaSuccess) {
    h->num_sms = prop.multiProcessorCount;
    h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (h->l2_persist && prop.persistingL2CacheMaxSize > 0 &&
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)prop.persistingL2CacheMaxSize) == cudaSuccess) {
      h->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
      h->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
    } else cudaGetLastError();
    // A PRIVATE stream-ordered pool that keeps its freed blocks: steady-state products do no cudaMalloc, and
    // the process-wide default pool (shared with torch, NCCL, ...) is left alone.
    cudaMemPoolProps pp = {};
    pp.allocType = cudaMemAllocationTypePinned;
    pp.handleTypes = cudaMemHandleTypeNone;
    pp.location.type = cudaMemLocationTypeDevice;
    pp.location.id = device;
    e = cudaMemPoolCreate(&h->pool, &pp);
    if (e == cudaSuccess) {
      unsigned long long thr = ~0ull;
      e = cudaMemPoolSetAttribute(h->pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    spam_cuda_destroy(h);
    return SPAM_ECUDA;
  }
  *out = h;
  return SPAM_OK;
}

int spam_cuda_destroy(spam_handle* h) {
  if (!h) return SPAM_EINVAL;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->comm) spam_comm_destroy(h);
  host_stage_free(h);
  drop_spgemm_state(h);
  drop_dok_state(h);
  if (h->scan_ws) { dev_free(h, h->scan_ws); h->scan_ws = nullptr; }
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto& set : h->evs) for (auto& e : set) if (e) cudaEventDestroy(e);
  for (auto& s : h->lane) if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
  for (auto& e : h->lane_ev) if (e) cudaEventDestroy(e);
  if (h->d_cnt) cudaFree(h->d_cnt);
  if (h->h_cnt) cudaFreeHost(h->h_cnt);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->pool) cudaMemPoolDestroy(h->pool);
  delete h;
  return SPAM_OK;
}

int spam_cuda_set_stream(spam_handle* h, void* s) {
  if (!h) return SPAM_EINVAL;
  cudaStream_t next = s ? (cudaStream_t)s : h->own_stream;
  if (next == h->stream) return SPAM_OK;
  // Everything queued so far (kernels, stream-ordered frees of scratch that the pool may hand out again) happens
  // before anything queued on the new stream.
  CKS(set_device(h));
  CK(cudaEventRecord(h->lane_ev[0], h->stream));
  CK(cudaStreamWaitEvent(next, h->lane_ev[0], 0));
  h->stream = next;
  return SPAM_OK;
}

int spam_cuda_set_timing(spam_handle* h, int enabled) {
  if (!h) return SPAM_EINVAL;
  h->timing = enabled != 0;
  return SPAM_OK;
}

int spam_cuda_get_stats(spam_handle* h, spam_stats* out) {
  if (!h || !out) return SPAM_EINVAL;
  finish_timing(h);
  // the rare-path counters of the last product live in the device counter block (zeroed per product)
  CKS(set_device(h));
  CK(cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->stats.fallbacks[0] = h->h_cnt->fb_warp_bitonic;
  h->stats.fallbacks[1] = h->h_cnt->fb_team_bitonic;
  h->stats.fallbacks[2] = h->h_cnt->fb_heavy_bitonic;
  h->stats.fallbacks[3] = h->h_cnt->fb_esc;
  *out = h->stats;
  return SPAM_OK;
}

int spam_cuda_get_phase_totals(spam_handle* h, double* ms5, uint64_t* products, int reset) {
  if (!h) return SPAM_EINVAL;
  finish_timing(h);
  if (h->ev_pending[h->ev_cur ^ 1] && cudaEventSynchronize(h->evs[h->ev_cur ^ 1][4]) == cudaSuccess)
    timing_harvest(h, h->ev_cur ^ 1);
  if (ms5) for (int i = 0; i < 5; ++i) ms5[i] = h->acc_ms[i];
  if (products) *products = h->acc_n;
  if (reset) {
    for (auto& a : h->acc_ms) a = 0.0;
    h->acc_n = 0;
  }
  return SPAM_OK;
}

int spam_cuda_synchronize(spam_handle* h) {
  if (!h) return SPAM_EINVAL;
  CKS(set_device(h));
  CK(cudaStreamSynchronize(h->stream));
  return SPAM_OK;
}

int spam_host_alloc(void** p, uint64_t bytes) {
  if (!p) return SPAM_EINVAL;
  cudaError_t e = cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault);
  if (e != cudaSuccess) { cudaGetLastError(); *p = nullptr; return SPAM_ENOMEM; }
  return SPAM_OK;
}
int spam_host_free(void* p) {
  if (!p) return SPAM_OK;
  return cudaFreeHost(p) == cudaSuccess ? SPAM_OK : SPAM_ECUDA;
}

/* ---------------- device-resident matrices ---------------- */

int spam_csr_upload(spam_handle* h, int dtype, uint64_t rows, uint64_t cols, uint64_t nnz, const uint64_t* ptr,
                    const uint64_t* idx, const void* val, spam_dcsr** out) {
  if (!h || !out || !ptr || (nnz && (!idx || !val)) || !valid_dtype(dtype)) return spam_fail(h, SPAM_EINVAL, "bad argument");
  *out = nullptr;
  if (rows >= 0xFFFFFFFFull || cols >= 0xFFFFFFFFull) return spam_fail(h, SPAM_ECOLS, "dimension >= 2^32-1");
  CKS(set_device(h));
  spam_dcsr* m = new spam_dcsr();
  m->dtype = dtype; m->rows = rows; m->cols = cols; m->nnz = nnz; m->owning = true; m->rows_sorted = -1; m->max_row_len = 0;
  m->ptr = nullptr; m->idx = nullptr; m->val = nullptr;
  u64* tmp = nullptr;
  int st = dev_alloc_t(h, &m->ptr, rows + 1);
  if (st == SPAM_OK) st = dev_alloc_t(h, &m->idx, nnz);
  if (st == SPAM_OK) st = dev_alloc(h, &m->val, nnz * dtype_size(dtype));
  if (st == SPAM_OK) st = dev_alloc_t(h, &tmp, nnz);
  if (st == SPAM_OK) st = host_to_dev(h, m->ptr, ptr, (rows + 1) * sizeof(u64));
  if (st == SPAM_OK && nnz) st = host_to_dev(h, tmp, idx, nnz * sizeof(u64));
  if (st == SPAM_OK) st = narrow_u64_to_u32(h, tmp, m->idx, nnz);   // runs while the values are still on the bus
  if (st == SPAM_OK && nnz) st = host_to_dev(h, m->val, val, nnz * dtype_size(dtype));
  dev_free(h, tmp);
  if (st != SPAM_OK) { free_dcsr(h, m); return st; }
  h->stats.bytes_h2d += (rows + 1) * 8 + nnz * (8 + dtype_size(dtype));
  *out = m;
  return SPAM_OK;
}

// Upload only rows [r0, r1) of a host CSR matrix; the device matrix keeps its full shape with every
// other row empty.  Used for B when A references a narrow band of its rows (a row block of a banded
// matrix times the whole matrix: each GPU of a row-sharded product needs only its halo of B).
static int upload_csr_window(spam_handle* h, int dtype, u64 rows, u64 cols, const u64* ptr, const u64* idx,
                             const void* val, u64 r0, u64 r1, spam_dcsr** out) {
  *out = nullptr;
  if (rows >= 0xFFFFFFFFull || cols >= 0xFFFFFFFFull) return spam_fail(h, SPAM_ECOLS, "dimension >= 2^32-1");
  const u64 base = ptr[r0], nnz_w = ptr[r1] - base;
  const size_t es = dtype_size(dtype);
  spam_dcsr* m = new spam_dcsr();
  m->dtype = dtype; m->rows = rows; m->cols = cols; m->nnz = nnz_w; m->owning = true; m->rows_sorted = -1; m->max_row_len = 0;
  m->ptr = nullptr; m->idx = nullptr; m->val = nullptr;
  u64* tmp = nullptr;
  int st = dev_alloc_t(h, &m->ptr, rows + 1);
  if (st == SPAM_OK) st = dev_alloc_t(h, &m->idx, nnz_w);
  if (st == SPAM_OK) st = dev_alloc(h, &m->val, nnz_w * es);
  if (st == SPAM_OK) st = dev_alloc_t(h, &tmp, nnz_w);
  cudaError_t e = cudaSuccess;
  if (st == SPAM_OK) {
    e = cudaMemcpyAsync(m->ptr + r0, ptr + r0, (r1 - r0 + 1) * sizeof(u64), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess && nnz_w) e = cudaMemcpyAsync(tmp, idx + base, nnz_w * sizeof(u64), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess && nnz_w)
      e = cudaMemcpyAsync(m->val, (const char*)val + base * es, nnz_w * es, cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) st = spam_fail(h, SPAM_ECUDA, "cudaMemcpyAsync H2D", e);
  }
  if (st == SPAM_OK) {
    const u64 nb = (rows + 1 + 255) / 256;
    k_window_ptr<<<(unsigned)(nb > 2048 ? 2048 : nb), 256, 0, h->stream>>>(m->ptr, rows + 1, r0, r1, base, nnz_w);
    count_launch(h);
    if ((e = cudaGetLastError()) != cudaSuccess) st = spam_fail(h, SPAM_ECUDA, "k_window_ptr", e);
  }
  if (st == SPAM_OK) st = narrow_u64_to_u32(h, tmp, m->idx, nnz_w);
  dev_free(h, tmp);
  if (st != SPAM_OK) { free_dcsr(h, m); return st; }
  h->stats.bytes_h2d += (r1 - r0 + 1) * 8 + nnz_w * (8 + es);
  *out = m;
  return SPAM_OK;
}

// Rows of B that A's columns reference: [*r0, *r1).  One small kernel + one sync.
static int referenced_rows(spam_handle* h, const spam_dcsr* a, u64 b_rows, u64* r0, u64* r1) {
  *r0 = 0; *r1 = 0;
  if (a->nnz == 0) return SPAM_OK;
  CK(cudaMemsetAsync(h->d_cnt, 0, sizeof(Counters), h->stream));
  const u64 nb = (a->nnz + 255) / 256;
  k_col_range<<<(unsigned)(nb > 1184 ? 1184 : nb), 256, 0, h->stream>>>(a->idx, a->nnz, h->d_cnt);
  count_launch(h);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const u64 cmin = (u32)~h->h_cnt->work_a, cmax = h->h_cnt->work_b;
  *r0 = cmin < b_rows ? cmin : b_rows;
  *r1 = cmax + 1 < b_rows ? cmax + 1 : b_rows;  // an out-of-range column is reported by the flop count (SPAM_EINDEX)
  if (*r1 < *r0) *r1 = *r0;
  return SPAM_OK;
}

int spam_dcsr_wrap(spam_handle* h, int dtype, uint64_t rows, uint64_t cols, uint64_t nnz, const void* d_ptr,
                   const void* d_idx, const void* d_val, spam_dcsr** out) {
  if (!h || !out || !d_ptr || !valid_dtype(dtype)) return spam_fail(h, SPAM_EINVAL, "bad argument");
  if (rows >= 0xFFFFFFFFull || cols >= 0xFFFFFFFFull) return spam_fail(h, SPAM_ECOLS, "dimension >= 2^32-1");
  spam_dcsr* m = new spam_dcsr();
  m->dtype = dtype; m->rows = rows; m->cols = cols; m->nnz = nnz; m->owning = false; m->rows_sorted = -1; m->max_row_len = 0;
  m->ptr = (u64*)d_ptr; m->idx = (u32*)d_idx; m->val = (void*)d_val;
  *out = m;
  return SPAM_OK;
}

int spam_dcsr_info(const spam_dcsr* m, int* dtype, uint64_t* rows, uint64_t* cols, uint64_t* nnz, void** d_ptr,
                   void** d_idx, void** d_val) {
  if (!m) return SPAM_EINVAL;
  if (dtype) *dtype = m->dtype;
  if (rows) *rows = m->rows;
  if (cols) *cols = m->cols;
  if (nnz) *nnz = m->nnz;
  if (d_ptr) *d_ptr = m->ptr;
  if (d_idx) *d_idx = m->idx;
  if (d_val) *d_val = m->val;
  return SPAM_OK;
}

int spam_dcsr_download(spam_handle* h, const spam_dcsr* m, uint64_t* ptr, uint64_t* idx, void* val) {
  if (!h || !m) return spam_fail(h, SPAM_EINVAL, "bad argument");
  CKS(set_device(h));
  DevGuard g(h);
  u64* wide = nullptr;
  if (idx && m->nnz) {
    CKS(g.alloc(&wide, m->nnz));
    CKS(widen_u32_to_u64(h, m->idx, wide, m->nnz));
  }
  if (ptr) CKS(dev_to_host(h, ptr, m->ptr, (m->rows + 1) * sizeof(u64)));
  if (val && m->nnz) CKS(dev_to_host(h, val, m->val, m->nnz * dtype_size(m->dtype)));
  if (idx && m->nnz) CKS(dev_to_host(h, idx, wide, m->nnz * sizeof(u64)));
  CK(cudaStreamSynchronize(h->stream));
  h->stats.bytes_d2h += (ptr ? (m->rows + 1) * 8 : 0) + (idx ? m->nnz * 8 : 0) + (val ? m->nnz * dtype_size(m->dtype) : 0);
  return SPAM_OK;
}

int spam_dcsr_free(spam_handle* h, spam_dcsr* m) {
  if (!h) return SPAM_EINVAL;
  if (!m) return SPAM_OK;
  CKS(set_device(h));
  free_dcsr(h, m);
  return SPAM_OK;
}

int spam_dcsr_slice_rows(spam_handle* h, const spam_dcsr* m, uint64_t r0, uint64_t r1, spam_dcsr** out) {
  if (!h || !m || !out || r0 > r1 || r1 > m->rows) return spam_fail(h, SPAM_EINVAL, "bad row range");
  *out = nullptr;
  CKS(set_device(h));
  u64 ends[2];
  CK(cudaMemcpyAsync(&ends[0], m->ptr + r0, sizeof(u64), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(&ends[1], m->ptr + r1, sizeof(u64), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const u64 nnz = ends[1] - ends[0], rows = r1 - r0;
  spam_dcsr* s = new spam_dcsr();
  s->dtype = m->dtype; s->rows = rows; s->cols = m->cols; s->nnz = nnz; s->owning = true; s->rows_sorted = m->rows_sorted; s->max_row_len = m->max_row_len;
  s->spread_sum = m->rows ? (u64)((double)m->spread_sum * (double)rows / (double)m->rows) : 0;
  s->ptr = nullptr; s->idx = nullptr; s->val = nullptr;
  int st = dev_alloc_t(h, &s->ptr, rows + 1);
  if (st == SPAM_OK) st = dev_alloc_t(h, &s->idx, nnz);
  if (st == SPAM_OK) st = dev_alloc(h, &s->val, nnz * dtype_size(m->dtype));
  if (st == SPAM_OK) {
    k_rebase_ptr<<<(unsigned)((rows + 1 + 255) / 256 > 2048 ? 2048 : (rows + 1 + 255) / 256), 256, 0, h->stream>>>(
        m->ptr + r0, s->ptr, rows + 1);
    count_launch(h);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(s->idx, m->idx + ends[0], nnz * sizeof(u32), cudaMemcpyDeviceToDevice, h->stream);
    if (e == cudaSuccess && nnz)
      e = cudaMemcpyAsync(s->val, (const char*)m->val + ends[0] * dtype_size(m->dtype), nnz * dtype_size(m->dtype),
                          cudaMemcpyDeviceToDevice, h->stream);
    if (e != cudaSuccess) st = spam_fail(h, SPAM_ECUDA, "slice copy", e);
  }
  if (st != SPAM_OK) { free_dcsr(h, s); return st; }
  *out = s;
  return SPAM_OK;
}

int spam_dcsr_select_rows(spam_handle* h, const spam_dcsr* m, const uint64_t* rows, uint64_t n, spam_dcsr** out) {
  if (!h || !m || !out || (n && !rows)) return spam_fail(h, SPAM_EINVAL, "bad argument");
  *out = nullptr;
  if (n >= 0xFFFFFFFFull) return spam_fail(h, SPAM_ECOLS, "more than 2^32-2 rows selected");
  CKS(set_device(h));
  DevGuard g(h);
  u64* d_rows = nullptr;
  u32* d_len = nullptr;
  u64* optr = nullptr;
  CKS(g.alloc(&d_rows, n));
  CKS(g.alloc(&d_len, n));
  CKS(g.alloc(&optr, n + 1));
  CK(cudaMemsetAsync(h->d_cnt, 0, sizeof(Counters), h->stream));
  if (n) {
    CK(cudaMemcpyAsync(d_rows, rows, n * sizeof(u64), cudaMemcpyHostToDevice, h->stream));
    k_select_len<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(m->ptr, d_rows, n, m->rows, d_len, h->d_cnt);
    count_launch(h);
    CK(cudaGetLastError());
  }
  CKS(scan_u32_to_u64(h, d_len, optr, n, &h->d_cnt->total_nnz));
  CK(cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->h_cnt->error & 2u) return spam_fail(h, SPAM_EINDEX, "a selected row index is >= rows");
  const u64 nnz = h->h_cnt->total_nnz;
  u32* oidx = nullptr;
  void* oval = nullptr;
  CKS(g.alloc(&oidx, nnz));
  CKS(g.alloc_bytes(&oval, nnz * dtype_size(m->dtype)));
  if (n && nnz) {
    const unsigned grid = (unsigned)((n * 32 + 255) / 256);
    if (dtype_size(m->dtype) == 4)
      k_select_copy<uint32_t><<<grid, 256, 0, h->stream>>>(m->ptr, m->idx, (const uint32_t*)m->val, d_rows, n, optr, oidx, (uint32_t*)oval);
    else
      k_select_copy<uint64_t><<<grid, 256, 0, h->stream>>>(m->ptr, m->idx, (const uint64_t*)m->val, d_rows, n, optr, oidx, (uint64_t*)oval);
    count_launch(h);
    CK(cudaGetLastError());
  }
  spam_dcsr* s = new spam_dcsr();
  s->dtype = m->dtype; s->rows = n; s->cols = m->cols; s->nnz = nnz; s->owning = true; s->rows_sorted = -1; s->max_row_len = 0;
  s->ptr = optr; s->idx = oidx; s->val = oval;
  g.release(optr); g.release(oidx); g.release(oval);
  *out = s;
  return SPAM_OK;
}

/* ---------------- transpose ---------------- */

int spam_dcsr_transpose(spam_handle* h, const spam_dcsr* m, spam_dcsr** out) {
  if (!h || !m || !out) return spam_fail(h, SPAM_EINVAL, "null argument");
  CKS(set_device(h));
  return transpose_dev(h, m, out);
}

int spam_csr_transpose(spam_handle* h, int dtype, uint64_t rows, uint64_t cols, const uint64_t* ptr, const uint64_t* idx,
                       const void* val, uint64_t* t_ptr, uint64_t* t_idx, void* t_val) {
  if (!h || !ptr || !t_ptr || !valid_dtype(dtype)) return spam_fail(h, SPAM_EINVAL, "bad argument");
  const u64 nnz = ptr[rows];
  if (nnz && (!idx || !val || !t_idx || !t_val)) return spam_fail(h, SPAM_EINVAL, "null buffer");
  CKS(set_device(h));
  spam_dcsr *a = nullptr, *t = nullptr;
  h->stats = spam_stats{};
  CKS(spam_csr_upload(h, dtype, rows, cols, nnz, ptr, idx, val, &a));
  const u64 h2d = h->stats.bytes_h2d;
  int st = transpose_dev(h, a, &t);  // resets stats
  if (st == SPAM_OK) {
    h->stats.bytes_h2d = h2d;
    st = spam_dcsr_download(h, t, t_ptr, t_idx, t_val);
  }
  free_dcsr(h, t);
  free_dcsr(h, a);
  return st;
}

/* ---------------- SpGEMM ---------------- */

int spam_spgemm_dev(spam_handle* h, const spam_dcsr* a, const spam_dcsr* b, spam_dcsr** c) {
  if (!h || !a || !b || !c) return spam_fail(h, SPAM_EINVAL, "null argument");
  *c = nullptr;
  CKS(set_device(h));
  CKS(spgemm_onepass_dev(h, a, b, c));
  if (*c) return SPAM_OK;
  SpgemmPending* p = nullptr;
  CKS(spgemm_symbolic_dev(h, a, b, &p));
  CKS(spgemm_numeric_dev(h, p, c));
  return SPAM_OK;
}

int spam_spgemm_dev_b2(spam_handle* h, const spam_dcsr* a, const spam_dcsr* b, int sorted, spam_dcsr** c) {
  if (!h || !a || !b || !c) return spam_fail(h, SPAM_EINVAL, "null argument");
  *c = nullptr;
  CKS(set_device(h));
  if (sorted) {
    CKS(spgemm_onepass_dev(h, a, b, c));
    if (*c) return SPAM_OK;
  }
  SpgemmPending* p = nullptr;
  CKS(spgemm_symbolic_dev(h, a, b, &p));
  CKS(spgemm_numeric_dev(h, p, c, sorted ? 1 : 0));
  return SPAM_OK;
}

int spam_spgemm_symbolic(spam_handle* h, int dtype, uint64_t a_rows, uint64_t a_cols, const uint64_t* a_ptr,
                         const uint64_t* a_idx, const void* a_val, uint64_t b_rows, uint64_t b_cols,
                         const uint64_t* b_ptr, const uint64_t* b_idx, const void* b_val, uint64_t* c_ptr,
                         uint64_t* c_nnz) {
  if (!h || !a_ptr || !b_ptr || !c_ptr || !c_nnz || !valid_dtype(dtype)) return spam_fail(h, SPAM_EINVAL, "bad argument");
  if (a_cols != b_rows) return spam_fail(h, SPAM_EDIM, "A.cols != B.rows");
  CKS(set_device(h));
  drop_spgemm_state(h);
  h->stats = spam_stats{};
  // a_ptr[a_rows] is nnz(A): offsets[rows] == indices.len() (invariant4, lib.rs:59-61)
  const u64 a_nnz = a_ptr[a_rows], b_nnz = b_ptr[b_rows];
  SpgemmHostState* s = new SpgemmHostState{nullptr, nullptr, nullptr};
  h->pending = s;
  const spam_stats keep = h->stats;
  int st = spam_csr_upload(h, dtype, a_rows, a_cols, a_nnz, a_ptr, a_idx, a_val, &s->a);
  const bool alias = (b_ptr == a_ptr && b_idx == a_idx && b_val == a_val && b_rows == a_rows && b_cols == a_cols);
  if (st == SPAM_OK) {
    if (alias) {
      s->b = s->a;
    } else {
      // A touches rows [r0, r1) of B only: when that is a narrow band (a row block of a banded matrix), upload
      // just the band over PCIe
      u64 r0 = 0, r1 = b_rows;
      st = referenced_rows(h, s->a, b_rows, &r0, &r1);
      if (st == SPAM_OK) {
        if ((r1 - r0) < b_rows - b_rows / 8) st = upload_csr_window(h, dtype, b_rows, b_cols, b_ptr, b_idx, b_val, r0, r1, &s->b);
        else st = spam_csr_upload(h, dtype, b_rows, b_cols, b_nnz, b_ptr, b_idx, b_val, &s->b);
      }
    }
  }
  const u64 h2d = h->stats.bytes_h2d;
  (void)keep;
  if (st == SPAM_OK) st = spgemm_symbolic_dev(h, s->a, s->b, &s->pend);  // resets stats
  if (st != SPAM_OK) { drop_spgemm_state(h); return st; }
  h->stats.bytes_h2d = h2d;
  st = dev_to_host(h, c_ptr, spgemm_pending_cptr(s->pend), (a_rows + 1) * sizeof(u64));
  cudaError_t e = st == SPAM_OK ? cudaStreamSynchronize(h->stream) : cudaSuccess;
  if (st != SPAM_OK || e != cudaSuccess) { drop_spgemm_state(h); return st != SPAM_OK ? st : spam_fail(h, SPAM_ECUDA, "D2H row_ptr", e); }
  h->stats.bytes_d2h += (a_rows + 1) * 8;
  *c_nnz = spgemm_pending_nnz(s->pend);
  return SPAM_OK;
}

int spam_spgemm_numeric(spam_handle* h, uint64_t* c_idx, void* c_val, int sorted) {
  if (!h) return SPAM_EINVAL;
  SpgemmHostState* s = static_cast<SpgemmHostState*>(h->pending);
  if (!s || !s->pend) return spam_fail(h, SPAM_ESTATE, "spam_spgemm_numeric without spam_spgemm_symbolic");
  const u64 nnz = spgemm_pending_nnz(s->pend);
  if (nnz && (!c_idx || !c_val)) return spam_fail(h, SPAM_EINVAL, "null output buffer");
  CKS(set_device(h));
  spam_dcsr* c = nullptr;
  SpgemmPending* p = s->pend;
  s->pend = nullptr;  // consumed by numeric whatever the outcome
  int st = spgemm_numeric_dev(h, p, &c, sorted ? 1 : 0);
  if (st == SPAM_OK) st = spam_dcsr_download(h, c, nullptr, c_idx, c_val);
  if (st == SPAM_OK) finish_timing(h);
  free_dcsr(h, c);
  drop_spgemm_state(h);
  return st;
}

/* ---------------- SpMV ---------------- */

int spam_spmv_dev(spam_handle* h, const spam_dcsr* a, const void* d_x, void* d_y) {
  if (!h || !a || !d_x || !d_y) return spam_fail(h, SPAM_EINVAL, "null argument");
  CKS(set_device(h));
  return spmv_dev(h, a, d_x, d_y);
}

int spam_spmv(spam_handle* h, int dtype, uint64_t a_rows, uint64_t a_cols, const uint64_t* a_ptr,
              const uint64_t* a_idx, const void* a_val, const void* x, void* y) {
  if (!h || !a_ptr || !x || !y || !valid_dtype(dtype)) return spam_fail(h, SPAM_EINVAL, "bad argument");
  CKS(set_device(h));
  spam_dcsr* a = nullptr;
  CKS(spam_csr_upload(h, dtype, a_rows, a_cols, a_ptr[a_rows], a_ptr, a_idx, a_val, &a));
  void *dx = nullptr, *dy = nullptr;
  const size_t es = dtype_size(dtype);
  int st = dev_alloc(h, &dx, a_cols * es);
  if (st == SPAM_OK) st = dev_alloc(h, &dy, a_rows * es);
  cudaError_t e = cudaSuccess;
  if (st == SPAM_OK) e = cudaMemcpyAsync(dx, x, a_cols * es, cudaMemcpyHostToDevice, h->stream);
  if (st == SPAM_OK && e == cudaSuccess) st = spmv_dev(h, a, dx, dy);
  if (st == SPAM_OK && e == cudaSuccess) e = cudaMemcpyAsync(y, dy, a_rows * es, cudaMemcpyDeviceToHost, h->stream);
  if (st == SPAM_OK && e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (st == SPAM_OK && e != cudaSuccess) st = spam_fail(h, SPAM_ECUDA, "spmv copies", e);
  dev_free(h, dx); dev_free(h, dy);
  free_dcsr(h, a);
  return st;
}

/* ---------------- DOK -> CSR ---------------- */

int spam_dok_to_csr_dev(spam_handle* h, int dtype, uint64_t rows, uint64_t cols, uint64_t n, const void* d_r,
                        const void* d_c, const void* d_v, spam_dcsr** out) {
  if (!h || !out || !valid_dtype(dtype) || (n && (!d_r || !d_c || !d_v))) return spam_fail(h, SPAM_EINVAL, "bad argument");
  *out = nullptr;
  if (rows >= 0xFFFFFFFFull || cols >= 0xFFFFFFFFull) return spam_fail(h, SPAM_ECOLS, "dimension >= 2^32-1");
  CKS(set_device(h));
  return dok_to_csr_dev(h, dtype, rows, cols, n, (const u64*)d_r, (const u64*)d_c, d_v, out);
}

int spam_dok_to_csr(spam_handle* h, int dtype, uint64_t rows, uint64_t cols, uint64_t n, const uint64_t* tr,
                    const uint64_t* tc, const void* tv, uint64_t* c_ptr, uint64_t* c_nnz) {
  if (!h || !c_ptr || !c_nnz || !valid_dtype(dtype) || (n && (!tr || !tc || !tv))) return spam_fail(h, SPAM_EINVAL, "bad argument");
  CKS(set_device(h));
  drop_dok_state(h);
  u64 *dr = nullptr, *dc = nullptr;
  void* dv = nullptr;
  const size_t es = dtype_size(dtype);
  int st = dev_alloc_t(h, &dr, n);
  if (st == SPAM_OK) st = dev_alloc_t(h, &dc, n);
  if (st == SPAM_OK) st = dev_alloc(h, &dv, n * es);
  cudaError_t e = cudaSuccess;
  if (st == SPAM_OK && n) {
    e = cudaMemcpyAsync(dr, tr, n * 8, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dc, tc, n * 8, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dv, tv, n * es, cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) st = spam_fail(h, SPAM_ECUDA, "triplet H2D", e);
  }
  spam_dcsr* c = nullptr;
  if (st == SPAM_OK) st = spam_dok_to_csr_dev(h, dtype, rows, cols, n, dr, dc, dv, &c);
  dev_free(h, dr); dev_free(h, dc); dev_free(h, dv);
  if (st != SPAM_OK) return st;
  e = cudaMemcpyAsync(c_ptr, c->ptr, (rows + 1) * 8, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) { free_dcsr(h, c); return spam_fail(h, SPAM_ECUDA, "D2H row_ptr", e); }
  *c_nnz = c->nnz;
  h->dok_pending = new DokPending{c};
  return SPAM_OK;
}

int spam_dok_to_csr_fetch(spam_handle* h, uint64_t* c_idx, void* c_val) {
  if (!h) return SPAM_EINVAL;
  if (!h->dok_pending) return spam_fail(h, SPAM_ESTATE, "spam_dok_to_csr_fetch without spam_dok_to_csr");
  CKS(set_device(h));
  spam_dcsr* c = h->dok_pending->c;
  int st = SPAM_OK;
  if (c->nnz && (!c_idx || !c_val)) st = spam_fail(h, SPAM_EINVAL, "null output buffer");
  if (st == SPAM_OK) st = spam_dcsr_download(h, c, nullptr, c_idx, c_val);
  drop_dok_state(h);
  return st;
}

/* ---------------- elementwise add / sub ---------------- */

int spam_dcsr_ewise(spam_handle* h, int op, const spam_dcsr* a, const spam_dcsr* b, spam_dcsr** out) {
  if (!h || !a || !b || !out) return spam_fail(h, SPAM_EINVAL, "null argument");
  CKS(set_device(h));
  return ewise_dev(h, op, a, b, out);
}

// host path, two phases like DOK -> CSR (the result size is known only after the count): phase 1 writes
// c_ptr and *c_nnz and parks the result on the device, spam_csr_ewise_fetch downloads it
int spam_csr_ewise(spam_handle* h, int op, int dtype, uint64_t rows, uint64_t cols, const uint64_t* a_ptr,
                   const uint64_t* a_idx, const void* a_val, const uint64_t* b_ptr, const uint64_t* b_idx,
                   const void* b_val, uint64_t* c_ptr, uint64_t* c_nnz) {
  if (!h || !a_ptr || !b_ptr || !c_ptr || !c_nnz || !valid_dtype(dtype)) return spam_fail(h, SPAM_EINVAL, "bad argument");
  CKS(set_device(h));
  drop_dok_state(h);
  spam_dcsr *a = nullptr, *b = nullptr, *c = nullptr;
  h->stats = spam_stats{};
  CKS(spam_csr_upload(h, dtype, rows, cols, a_ptr[rows], a_ptr, a_idx, a_val, &a));
  const bool alias = (b_ptr == a_ptr && b_idx == a_idx && b_val == a_val);
  int st = SPAM_OK;
  if (alias) b = a; else st = spam_csr_upload(h, dtype, rows, cols, b_ptr[rows], b_ptr, b_idx, b_val, &b);
  const u64 h2d = h->stats.bytes_h2d;
  if (st == SPAM_OK) st = ewise_dev(h, op, a, b, &c);  // resets stats
  if (st == SPAM_OK) {
    h->stats.bytes_h2d = h2d;
    st = spam_dcsr_download(h, c, c_ptr, nullptr, nullptr);
  }
  if (b != a) free_dcsr(h, b);
  free_dcsr(h, a);
  if (st != SPAM_OK) { free_dcsr(h, c); return st; }
  *c_nnz = c->nnz;
  h->dok_pending = new DokPending{c};
  return SPAM_OK;
}

int spam_csr_ewise_fetch(spam_handle* h, uint64_t* c_idx, void* c_val) { return spam_dok_to_csr_fetch(h, c_idx, c_val); }

/* ---------------- multi-GPU helpers ---------------- */

static int rows_to_parts_impl(spam_handle* h, const spam_dcsr* a, const spam_dcsr* b, uint32_t parts,
                              uint64_t* row_starts, uint64_t* total_flops, bool by_cost) {
  if (!h || !a || !b || !row_starts || parts == 0) return spam_fail(h, SPAM_EINVAL, "bad argument");
  if (a->cols != b->rows) return spam_fail(h, SPAM_EDIM, "A.cols != B.rows");
  CKS(set_device(h));
  const u64 m = a->rows;
  CKS(ensure_matrix_stats(h, a));
  if (b != a) CKS(ensure_matrix_stats(h, b));
  DevGuard g(h);
  u32* flop = nullptr;
  u64 *ps = nullptr, *d_out = nullptr;
  CKS(g.alloc(&flop, m));
  CKS(g.alloc(&ps, m + 1));
  CKS(g.alloc(&d_out, (u64)parts + 1));
  CK(cudaMemsetAsync(h->d_cnt, 0, sizeof(Counters), h->stream));
  CKS(flop_count_dev(h, a, b, flop, false, 0));
  if (by_cost && m) {
    k_flop_to_cost<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(flop, a->ptr, m);
    count_launch(h);
    CK(cudaGetLastError());
  }
  CKS(scan_u32_to_u64(h, flop, ps, m, nullptr));
  k_partition_points<<<(parts + 1 + 127) / 128, 128, 0, h->stream>>>(ps, m, parts, d_out);
  count_launch(h);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(row_starts, d_out, ((u64)parts + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->h_cnt->error & 1u) return spam_fail(h, SPAM_EINDEX, "a column index of A is >= rows(B)");
  if (total_flops) *total_flops = h->h_cnt->total_flops;
  return SPAM_OK;
}

int spam_rows_to_parts(spam_handle* h, const spam_dcsr* a, const spam_dcsr* b, uint32_t parts, uint64_t* row_starts,
                       uint64_t* total_flops) {
  return rows_to_parts_impl(h, a, b, parts, row_starts, total_flops, false);
}

int spam_rows_to_parts_cost(spam_handle* h, const spam_dcsr* a, const spam_dcsr* b, uint32_t parts,
                            uint64_t* row_starts, uint64_t* total_flops) {
  return rows_to_parts_impl(h, a, b, parts, row_starts, total_flops, true);
}

int spam_offset_u64(spam_handle* h, void* d_ptr_u64, uint64_t n, uint64_t offset) {
  if (!h || (!d_ptr_u64 && n)) return spam_fail(h, SPAM_EINVAL, "bad argument");
  CKS(set_device(h));
  return add_offset_u64(h, (u64*)d_ptr_u64, n, offset);
}

}  // extern "C"
