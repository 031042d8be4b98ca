// common.cuh — shared types for libspam_cuda (sm_100a).  Device layout of a CSR matrix:
//   row_ptr : u64[rows+1]   (nnz(C) and flop prefix sums exceed 2^32 on the big configs)
//   col_idx : u32[nnz]      (the reference truncates keys to u32: spam_csr/src/mul_hash.rs:92,157;
//                            u32::MAX is the empty-slot sentinel: linprobe/src/set.rs:45,110)
//   val     : T[nnz]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "../../include/spam_cuda.h"

typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef unsigned long long ull;

constexpr u32 EMPTY_KEY = 0xFFFFFFFFu;  // linprobe/src/set.rs:45
constexpr u32 HASH_SCAL = 107u;         // linprobe/src/lib.rs:13
constexpr u32 MIN_TABLE = 16u;          // linprobe/src/lib.rs:14

constexpr int NBINS = 16;

// ---- row bins ---------------------------------------------------------------------------
// One hash bin per power of two, so a row's shared-memory table is at most 2x what linprobe's rule
// max(16, 2*npow2(n)) gives it: the team kernels are occupancy-bound (profiles/r01_rmat18_v1_rowhash.txt),
// and a bin that lumps z = 513..2048 together makes every row pay for a 4096-slot table.
//   0        tiny   f <= 32 / (z <= 16 and f <= 128)   thread-per-row private table (k_*_tiny)
//   1..8     hash   symbolic bin b: f <= 64 << b  (128 .. 16384), table 128 << b keys
//                   numeric  bin b: z <= 32 << b  (64 .. 8192),   table 64 << b (key, value) slots
//                   one warp per row while the table is <= 12 KB, then teams of 4 .. 32 warps (k_*_row)
//   9        heavy  above: global-memory table, persistent 1024-thread blocks (k_*_heavy)
//   10       merge  B sorted, len(A row) <= MERGE_K, f <= 128: k-way merge of sorted runs (merge.cuh)
//   11..15   esc    numeric only: rows that do not compress (z > 256 and 2 z >= f) are bucket-sorted instead of
//                   hashed (esc.cuh): f <= 1024 / 2048 / 4096 / 8192 one block per row, 15 = longer rows in
//                   column ranges
constexpr int NHASH = 8, HEAVY_BIN = 9, MERGE_BIN = 10, ESC_BIN0 = 11, ESC_HEAVY_BIN = 15;
constexpr u32 ESC_ZMIN = 256;
// `mode` bits of the binning functions: which optional bins the product may use
constexpr int MODE_MERGE = 1, MODE_ESC = 2, MODE_ESC_HEAVY = 4;
constexpr u32 SYM_TINY_MAX = 32, NUM_TINY_MAX = 16, NUM_TINY_FLOP_MAX = 128;
constexpr u32 SYM_HASH_FMAX = 64u << NHASH, NUM_HASH_ZMAX = 32u << NHASH;
constexpr u32 MERGE_K = 8, MERGE_FLOP_MAX = 128;

__host__ __device__ __forceinline__ int ceil_log2_u32(u32 x) {  // x >= 1
#ifdef __CUDA_ARCH__
  return x <= 1 ? 0 : 32 - __clz(x - 1);
#else
  int l = 0;
  while ((1ull << l) < x) ++l;
  return l;
#endif
}
__host__ __device__ __forceinline__ int sym_bin_of(u32 f, u32 alen = 0xFFFFFFFFu, bool merge_ok = false) {
  if (merge_ok && alen <= MERGE_K && f <= MERGE_FLOP_MAX) return MERGE_BIN;
  if (f <= SYM_TINY_MAX) return 0;
  if (f > SYM_HASH_FMAX) return HEAVY_BIN;
  const int b = ceil_log2_u32(f) - 6;  // smallest b with f <= 64 << b
  return b < 1 ? 1 : b;
}
__host__ __device__ __forceinline__ int num_bin_of(u32 z, u32 f, u32 alen = 0xFFFFFFFFu, int mode = 0) {
  if ((mode & MODE_MERGE) && alen <= MERGE_K && f <= MERGE_FLOP_MAX) return MERGE_BIN;
  if (z <= NUM_TINY_MAX && f <= NUM_TINY_FLOP_MAX) return 0;
  if ((mode & MODE_ESC) && z > ESC_ZMIN && 2ull * z >= f && f != 0xFFFFFFFFu && alen <= f) {  // f saturates at u32::MAX
    if (f <= 1024) return ESC_BIN0;
    if (f <= 2048) return ESC_BIN0 + 1;
    if (f <= 4096) return ESC_BIN0 + 2;
    if (f <= 8192) return ESC_BIN0 + 3;
    if (mode & MODE_ESC_HEAVY) return ESC_HEAVY_BIN;
  }
  if (z > NUM_HASH_ZMAX) return HEAVY_BIN;
  const int b = ceil_log2_u32(z < 1 ? 1 : z) - 5;  // smallest b with z <= 32 << b
  return b < 1 ? 1 : b;
}

// ---- device counters block (one per handle, zeroed per product) ---------------------------
struct Counters {
  ull total_flops;
  ull total_nnz;
  u32 sym_bins[NBINS];
  u32 num_bins[NBINS];
  u32 sym_cursor[NBINS];
  u32 num_cursor[NBINS];
  u32 max_flop;
  u32 max_nnz;
  u32 error;      // bit0: column index of A >= rows(B); bit1: triplet index out of range
  u32 work_a;     // dynamic work counters for the persistent heavy-row kernels
  u32 work_b;
  u32 work_c;     // row queue of the global-table kernel when it runs the bucket-sort bins' fallback list
  u32 max_alen;   // longest A row among the rows of the merge bin (picks the head-count template)
  u32 unsorted;   // set by k_rows_sorted when some row is not strictly increasing
  u32 max_rowlen; // longest row seen by k_rows_sorted
  u32 invalid;    // set by k_rows_sorted: bit0 row_ptr not monotone / out of range, bit1 a column index >= cols
  // how often the rarely taken code paths ran during the last product (spam_stats.fallbacks, tests assert them)
  u32 fb_warp_bitonic;   // NW = 1 rows sorted by the shared-memory bitonic network (columns too wide to pack)
  u32 fb_team_bitonic;   // team rows whose bucket drain overflowed (compaction + block bitonic network)
  u32 fb_heavy_bitonic;  // global-table rows whose bucket drain overflowed (global-memory bitonic network)
  u32 fb_esc;            // rows the bucket-sort (ESC) bins handed back to the global-table kernel
  u32 fb_list_n;         // number of row ids in the ESC fallback list
  u32 bk_flags;          // bucket.cuh: bit0 a bucket overflowed, bit1 a segment is too long for the all-pairs loops
  ull spread_sum;        // k_rows_sorted: sum over the rows of <= 32 entries of (largest - smallest column index)
};

struct BinBase { u32 v[NBINS]; };

// ---- TMA -------------------------------------------------------------------------------------
// 1-D bulk copies global -> shared through the TMA unit (SASS UBLKCP), completion on an mbarrier.
static __device__ __forceinline__ u32 msm_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(msm_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
static __device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(msm_addr(bar)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void bulk_g2s(void* dst, const void* src, u32 bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(msm_addr(dst)),
               "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(msm_addr(bar))
               : "memory");
}
static __device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
  u32 ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(msm_addr(bar)), "r"(parity)
                 : "memory");
  } while (!ok);
}

// ---- element arithmetic --------------------------------------------------------------------
// mul_hash.rs:154-161: product `t * t1`, then `*t += t1`: two roundings, never fused.
template <class T> struct Num;
template <> struct Num<float> {
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ void atomic_add(float* p, float v) { atomicAdd(p, v); }
  static __device__ __forceinline__ float zero() { return 0.f; }
};
template <> struct Num<double> {
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ void atomic_add(double* p, double v) { atomicAdd(p, v); }
  static __device__ __forceinline__ double zero() { return 0.0; }
};
// integers wrap (release-mode Rust / Wrapping<iN>): do the arithmetic unsigned
template <> struct Num<int32_t> {
  static __device__ __forceinline__ int32_t mul(int32_t a, int32_t b) { return (int32_t)((u32)a * (u32)b); }
  static __device__ __forceinline__ int32_t add(int32_t a, int32_t b) { return (int32_t)((u32)a + (u32)b); }
  static __device__ __forceinline__ int32_t sub(int32_t a, int32_t b) { return (int32_t)((u32)a - (u32)b); }
  static __device__ __forceinline__ void atomic_add(int32_t* p, int32_t v) { atomicAdd((u32*)p, (u32)v); }
  static __device__ __forceinline__ int32_t zero() { return 0; }
};
template <> struct Num<int64_t> {
  static __device__ __forceinline__ int64_t mul(int64_t a, int64_t b) { return (int64_t)((u64)a * (u64)b); }
  static __device__ __forceinline__ int64_t add(int64_t a, int64_t b) { return (int64_t)((u64)a + (u64)b); }
  static __device__ __forceinline__ int64_t sub(int64_t a, int64_t b) { return (int64_t)((u64)a - (u64)b); }
  static __device__ __forceinline__ void atomic_add(int64_t* p, int64_t v) { atomicAdd((ull*)p, (ull)v); }
  static __device__ __forceinline__ int64_t zero() { return 0; }
};

// next power of two, npow2(0) = npow2(1) = 1  (usize::next_power_of_two)
__host__ __device__ __forceinline__ u32 npow2_u32(u32 x) {
  if (x <= 1) return 1;
#ifdef __CUDA_ARCH__
  return 1u << (32 - __clz(x - 1));
#else
  u32 p = 1;
  while (p < x) p <<= 1;
  return p;
#endif
}
__host__ __device__ __forceinline__ u64 npow2_u64(u64 x) {
  u64 p = 1;
  while (p < x) p <<= 1;
  return p;
}
// linprobe table size for `capacity` items: max(16, 2*npow2(capacity))  (map.rs:33-38)
__host__ __device__ __forceinline__ u32 table_size_u32(u32 capacity) {
  u32 t = 2u * npow2_u32(capacity);
  return t < MIN_TABLE ? MIN_TABLE : t;
}
// linprobe/src/lib.rs:29-31 + set.rs:131 / map.rs:68: (key * 107) & (len - 1)
__device__ __forceinline__ u32 slot_of(u32 key, u32 mask) { return (key * HASH_SCAL) & mask; }

// ---- matrices and the handle -----------------------------------------------------------------
struct spam_dcsr {
  int dtype;
  u64 rows, cols, nnz;
  u64* ptr;   // device
  u32* idx;   // device
  void* val;  // device
  bool owning;
  int rows_sorted;  // cached property: -1 unknown, 0 no, 1 every row strictly increasing (IS_SORTED)
  u64 max_row_len;  // cached with rows_sorted: longest row (picks DIRECT vs FLAT product enumeration)
  int invalid;      // cached with rows_sorted: 0 = row_ptr monotone from 0 to nnz and every column < cols
                    // (invariants 3, 4, 5, 7 of spam_csr/src/lib.rs:47-81); bit0 row_ptr, bit1 column range
  u64 spread_sum = 0;  // cached with rows_sorted: sum over the short rows of (largest - smallest column); as a left operand,
                       // spread_sum / rows * mean(B row) says whether the B rows a block touches can stay in L1/L2
  spam_dcsr* sorted_copy;  // rows_sorted == 0: the same matrix with every row in column order, made on first need
                           // (sorted_rows_of) and owned by this object
};

struct HostStage {  // hostio.cu: ring of pinned slots for copies from / to pageable caller memory
  unsigned char* buf;
  cudaEvent_t ev[4];
  int nthreads;
};

struct SpgemmPending;  // state between the two host phases
struct DokPending;

struct spam_handle {
  int device;
  cudaStream_t own_stream, stream;
  bool timing;
  std::string err;
  spam_stats stats;
  Counters* d_cnt;   // device
  Counters* h_cnt;   // pinned host mirror
  int num_sms;
  int max_smem_optin;
  void* pending;             // SpgemmHostState* (api.cu) between the two host phases
  DokPending* dok_pending;
  // phase events, two sets: a product records into one set while the previous product's set is still
  // un-harvested; a set is harvested (elapsed times -> stats + running totals) after the next host sync
  // that covers it, so per-phase timing adds no synchronisation of its own
  cudaEvent_t evs[2][6];
  cudaEvent_t* ev;       // = evs[ev_cur]
  int ev_cur;
  bool ev_pending[2];
  double acc_ms[5];      // flop, symbolic, scan, numeric, total over the harvested products
  u64 acc_n;
  // side lanes: the per-bin kernels of one product touch disjoint rows, so they are spread over the main
  // stream and NLANES-1 internal streams (fork/join by events) and the tail of one bin overlaps the next
  cudaStream_t lane[3];
  cudaEvent_t lane_ev[4];  // [0] fork, [1..3] join
  u64* scan_ws;      // look-back scan tile states + tile counter (grow-only, stream-ordered reuse)
  u64 scan_ws_cap;   // in u64 words
  cudaMemPool_t pool;  // private stream-ordered pool: freed blocks stay with this handle, not with the process
  bool use_lanes;      // SPAM_LANES=0 in the environment at create time keeps every bin on the main stream
  int use_esc;         // SPAM_ESC at create time: 0 = hash bins only (default: measured faster on B200), 1 = bucket-sort
                       // bins for non-compressing rows up to 8192 products, 2 = also the column-range kernel for longer rows
  bool sort_b;         // SPAM_SORT_B=0 at create time: never multiply by a sorted copy of an unsorted B (tests of the hash bins)
  int merge_win;       // SPAM_MERGE_WIN at create time: bit 0 numeric, bit 1 symbolic merge kernels stage the block's B window in shared memory with cp.async.bulk (default 0: measured slower)
  int merge_pf;        // SPAM_MERGE_PF at create time: low two bits = PF of k_num_merge (0..2), bit 2 = one-ahead columns in k_flop_sym_merge; -1 (default): 6 when A's rows scatter over B, else 0
  bool spmv_tma;       // SPAM_SPMV_TMA=1 at create time: the persistent TMA-pipelined SpMV kernel (spmv.cu) instead of k_spmv_stream
  bool onepass;        // SPAM_ONEPASS=1 at create time: device-resident products whose rows are all merge rows by the cached statistics run as one kernel (spgemm_onepass_dev; measured slower, off by default)
  int merge_persist;   // SPAM_MERGE_PERSIST=N at create time (experiment, measured slower): k_num_merge as a persistent grid of N blocks per SM
  bool dok_bucket;     // SPAM_DOK_BUCKET=0 at create time: DOK -> CSR and transpose without the bucket path (bucket.cuh)
  bool ewise_tma;      // SPAM_EWISE_TMA=0 at create time: elementwise fill without the TMA-staged spans (k_ewise_fill)
  int l2_persist;      // SPAM_L2_PERSIST at create time (experiment, spgemm.cu): 1 = B's col_idx, 2 = B's values persisting in L2
  size_t l2_persist_max, l2_window_max;
  HostStage* stage;    // created on the first copy that involves a pageable host buffer
  struct CommState* comm;  // comm.cu: NCCL communicator + peer-mapped gather buffers (spam_comm_init), or null
};

static inline size_t dtype_size(int dt) { return (dt == SPAM_F32 || dt == SPAM_I32) ? 4 : 8; }

// error plumbing: record message, return status
int spam_fail(spam_handle* h, int status, const char* what, cudaError_t ce = cudaSuccess);
#define CK(call)                                                             \
  do {                                                                       \
    cudaError_t _e = (call);                                                 \
    if (_e != cudaSuccess) return spam_fail(h, SPAM_ECUDA, #call, _e);       \
  } while (0)
#define CKS(call)                                \
  do {                                           \
    int _s = (call);                             \
    if (_s != SPAM_OK) return _s;                \
  } while (0)

// stream-ordered allocation on the handle's stream (pool keeps freed blocks: no cudaMalloc
// in steady state, the analogue of the reference reusing its allocator's free lists)
int dev_alloc(spam_handle* h, void** p, size_t bytes);
int dev_free(spam_handle* h, void* p);
template <class T>
static inline int dev_alloc_t(spam_handle* h, T** p, size_t n) { return dev_alloc(h, (void**)p, n * sizeof(T)); }

// Frees every buffer it allocated when it goes out of scope, unless release()d: early error returns cannot
// leak stream-ordered allocations.
struct DevGuard {
  spam_handle* h;
  std::vector<void*> owned;
  explicit DevGuard(spam_handle* hh) : h(hh) {}
  DevGuard(const DevGuard&) = delete;
  DevGuard& operator=(const DevGuard&) = delete;
  ~DevGuard() { for (void* p : owned) dev_free(h, p); }
  template <class T>
  int alloc(T** p, size_t n) {
    const int st = dev_alloc(h, (void**)p, n * sizeof(T));
    if (st == SPAM_OK) owned.push_back((void*)*p);
    return st;
  }
  int alloc_bytes(void** p, size_t bytes) {
    const int st = dev_alloc(h, p, bytes);
    if (st == SPAM_OK) owned.push_back(*p);
    return st;
  }
  void release(void* p) {  // ownership moves to the caller
    for (auto& q : owned) if (q == p) { q = owned.back(); owned.pop_back(); return; }
  }
};

static inline void count_launch(spam_handle* h, u64 n = 1) { h->stats.kernel_launches += n; }

// fork: the side lanes wait for everything queued on the main stream so far; join: the main stream waits
// for everything queued on the side lanes.
static inline cudaError_t lanes_fork(spam_handle* h) {
  cudaError_t e = cudaEventRecord(h->lane_ev[0], h->stream);
  for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaStreamWaitEvent(h->lane[i], h->lane_ev[0], 0);
  return e;
}
static inline cudaError_t lanes_join(spam_handle* h) {
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 3 && e == cudaSuccess; ++i) {
    e = cudaEventRecord(h->lane_ev[1 + i], h->lane[i]);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(h->stream, h->lane_ev[1 + i], 0);
  }
  return e;
}
// stream of row bin `bin`: the team bins go to the side lanes, the rest stays on the main stream.  The
// global-table (heavy) kernels always run alone, after the join: overlapping them with the team kernels cost
// 10% of the whole product on R-MAT 22 (their L2 atomics and the teams' B gathers fight over L2).
static inline cudaStream_t lane_of(spam_handle* h, int bin) {
  if (!h->use_lanes) return h->stream;
  switch (bin) {
    case 8: case 5: case 14: case 11: return h->lane[0];
    case 7: case 4: case 13: return h->lane[1];
    case 6: case 12: case 15: return h->lane[2];
    default: return h->stream;
  }
}

// api.cu : phase-event bookkeeping
void timing_begin_product(spam_handle* h);          // switch to the other event set
void timing_harvest(spam_handle* h, int set);       // events of `set` are known complete
void finish_timing(spam_handle* h);                 // wait for the current set, harvest it

// ---- entry points implemented across the .cu files ---------------------------------------------
// scan.cu : exclusive scan of u32 counts into u64 offsets (out has n+1 entries), decoupled look-back
int lookback_workspace(spam_handle* h, u64 tiles, u64** state, u32** tile_counter);  // scan.cu
int scan_u32_to_u64(spam_handle* h, const u32* in, u64* out, u64 n, ull* d_total, u32* d_max = nullptr);
// spgemm.cu
int spgemm_symbolic_dev(spam_handle* h, const spam_dcsr* a, const spam_dcsr* b, SpgemmPending** out);
int spgemm_numeric_dev(spam_handle* h, SpgemmPending* p, spam_dcsr** c, int sorted = 1);
// slotorder.cu : rows of a sorted product permuted into the reference's B2 = false (linear-probe slot) order
int slot_order_dev(spam_handle* h, const spam_dcsr* a, const spam_dcsr* b_original, spam_dcsr* c);
int spgemm_onepass_dev(spam_handle* h, const spam_dcsr* a, const spam_dcsr* b, spam_dcsr** cout);  // *cout null: not applicable
int spgemm_numeric_into(spam_handle* h, SpgemmPending* p, const u64* c_ptr, u32* c_idx, void* c_val);
void spgemm_pending_free(spam_handle* h, SpgemmPending* p);
u64 spgemm_pending_nnz(const SpgemmPending* p);
const u64* spgemm_pending_cptr(const SpgemmPending* p);
int flop_count_dev(spam_handle* h, const spam_dcsr* a, const spam_dcsr* b, u32* d_flop, bool do_bins, int mode);
// cached per-matrix properties (rows sorted? longest row), one pass over col_idx on first use
int ensure_matrix_stats(spam_handle* h, const spam_dcsr* m);
// spmv.cu
int spmv_dev(spam_handle* h, const spam_dcsr* a, const void* d_x, void* d_y);
// dok.cu
int dok_to_csr_dev(spam_handle* h, int dtype, u64 rows, u64 cols, u64 n, const u64* d_r, const u64* d_c,
                   const void* d_v, spam_dcsr** out);
int transpose_dev(spam_handle* h, const spam_dcsr* a, spam_dcsr** out);
int dok_partition_dev(spam_handle* h, int dtype, u64 rows, u64 cols, u64 rows_per, int world, u64 n, const u64* d_r,
                      const u64* d_c, const void* d_v, u64* o_r, u64* o_c, void* o_v, u64* counts_host);
// m itself when its rows are sorted by column, else a cached copy with sorted rows (two stable transposes)
int sorted_rows_of(spam_handle* h, const spam_dcsr* m, const spam_dcsr** view);
void free_dcsr_tree(spam_handle* h, spam_dcsr* m);
// ewise.cu : C = A + B (op 0) / A - B (op 1)
int ewise_dev(spam_handle* h, int op, const spam_dcsr* a, const spam_dcsr* b, spam_dcsr** out);
// hostio.cu : host <-> device copies; pageable host memory goes through the handle's pinned staging ring
int host_to_dev(spam_handle* h, void* d_dst, const void* h_src, size_t bytes);
int dev_to_host(spam_handle* h, void* h_dst, const void* d_src, size_t bytes);
void host_stage_free(spam_handle* h);
// convert.cu : index width conversion at the host boundary
int narrow_u64_to_u32(spam_handle* h, const u64* in, u32* out, u64 n);
int widen_u32_to_u64(spam_handle* h, const u32* in, u64* out, u64 n);
int add_offset_u64(spam_handle* h, u64* p, u64 n, u64 off, const u64* src = nullptr);
