// ewise.cu — elementwise C = A + B / A - B on the device.
//
// Replaces `impl Add / Sub for CsrMatrix` -> `apply_elementwise` (spam_csr/src/lib.rs:83-149, 276-290).
// IS_SORTED branch of the reference: per row, a merge-join of the two sorted column lists; a column in both
// rows gives f(t1, t2), only in the left row f(t, 0), only in the right row f(0, t) (lib.rs:112-116).  Nothing
// is filtered: cancellation zeros and explicit zeros stay (there is no is_zero test in apply_elementwise).
// The unsorted branch (lib.rs:119-137) collects the left row into a std HashMap and folds the right row in:
// same columns, f(t1, t2) and f(0, t) as above, but an entry only in the left row keeps its value t
// untouched (for f = add that differs from f(t, 0) exactly when t is -0.0).  Its iteration order is
// unspecified (RandomState), so rows sorted by column are a valid result for both branches.  `op` bit 1
// selects the unsorted branch's rule for left-only entries.
//
// Two passes like the product: count the union per row, look-back scan (scan.cu), fill.  One thread per row:
// HBM-bound on stencil-like matrices (reads A and B once, writes C once); rows whose columns are not sorted
// are first put in order by two transposes (dok.cu).
#include "common.cuh"

namespace {

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_ewise_count(u64 m, const u64* __restrict__ ap, const u32* __restrict__ ac,
                                                       const u64* __restrict__ bp, const u32* __restrict__ bc,
                                                       u32* __restrict__ row_nnz) {
  const u64 row = (u64)blockIdx.x * BLOCK + threadIdx.x;
  if (row >= m) return;
  u64 i = ap[row], j = bp[row];
  const u64 ie = ap[row + 1], je = bp[row + 1];
  u32 z = 0;
  while (i < ie && j < je) {
    const u32 ca = ac[i], cb = bc[j];
    i += ca <= cb ? 1 : 0;
    j += cb <= ca ? 1 : 0;
    ++z;
  }
  row_nnz[row] = z + (u32)(ie - i) + (u32)(je - j);
}

// OP 0: f = t1 + t2, OP 1: f = t1 - t2; the one-sided cases still go through f with a zero operand
// (lib.rs:114-115), so -0.0 + 0.0 = +0.0 comes out as in the reference.
template <class V, int OP>
__device__ __forceinline__ V apply(V x, V y) {
  if (OP == 0) return Num<V>::add(x, y);
  return Num<V>::sub(x, y);
}

constexpr int EW_STAGE = 3072;  // output entries a block stages in shared memory (128 rows x 24)

// One thread per row walks the two sorted rows.  The rows of a block are consecutive, so their output is one
// contiguous span of C: the threads write it into shared memory and the block then stores it with full sectors
// (thread-per-row stores straight to C are row-length strided: 32 sectors per store instruction, 0.18 of the copy
// peak in r1, 0.30 with the staged stores).  Staging the two INPUT spans the same way was tried and was slower (0.96 ms
// against 0.87 ms on A*A + A: 90 KB of shared memory per block leave two blocks per SM).  A block whose span does not
// fit (long rows) stores directly.
template <class V, int OP, bool KEEP_LEFT, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_ewise_fill(u64 m, const u64* __restrict__ ap, const u32* __restrict__ ac,
                                                      const V* __restrict__ av, const u64* __restrict__ bp,
                                                      const u32* __restrict__ bc, const V* __restrict__ bv,
                                                      const u64* __restrict__ cp, u32* __restrict__ cc,
                                                      V* __restrict__ cv) {
  __shared__ u32 sk[EW_STAGE];
  __shared__ V sv[EW_STAGE];
  const u64 row0 = (u64)blockIdx.x * BLOCK;
  const u64 row = row0 + threadIdx.x;
  const u64 rend = row0 + BLOCK < m ? row0 + BLOCK : m;
  const u64 base = cp[row0], span = cp[rend] - base;
  const bool staged = span <= (u64)EW_STAGE;  // block-uniform
  if (row < m) {
    u64 i = ap[row], j = bp[row], o = cp[row];
    const u64 ie = ap[row + 1], je = bp[row + 1];
    const V zero = Num<V>::zero();
    auto put = [&](u32 c, V v) {
      if (staged) { sk[o - base] = c; sv[o - base] = v; } else { cc[o] = c; cv[o] = v; }
      ++o;
    };
    while (i < ie && j < je) {
      const u32 ca = ac[i], cb = bc[j];
      if (ca == cb) { put(ca, apply<V, OP>(av[i], bv[j])); ++i; ++j; }
      else if (ca < cb) { put(ca, KEEP_LEFT ? av[i] : apply<V, OP>(av[i], zero)); ++i; }
      else { put(cb, apply<V, OP>(zero, bv[j])); ++j; }
    }
    for (; i < ie; ++i) put(ac[i], KEEP_LEFT ? av[i] : apply<V, OP>(av[i], zero));
    for (; j < je; ++j) put(bc[j], apply<V, OP>(zero, bv[j]));
  }
  if (!staged) return;
  __syncthreads();
  for (u64 q = threadIdx.x; q < span; q += BLOCK) { cc[base + q] = sk[q]; cv[base + q] = sv[q]; }
}

// The same walk with the block's three spans in shared memory.  The rows of a block are consecutive, so what it reads
// of A and of B are two contiguous entry ranges: one thread fetches them (columns and values: four 1-D bulk copies,
// cp.async.bulk on an mbarrier, SASS UBLKCP) while the others wait; the per-row merge then reads shared memory only,
// writes the output span to shared memory, and the block stores it with full sectors.  k_ewise_fill lets every thread
// walk its own row in global memory: 32 sectors per load instruction and, with 36 KB of staging per block, an L1 too
// small to keep them (hit rate 4 %, 5.3 GB moved from L2 for 1.0 GB of input: profiles/r02_ewise_fill.txt).
// capA / capB / capC (multiples of 4 entries) are sized by the host from the mean row lengths; a block whose spans do
// not fit walks global memory and stores directly.  Layout: [mbarrier][A cols][B cols][C cols][A vals][B vals][C vals].
template <class V, int OP, bool KEEP_LEFT, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_ewise_fill_tma(u64 m, const u64* __restrict__ ap, const u32* __restrict__ ac,
                                                          const V* __restrict__ av, u64 a_nnz,
                                                          const u64* __restrict__ bp, const u32* __restrict__ bc,
                                                          const V* __restrict__ bv, u64 b_nnz,
                                                          const u64* __restrict__ cp, u32* __restrict__ cc,
                                                          V* __restrict__ cv, u32 capA, u32 capB, u32 capC) {
  extern __shared__ __align__(16) unsigned char sm_ew[];
  u64* bar = reinterpret_cast<u64*>(sm_ew);
  u32* sAk = reinterpret_cast<u32*>(sm_ew + 16);
  u32* sBk = sAk + capA;
  u32* sCk = sBk + capB;
  V* sAv = reinterpret_cast<V*>(sCk + capC);
  V* sBv = sAv + capA;
  V* sCv = sBv + capB;
  const int tid = threadIdx.x;
  if (tid == 0) mbar_init(bar, 1);
  const u64 row0 = (u64)blockIdx.x * BLOCK;
  const u64 row = row0 + tid;
  const u64 rend = row0 + BLOCK < m ? row0 + BLOCK : m;
  const u64 a0 = ap[row0], a1 = ap[rend], b0 = bp[row0], b1 = bp[rend], c0 = cp[row0], c1 = cp[rend];
  const u64 a0a = a0 & ~(u64)3, b0a = b0 & ~(u64)3;  // 16-byte aligned starts
  const bool staged = a1 - a0a <= (u64)capA && b1 - b0a <= (u64)capB && c1 - c0 <= (u64)capC;  // block-uniform
  const V zero = Num<V>::zero();
  __syncthreads();
  if (!staged) {
    if (row < m) {
      u64 i = ap[row], j = bp[row], o = cp[row];
      const u64 ie = ap[row + 1], je = bp[row + 1];
      auto put = [&](u32 c, V v) { cc[o] = c; cv[o] = v; ++o; };
      while (i < ie && j < je) {
        const u32 ca = ac[i], cb = bc[j];
        if (ca == cb) { put(ca, apply<V, OP>(av[i], bv[j])); ++i; ++j; }
        else if (ca < cb) { put(ca, KEEP_LEFT ? av[i] : apply<V, OP>(av[i], zero)); ++i; }
        else { put(cb, apply<V, OP>(zero, bv[j])); ++j; }
      }
      for (; i < ie; ++i) put(ac[i], KEEP_LEFT ? av[i] : apply<V, OP>(av[i], zero));
      for (; j < je; ++j) put(bc[j], apply<V, OP>(zero, bv[j]));
    }
    return;
  }
  const u64 a4 = a_nnz & ~(u64)3, b4 = b_nnz & ~(u64)3;  // a 16-byte copy must not run past the end of the arrays
  if (tid == 0) {
    const u64 ra = (a1 + 3) & ~(u64)3, rb = (b1 + 3) & ~(u64)3;
    const u64 ea = ra < a4 ? ra : a4, eb = rb < b4 ? rb : b4;
    const u32 la = ea > a0a ? (u32)(ea - a0a) : 0, lb = eb > b0a ? (u32)(eb - b0a) : 0;
    mbar_expect_tx(bar, (la + lb) * (4u + (u32)sizeof(V)));
    if (la) { bulk_g2s(sAk, ac + a0a, la * 4u, bar); bulk_g2s(sAv, av + a0a, la * (u32)sizeof(V), bar); }
    if (lb) { bulk_g2s(sBk, bc + b0a, lb * 4u, bar); bulk_g2s(sBv, bv + b0a, lb * (u32)sizeof(V), bar); }
  }
  for (u64 e = (a0a > a4 ? a0a : a4) + tid; e < a1; e += BLOCK) { sAk[e - a0a] = ac[e]; sAv[e - a0a] = av[e]; }
  for (u64 e = (b0a > b4 ? b0a : b4) + tid; e < b1; e += BLOCK) { sBk[e - b0a] = bc[e]; sBv[e - b0a] = bv[e]; }
  u32 i = 0, ie = 0, j = 0, je = 0, o = 0;
  if (row < m) {
    i = (u32)(ap[row] - a0a); ie = (u32)(ap[row + 1] - a0a);
    j = (u32)(bp[row] - b0a); je = (u32)(bp[row + 1] - b0a);
    o = (u32)(cp[row] - c0);
  }
  mbar_wait(bar, 0);
  __syncthreads();
  {
    auto put = [&](u32 c, V v) { sCk[o] = c; sCv[o] = v; ++o; };
    while (i < ie && j < je) {
      const u32 ca = sAk[i], cb = sBk[j];
      if (ca == cb) { put(ca, apply<V, OP>(sAv[i], sBv[j])); ++i; ++j; }
      else if (ca < cb) { put(ca, KEEP_LEFT ? sAv[i] : apply<V, OP>(sAv[i], zero)); ++i; }
      else { put(cb, apply<V, OP>(zero, sBv[j])); ++j; }
    }
    for (; i < ie; ++i) put(sAk[i], KEEP_LEFT ? sAv[i] : apply<V, OP>(sAv[i], zero));
    for (; j < je; ++j) put(sBk[j], apply<V, OP>(zero, sBv[j]));
  }
  __syncthreads();
  const u32 span = (u32)(c1 - c0);
  for (u32 q = tid; q < span; q += BLOCK) { cc[c0 + q] = sCk[q]; cv[c0 + q] = sCv[q]; }
}

template <class V>
int fill_typed(spam_handle* h, int op, const spam_dcsr* a, const spam_dcsr* b, spam_dcsr* c) {
  constexpr int BL = 128;
  const unsigned grid = (unsigned)((a->rows + BL - 1) / BL);
  // span capacities: a block of BL mean rows plus 1/16 and a few entries of slack
  auto cap_of = [&](u64 nnz) { return (u32)(((u64)((double)nnz / (double)a->rows * BL * 1.0625) + 64 + 3) & ~3ull); };
  const u32 capA = cap_of(a->nnz), capB = cap_of(b->nnz), capC = cap_of(c->nnz);
  const size_t tsmem = 16 + (size_t)(capA + capB + capC) * (4 + sizeof(V));
  const bool aligned = (((uintptr_t)a->idx | (uintptr_t)a->val | (uintptr_t)b->idx | (uintptr_t)b->val) & 15) == 0;
  if (h->ewise_tma && aligned && tsmem <= 72 * 1024) {
#define EW_TMA(OP, KEEP)                                                                                              \
  {                                                                                                                   \
    CK(cudaFuncSetAttribute(k_ewise_fill_tma<V, OP, KEEP, BL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024)); \
    k_ewise_fill_tma<V, OP, KEEP, BL><<<grid, BL, tsmem, h->stream>>>(a->rows, a->ptr, a->idx, (const V*)a->val,      \
        a->nnz, b->ptr, b->idx, (const V*)b->val, b->nnz, c->ptr, c->idx, (V*)c->val, capA, capB, capC);              \
  }
    switch (op) {
      case 0: EW_TMA(0, false); break;
      case 1: EW_TMA(1, false); break;
      case 2: EW_TMA(0, true); break;
      default: EW_TMA(1, true); break;
    }
#undef EW_TMA
    count_launch(h);
    CK(cudaGetLastError());
    return SPAM_OK;
  }
#define EW_LAUNCH(OP, KEEP)                                                                                          \
  k_ewise_fill<V, OP, KEEP, BL><<<grid, BL, 0, h->stream>>>(a->rows, a->ptr, a->idx, (const V*)a->val, b->ptr, b->idx, \
                                                            (const V*)b->val, c->ptr, c->idx, (V*)c->val)
  switch (op) {
    case 0: EW_LAUNCH(0, false); break;
    case 1: EW_LAUNCH(1, false); break;
    case 2: EW_LAUNCH(0, true); break;
    default: EW_LAUNCH(1, true); break;
  }
#undef EW_LAUNCH
  count_launch(h);
  CK(cudaGetLastError());
  return SPAM_OK;
}

void free_owned(spam_handle* h, spam_dcsr* m) {
  if (!m) return;
  dev_free(h, m->ptr); dev_free(h, m->idx); dev_free(h, m->val);
  delete m;
}

}  // namespace

int ewise_dev(spam_handle* h, int op, const spam_dcsr* a_in, const spam_dcsr* b_in, spam_dcsr** out) {
  *out = nullptr;
  if (op < 0 || op > 3) return spam_fail(h, SPAM_EINVAL, "op must be 0 (add), 1 (sub), optionally | 2 (IS_SORTED = false rule)");
  // lib.rs:87-91: assert_eq!((self.rows, self.cols), (rhs.rows, rhs.cols))
  if (a_in->rows != b_in->rows || a_in->cols != b_in->cols) return spam_fail(h, SPAM_EDIM, "matrices must have identical dimensions");
  if (a_in->dtype != b_in->dtype) return spam_fail(h, SPAM_EDTYPE, "operand dtypes differ");
  // rows that are not sorted by column: the cached sorted copy of the operand (two stable transposes, dok.cu)
  const spam_dcsr *a = nullptr, *b = nullptr;
  int st = sorted_rows_of(h, a_in, &a);
  if (st == SPAM_OK) st = (b_in == a_in) ? (b = a, SPAM_OK) : sorted_rows_of(h, b_in, &b);
  if (st != SPAM_OK) return st;
  h->stats = spam_stats{};
  const u64 m = a->rows;
  spam_dcsr* c = new spam_dcsr();
  c->dtype = a->dtype; c->rows = m; c->cols = a->cols; c->nnz = 0; c->owning = true; c->rows_sorted = -1; c->max_row_len = 0;  // stats taken lazily if C becomes an operand
  c->ptr = nullptr; c->idx = nullptr; c->val = nullptr;
  u32* row_nnz = nullptr;
  auto fail = [&](int s) {
    dev_free(h, row_nnz);
    free_owned(h, c);
    return s;
  };
  st = dev_alloc_t(h, &row_nnz, m);
  if (st == SPAM_OK) st = dev_alloc_t(h, &c->ptr, m + 1);
  if (st != SPAM_OK) return fail(st);
  cudaError_t e = cudaMemsetAsync(h->d_cnt, 0, sizeof(Counters), h->stream);
  if (e != cudaSuccess) return fail(spam_fail(h, SPAM_ECUDA, "cudaMemsetAsync", e));
  constexpr int BL = 128;
  k_ewise_count<BL><<<(unsigned)((m + BL - 1) / BL), BL, 0, h->stream>>>(m, a->ptr, a->idx, b->ptr, b->idx, row_nnz);
  count_launch(h);
  if ((e = cudaGetLastError()) != cudaSuccess) return fail(spam_fail(h, SPAM_ECUDA, "k_ewise_count", e));
  st = scan_u32_to_u64(h, row_nnz, c->ptr, m, &h->d_cnt->total_nnz);
  if (st != SPAM_OK) return fail(st);
  e = cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return fail(spam_fail(h, SPAM_ECUDA, "ewise sync", e));
  c->nnz = h->h_cnt->total_nnz;
  st = dev_alloc_t(h, &c->idx, c->nnz);
  if (st == SPAM_OK) st = dev_alloc(h, &c->val, c->nnz * dtype_size(c->dtype));
  if (st == SPAM_OK && c->nnz) {
    switch (c->dtype) {
      case SPAM_F32: st = fill_typed<float>(h, op, a, b, c); break;
      case SPAM_F64: st = fill_typed<double>(h, op, a, b, c); break;
      case SPAM_I32: st = fill_typed<int32_t>(h, op, a, b, c); break;
      case SPAM_I64: st = fill_typed<int64_t>(h, op, a, b, c); break;
      default: st = spam_fail(h, SPAM_EINVAL, "bad dtype");
    }
  }
  if (st != SPAM_OK) return fail(st);
  dev_free(h, row_nnz);
  h->stats.nnz_c = c->nnz;
  *out = c;
  return SPAM_OK;
}
