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

constexpr int EW_IN = 2304;   // entries of each operand a block stages in shared memory (128 rows x 18)
constexpr int EW_OUT = 3072;  // output entries a block stages (128 rows x 24)
template <class V>
constexpr size_t ewise_smem() { return (size_t)(2 * EW_IN + EW_OUT) * (sizeof(V) + 4); }

// One thread per row walks the two sorted rows.  The rows of a block are consecutive, so its slices of A, B and C
// are three contiguous spans: the block loads the two input spans into shared memory with full-sector loads, the
// threads merge out of shared memory into a staged output span, and the block stores that with full sectors.
// (Thread-per-row loads and stores straight from / to global memory are row-length strided — 32 sectors per
// instruction: 0.18 of the copy peak in r1, 0.30 with only the stores staged.)  A block whose spans do not fit
// (long rows) works on global memory directly.
template <class V, int OP, bool KEEP_LEFT, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_ewise_fill(u64 m, const u64* __restrict__ ap, const u32* __restrict__ ac,
                                                      const V* __restrict__ av, const u64* __restrict__ bp,
                                                      const u32* __restrict__ bc, const V* __restrict__ bv,
                                                      const u64* __restrict__ cp, u32* __restrict__ cc,
                                                      V* __restrict__ cv) {
  extern __shared__ __align__(16) unsigned char sm_ew[];
  V* sva = reinterpret_cast<V*>(sm_ew);
  V* svb = sva + EW_IN;
  V* svo = svb + EW_IN;
  u32* ska = reinterpret_cast<u32*>(svo + EW_OUT);
  u32* skb = ska + EW_IN;
  u32* sko = skb + EW_IN;
  const u64 row0 = (u64)blockIdx.x * BLOCK;
  const u64 row = row0 + threadIdx.x;
  const u64 rend = row0 + BLOCK < m ? row0 + BLOCK : m;
  const u64 a0 = ap[row0], b0 = bp[row0], c0 = cp[row0];
  const u64 na = ap[rend] - a0, nb = bp[rend] - b0, nc = cp[rend] - c0;
  const bool staged = na <= (u64)EW_IN && nb <= (u64)EW_IN && nc <= (u64)EW_OUT;  // block-uniform
  const V zero = Num<V>::zero();
  if (staged) {
    for (u64 q = threadIdx.x; q < na; q += BLOCK) { ska[q] = ac[a0 + q]; sva[q] = av[a0 + q]; }
    for (u64 q = threadIdx.x; q < nb; q += BLOCK) { skb[q] = bc[b0 + q]; svb[q] = bv[b0 + q]; }
    __syncthreads();
    if (row < m) {
      u32 i = (u32)(ap[row] - a0), j = (u32)(bp[row] - b0), o = (u32)(cp[row] - c0);
      const u32 ie = (u32)(ap[row + 1] - a0), je = (u32)(bp[row + 1] - b0);
      while (i < ie && j < je) {
        const u32 ca = ska[i], cb = skb[j];
        if (ca == cb) { sko[o] = ca; svo[o] = apply<V, OP>(sva[i], svb[j]); ++i; ++j; }
        else if (ca < cb) { sko[o] = ca; svo[o] = KEEP_LEFT ? sva[i] : apply<V, OP>(sva[i], zero); ++i; }
        else { sko[o] = cb; svo[o] = apply<V, OP>(zero, svb[j]); ++j; }
        ++o;
      }
      for (; i < ie; ++i, ++o) { sko[o] = ska[i]; svo[o] = KEEP_LEFT ? sva[i] : apply<V, OP>(sva[i], zero); }
      for (; j < je; ++j, ++o) { sko[o] = skb[j]; svo[o] = apply<V, OP>(zero, svb[j]); }
    }
    __syncthreads();
    for (u64 q = threadIdx.x; q < nc; q += BLOCK) { cc[c0 + q] = sko[q]; cv[c0 + q] = svo[q]; }
    return;
  }
  if (row >= m) return;
  u64 i = ap[row], j = bp[row], o = cp[row];
  const u64 ie = ap[row + 1], je = bp[row + 1];
  while (i < ie && j < je) {
    const u32 ca = ac[i], cb = bc[j];
    if (ca == cb) { cc[o] = ca; cv[o] = apply<V, OP>(av[i], bv[j]); ++i; ++j; }
    else if (ca < cb) { cc[o] = ca; cv[o] = KEEP_LEFT ? av[i] : apply<V, OP>(av[i], zero); ++i; }
    else { cc[o] = cb; cv[o] = apply<V, OP>(zero, bv[j]); ++j; }
    ++o;
  }
  for (; i < ie; ++i, ++o) { cc[o] = ac[i]; cv[o] = KEEP_LEFT ? av[i] : apply<V, OP>(av[i], zero); }
  for (; j < je; ++j, ++o) { cc[o] = bc[j]; cv[o] = apply<V, OP>(zero, bv[j]); }
}

template <class V>
int fill_typed(spam_handle* h, int op, const spam_dcsr* a, const spam_dcsr* b, spam_dcsr* c) {
  constexpr int BL = 128;
  const unsigned grid = (unsigned)((a->rows + BL - 1) / BL);
  constexpr size_t smem = ewise_smem<V>();
#define EW_LAUNCH(OP, KEEP)                                                                                          \
  CK(cudaFuncSetAttribute(k_ewise_fill<V, OP, KEEP, BL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
  k_ewise_fill<V, OP, KEEP, BL><<<grid, BL, smem, h->stream>>>(a->rows, a->ptr, a->idx, (const V*)a->val, b->ptr, b->idx, \
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
