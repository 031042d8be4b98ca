// spmv.cu — CSR SpMV  y = A x  (dense x, dense y).
// The reference has no SpMV; the semantics are those of
// a.mul_hash::<_, true>(&x) with x an n x 1 CsrMatrix: y_i = sum over A's row i of a_ik * x_k,
// product rounded, then added (mul_hash.rs:154-161); rows without entries give zero.
//
// HBM-bound: every A entry (4 + s bytes) is read once, x is gathered (L2/L1 resident for banded
// matrices).  L = 2^k lanes cooperate on one row so that a warp's loads of col_idx/val cover a
// contiguous span of the arrays; partial sums are folded with warp shuffles.
#include "common.cuh"

namespace {

template <class V, int L>
__global__ void __launch_bounds__(256) k_spmv(u64 m, const u64* __restrict__ ptr, const u32* __restrict__ idx,
                                              const V* __restrict__ val, const V* __restrict__ x, V* __restrict__ y) {
  const u64 gtid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u64 row = gtid / L;
  const int lane = (int)(gtid % L);
  V acc = Num<V>::zero();
  if (row < m) {
    const u64 lo = ptr[row], hi = ptr[row + 1];
    for (u64 e = lo + lane; e < hi; e += L) acc = Num<V>::add(acc, Num<V>::mul(val[e], x[idx[e]]));
  }
#pragma unroll
  for (int d = L >> 1; d > 0; d >>= 1) acc = Num<V>::add(acc, __shfl_xor_sync(0xffffffffu, acc, d));
  if (row < m && lane == 0) y[row] = acc;
}

template <class V>
int launch_spmv(spam_handle* h, const spam_dcsr* a, const V* x, V* y) {
  const u64 m = a->rows;
  if (m == 0) return SPAM_OK;
  const double mean = (double)a->nnz / (double)m;
  int L = 1;
  while (L < 32 && (double)(L * 2) <= mean) L <<= 1;  // largest power of two <= mean row length
  const u64 threads = m * (u64)L;
  const unsigned grid = (unsigned)((threads + 255) / 256);
  const V* av = (const V*)a->val;
  switch (L) {
    case 1: k_spmv<V, 1><<<grid, 256, 0, h->stream>>>(m, a->ptr, a->idx, av, x, y); break;
    case 2: k_spmv<V, 2><<<grid, 256, 0, h->stream>>>(m, a->ptr, a->idx, av, x, y); break;
    case 4: k_spmv<V, 4><<<grid, 256, 0, h->stream>>>(m, a->ptr, a->idx, av, x, y); break;
    case 8: k_spmv<V, 8><<<grid, 256, 0, h->stream>>>(m, a->ptr, a->idx, av, x, y); break;
    case 16: k_spmv<V, 16><<<grid, 256, 0, h->stream>>>(m, a->ptr, a->idx, av, x, y); break;
    default: k_spmv<V, 32><<<grid, 256, 0, h->stream>>>(m, a->ptr, a->idx, av, x, y); break;
  }
  count_launch(h);
  CK(cudaGetLastError());
  return SPAM_OK;
}

}  // namespace

int spmv_dev(spam_handle* h, const spam_dcsr* a, const void* d_x, void* d_y) {
  switch (a->dtype) {
    case SPAM_F32: return launch_spmv<float>(h, a, (const float*)d_x, (float*)d_y);
    case SPAM_F64: return launch_spmv<double>(h, a, (const double*)d_x, (double*)d_y);
    case SPAM_I32: return launch_spmv<int32_t>(h, a, (const int32_t*)d_x, (int32_t*)d_y);
    case SPAM_I64: return launch_spmv<int64_t>(h, a, (const int64_t*)d_x, (int64_t*)d_y);
    default: return spam_fail(h, SPAM_EINVAL, "bad dtype");
  }
}
