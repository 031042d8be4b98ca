// spmv.cu — CSR SpMV  y = A x  (dense x, dense y).
// The reference has no SpMV; the semantics are those of
// a.mul_hash::<_, true>(&x) with x an n x 1 CsrMatrix: y_i = sum over A's row i of a_ik * x_k,
// product rounded, then added (mul_hash.rs:154-161), in A-row storage order; rows without entries
// give zero.
//
// HBM-bound: every A entry (4 + s bytes) is read once, x is gathered (L1/L2 resident for banded
// matrices), y written once: nnz*(4+s) + (m+1)*8 + n*s + m*s bytes.
//
//  * k_spmv_stream (short rows, max row length <= 64): a block owns 256 consecutive rows; their entries
//    are one contiguous span of col_idx/val, which the block streams with fully coalesced loads,
//    multiplies by the gathered x and parks in shared memory; then thread t adds up row t's products
//    sequentially — the reference's order, so the result is bit-identical to the oracle.
//  * k_spmv_vector (long or skewed rows): L = 2^k lanes per row, partial sums folded by warp shuffles.
#include "common.cuh"

namespace {

constexpr int SP_BLOCK = 256;
constexpr int SP_TILE = 2048;  // products parked per pass

template <class V>
__global__ void __launch_bounds__(SP_BLOCK) k_spmv_stream(u64 m, const u64* __restrict__ ptr,
                                                          const u32* __restrict__ idx, const V* __restrict__ val,
                                                          const V* __restrict__ x, V* __restrict__ y) {
  __shared__ u64 s_ptr[SP_BLOCK + 1];
  __shared__ V s_prod[SP_TILE];
  const int tid = threadIdx.x;
  const u64 r0 = (u64)blockIdx.x * SP_BLOCK;
  const u64 row = r0 + tid;
  s_ptr[tid] = ptr[row < m ? row : m];
  if (tid == 0) s_ptr[SP_BLOCK] = ptr[(r0 + SP_BLOCK) < m ? (r0 + SP_BLOCK) : m];
  __syncthreads();
  const u64 base = s_ptr[0], end = s_ptr[SP_BLOCK];
  const u64 lo = s_ptr[tid], hi = s_ptr[tid + 1];
  V acc = Num<V>::zero();
  bool first = true;
  for (u64 t0 = base; t0 < end; t0 += SP_TILE) {
    const u64 t1 = (t0 + SP_TILE < end) ? t0 + SP_TILE : end;
    // all loads of the tile in flight at once: SP_TILE / SP_BLOCK independent (col, val) pairs per thread, then the
    // gathers of x (a loop with a data-dependent trip count kept one pair in flight: 0.55 of the copy peak)
    {
      u32 ci[SP_TILE / SP_BLOCK];
      V vv[SP_TILE / SP_BLOCK];
#pragma unroll
      for (int i = 0; i < SP_TILE / SP_BLOCK; ++i) {
        const u64 e = t0 + tid + (u64)i * SP_BLOCK;
        ci[i] = 0; vv[i] = Num<V>::zero();
        if (e < t1) { ci[i] = idx[e]; vv[i] = val[e]; }
      }
#pragma unroll
      for (int i = 0; i < SP_TILE / SP_BLOCK; ++i) {
        const u64 e = t0 + tid + (u64)i * SP_BLOCK;
        if (e < t1) s_prod[e - t0] = Num<V>::mul(vv[i], x[ci[i]]);
      }
    }
    __syncthreads();
    const u64 a = lo > t0 ? lo : t0, b = hi < t1 ? hi : t1;
    for (u64 e = a; e < b; ++e) {
      const V p = s_prod[e - t0];
      acc = first ? p : Num<V>::add(acc, p);  // first product stored, not added to 0
      first = false;
    }
    __syncthreads();
  }
  if (row < m) y[row] = acc;
}

template <class V, int L>
__global__ void __launch_bounds__(256) k_spmv_vector(u64 m, const u64* __restrict__ ptr, const u32* __restrict__ idx,
                                                     const V* __restrict__ val, const V* __restrict__ x,
                                                     V* __restrict__ y) {
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
  const V* av = (const V*)a->val;
  const double mean = (double)a->nnz / (double)m;
  if (a->rows_sorted >= 0 && a->max_row_len <= 64) {
    k_spmv_stream<V><<<(unsigned)((m + SP_BLOCK - 1) / SP_BLOCK), SP_BLOCK, 0, h->stream>>>(m, a->ptr, a->idx, av, x, y);
    count_launch(h);
    CK(cudaGetLastError());
    return SPAM_OK;
  }
  int L = 1;
  while (L < 32 && (double)(L * 2) <= mean) L <<= 1;  // largest power of two <= mean row length
  const u64 threads = m * (u64)L;
  const unsigned grid = (unsigned)((threads + 255) / 256);
  switch (L) {
    case 1: k_spmv_vector<V, 1><<<grid, 256, 0, h->stream>>>(m, a->ptr, a->idx, av, x, y); break;
    case 2: k_spmv_vector<V, 2><<<grid, 256, 0, h->stream>>>(m, a->ptr, a->idx, av, x, y); break;
    case 4: k_spmv_vector<V, 4><<<grid, 256, 0, h->stream>>>(m, a->ptr, a->idx, av, x, y); break;
    case 8: k_spmv_vector<V, 8><<<grid, 256, 0, h->stream>>>(m, a->ptr, a->idx, av, x, y); break;
    case 16: k_spmv_vector<V, 16><<<grid, 256, 0, h->stream>>>(m, a->ptr, a->idx, av, x, y); break;
    default: k_spmv_vector<V, 32><<<grid, 256, 0, h->stream>>>(m, a->ptr, a->idx, av, x, y); break;
  }
  count_launch(h);
  CK(cudaGetLastError());
  return SPAM_OK;
}

}  // namespace

int spmv_dev(spam_handle* h, const spam_dcsr* a, const void* d_x, void* d_y) {
  CKS(ensure_matrix_stats(h, a));  // cached per matrix: longest row picks the kernel
  switch (a->dtype) {
    case SPAM_F32: return launch_spmv<float>(h, a, (const float*)d_x, (float*)d_y);
    case SPAM_F64: return launch_spmv<double>(h, a, (const double*)d_x, (double*)d_y);
    case SPAM_I32: return launch_spmv<int32_t>(h, a, (const int32_t*)d_x, (int32_t*)d_y);
    case SPAM_I64: return launch_spmv<int64_t>(h, a, (const int64_t*)d_x, (int64_t*)d_y);
    default: return spam_fail(h, SPAM_EINVAL, "bad dtype");
  }
}
