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

// k_spmv_tma: the same scheme as a persistent, software-pipelined kernel.  In k_spmv_stream a block's three dependent
// round trips (row_ptr -> the span of col_idx / val -> the gathers of x) run one after the other, hidden only by the
// other resident blocks: 0.63 of the copy peak.  Here a block walks row blocks blockIdx.x, + gridDim.x, ...; thread 0
// fetches the span of the item ST_STAGES - 1 ahead with two 1-D bulk copies (cp.async.bulk, completion on the
// stage's mbarrier, SASS UBLKCP) while all threads gather x for the current one, so the stream of A never stops.
// A row block whose span exceeds ST_TILE entries is fetched in several chunks; the per-row sums carry over in
// registers.  Products and their order are those of k_spmv_stream (bit-identical results).
// Measured on B200 (Poisson 2048^2 f64): 0.0846 ms against 0.0861 ms — both wait for the gathers of x (long scoreboard
// 30 % of the samples; DRAM 50 %, L1 57 %, issue 41 % busy: profiles/r02_spmv.txt), so this kernel is opt-in
// (SPAM_SPMV_TMA=1).  A first version with three stages of 2048 entries (3 blocks per SM, 78 registers) ran 0.0969 ms.
constexpr int ST_TILE = 1536, ST_STAGES = 2, ST_BLOCK = 256, ST_OCC = 6;
struct StMeta { u64 e0, lo, hi, rb; };  // entries [lo, hi) of row block rb are at offsets (e - e0) of the stage; rb = ~0: no more work
constexpr u64 ST_DONE = ~(u64)0, ST_LAST = (u64)1 << 63;  // ST_LAST in rb: the row block ends with this chunk

template <class V>
constexpr size_t spmv_tma_smem() { return 192 + (size_t)ST_STAGES * ST_TILE * (4 + sizeof(V)); }

template <class V>
__global__ void __launch_bounds__(ST_BLOCK, ST_OCC) k_spmv_tma(u64 m, u64 nnz, const u64* __restrict__ ptr,
                                                       const u32* __restrict__ idx, const V* __restrict__ val,
                                                       const V* __restrict__ x, V* __restrict__ y) {
  extern __shared__ __align__(16) unsigned char sm_spmv[];
  u64* bar = reinterpret_cast<u64*>(sm_spmv);                   // [ST_STAGES]
  StMeta* meta = reinterpret_cast<StMeta*>(sm_spmv + 64);       // [ST_STAGES]
  u32* s_idx = reinterpret_cast<u32*>(sm_spmv + 192);           // [ST_STAGES][ST_TILE]
  V* s_val = reinterpret_cast<V*>(s_idx + ST_STAGES * ST_TILE); // [ST_STAGES][ST_TILE]; products are written over the values
  const int tid = threadIdx.x;
  const u64 nrb = (m + ST_BLOCK - 1) / ST_BLOCK;
  const u64 nnz4 = nnz & ~(u64)3;  // a 16-byte copy must not run past the end of the arrays
  if (tid == 0)
    for (int s = 0; s < ST_STAGES; ++s) mbar_init(&bar[s], 1);
  __syncthreads();

  // ---- producer (thread 0): the next chunk to fetch; the bounds of the following row block are loaded one ahead
  u64 p_rb = blockIdx.x, p_b0 = 0, p_b1 = 0, p_e = 0, p_nb0 = 0, p_nb1 = 0;
  bool p_have = false, p_done = false;
  auto bounds = [&](u64 rb, u64& b0, u64& b1) {
    if (rb < nrb) {
      const u64 r0 = rb * ST_BLOCK, r1 = r0 + ST_BLOCK < m ? r0 + ST_BLOCK : m;
      b0 = ptr[r0]; b1 = ptr[r1];
    }
  };
  if (tid == 0) bounds(p_rb, p_nb0, p_nb1);
  auto produce = [&](int stage) {
    if (p_done) return;
    if (!p_have) {
      if (p_rb >= nrb) {
        meta[stage].rb = ST_DONE;
        mbar_expect_tx(&bar[stage], 0);
        p_done = true;
        return;
      }
      p_b0 = p_nb0; p_b1 = p_nb1;
      bounds(p_rb + gridDim.x, p_nb0, p_nb1);
      p_e = p_b0 & ~(u64)3;
      p_have = true;
    }
    const u64 hi = p_e + ST_TILE < p_b1 ? p_e + ST_TILE : p_b1;
    const u64 lo = p_e > p_b0 ? p_e : p_b0;
    u64 ce = (hi + 3) & ~(u64)3;
    if (ce > nnz4) ce = nnz4;
    const u32 len = ce > p_e ? (u32)(ce - p_e) : 0;
    const bool last = hi >= p_b1;
    meta[stage].e0 = p_e; meta[stage].lo = lo; meta[stage].hi = hi; meta[stage].rb = p_rb | (last ? ST_LAST : 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the stage was written (products) by ordinary stores
    mbar_expect_tx(&bar[stage], len * (4u + (u32)sizeof(V)));
    if (len) {
      bulk_g2s(s_idx + (size_t)stage * ST_TILE, idx + p_e, len * 4u, &bar[stage]);
      bulk_g2s(s_val + (size_t)stage * ST_TILE, val + p_e, len * (u32)sizeof(V), &bar[stage]);
    }
    p_e += ST_TILE;
    if (last) { p_have = false; p_rb += gridDim.x; }
  };
  if (tid == 0)
    for (int s = 0; s < ST_STAGES - 1; ++s) produce(s);

  // ---- consumers (all threads): row tid of the current row block; its extent is loaded one row block ahead
  u64 c_next = blockIdx.x, n_lo = 0, n_hi = 0, lo_r = 0, hi_r = 0, cur = ST_DONE;
  auto row_extent = [&](u64 rb, u64& lo, u64& hi) {
    if (rb < nrb) {
      const u64 row = rb * ST_BLOCK + tid;
      lo = ptr[row < m ? row : m]; hi = ptr[row + 1 < m ? row + 1 : m];
    }
  };
  row_extent(c_next, n_lo, n_hi);
  V acc = Num<V>::zero();
  bool first = true;
  for (u32 it = 0;; ++it) {
    const int stage = it % ST_STAGES;
    if (tid == 0) produce((it + ST_STAGES - 1) % ST_STAGES);
    mbar_wait(&bar[stage], (it / ST_STAGES) & 1);
    const StMeta mt = meta[stage];
    if (mt.rb == ST_DONE) break;
    const u64 rb = mt.rb & ~ST_LAST;
    if (rb != cur) {
      cur = rb; lo_r = n_lo; hi_r = n_hi;
      c_next += gridDim.x;
      row_extent(c_next, n_lo, n_hi);
      acc = Num<V>::zero(); first = true;
    }
    u32* ki = s_idx + (size_t)stage * ST_TILE;
    V* vi = s_val + (size_t)stage * ST_TILE;
    // offsets within the stage: entries [k_lo, k_hi) are this chunk's; those from k_bulk on lie past the last
    // 16-byte boundary of the arrays and were not part of the bulk copy
    const u32 k_lo = (u32)(mt.lo - mt.e0), k_hi = (u32)(mt.hi - mt.e0);
    const u32 k_bulk = nnz4 > mt.e0 ? (nnz4 - mt.e0 < (u64)ST_TILE ? (u32)(nnz4 - mt.e0) : (u32)ST_TILE) : 0u;
    {
      V xv[ST_TILE / ST_BLOCK];
#pragma unroll
      for (int i = 0; i < ST_TILE / ST_BLOCK; ++i) {
        const u32 k = tid + i * ST_BLOCK;
        xv[i] = Num<V>::zero();
        if (k >= k_lo && k < k_hi) {
          if (k >= k_bulk) { ki[k] = idx[mt.e0 + k]; vi[k] = val[mt.e0 + k]; }  // own slot: no other thread reads it before the barrier
          xv[i] = x[ki[k]];
        }
      }
#pragma unroll
      for (int i = 0; i < ST_TILE / ST_BLOCK; ++i) {
        const u32 k = tid + i * ST_BLOCK;
        if (k >= k_lo && k < k_hi) vi[k] = Num<V>::mul(vi[k], xv[i]);
      }
    }
    __syncthreads();
    {
      const u64 a = lo_r > mt.lo ? lo_r : mt.lo, b = hi_r < mt.hi ? hi_r : mt.hi;
      const u32 ka = (u32)(a - mt.e0);
      const u32 kb = b > a ? (u32)(b - mt.e0) : ka;
      for (u32 k = ka; k < kb; ++k) {
        const V p = vi[k];
        acc = first ? p : Num<V>::add(acc, p);  // first product stored, not added to 0
        first = false;
      }
    }
    if (mt.rb & ST_LAST) {
      const u64 row = rb * ST_BLOCK + tid;
      if (row < m) y[row] = acc;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // products were ordinary stores; the refill is an async-proxy write
    __syncthreads();  // the stage may be refilled
  }
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
  if (h->spmv_tma && a->rows_sorted >= 0 && a->max_row_len <= 64 && (((uintptr_t)a->idx | (uintptr_t)a->val) & 15) == 0) {
    constexpr size_t smem = spmv_tma_smem<V>();
    const u64 nrb = (m + ST_BLOCK - 1) / ST_BLOCK;
    const int per_sm_smem = (int)((227 * 1024) / (smem + 1024));
    const int per_sm = per_sm_smem < ST_OCC ? per_sm_smem : ST_OCC;
    const u64 cap = (u64)h->num_sms * (u64)(per_sm > 0 ? per_sm : 1);
    CK(cudaFuncSetAttribute(k_spmv_tma<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_spmv_tma<V><<<(unsigned)(nrb < cap ? nrb : cap), ST_BLOCK, smem, h->stream>>>(m, a->nnz, a->ptr, a->idx, av, x, y);
    count_launch(h);
    CK(cudaGetLastError());
    return SPAM_OK;
  }
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
