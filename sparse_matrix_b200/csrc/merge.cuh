// merge.cuh — the "merge" bin: SpGEMM rows as a k-way merge of sorted runs, one thread per row.
//
// When B's rows are sorted by column (CsrMatrix<T, true>; everything From<DokMatrix> builds,
// spam_csr/src/lib.rs:315-334), row i of C is the union of the runs  a_ik * B[k,:]  over the entries
// k of A's row i, each run already ordered.  For short A rows (<= MERGE_K entries) a thread keeps the
// run heads (position, end, current column) in registers and repeatedly takes the smallest column:
//   * no hash table, no shared-memory atomics, no per-row sort: the output comes out ordered
//     (the reference's B2=true branch, mul_hash.rs:164-175);
//   * runs are visited in A-row storage order and equal columns are folded first-stored-then-added
//     with separate mul and add, i.e. exactly the reference's accumulation order
//     (mul_hash.rs:145-162): sums are bit-identical to the reference, floats included.
// HBM-bound by design: A and the B rows it touches are read once (B mostly from L1/L2 for banded
// matrices), C is written once through a shared-memory transpose so that each warp stores the
// contiguous output of its 32 consecutive rows.
#pragma once
#include "common.cuh"

namespace {

constexpr u32 INF_COL = 0xFFFFFFFFu;

// One pass over a matrix the first time it is used (the result is cached with the matrix):
//  * are all rows strictly increasing by column?  (invariant6 with IS_SORTED, lib.rs:69-77)
//  * longest row; how far the entries of a short row are apart (picks the prefetching merge kernels)
//  * the invariants the kernels rely on for memory safety (lib.rs:47-81): row_ptr starts at 0, is monotone and
//    ends at nnz (invariants 3, 4, 7), every column index < cols (invariant 5).  The reference panics safely
//    on a bad matrix; here it is SPAM_EINVAL / SPAM_EINDEX instead of a device fault.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_rows_sorted(u64 m, u64 nnz, u64 cols, const u64* __restrict__ ptr,
                                                       const u32* __restrict__ idx, Counters* cnt) {
  const int lane = threadIdx.x & 31;
  const u64 row = (u64)blockIdx.x * BLOCK + threadIdx.x;
  bool valid = row < m;
  u64 lo = 0, hi = 0;
  u32 inval = 0;
  if (valid) {
    lo = ptr[row]; hi = ptr[row + 1];
    if (lo > hi || hi > nnz || (row == 0 && lo != 0) || (row + 1 == m && hi != nnz)) { inval |= 1u; valid = false; lo = hi = 0; }
  }
  bool bad = false;
  u64 spread = 0;
  if (valid && hi - lo <= 32) {
    u32 prev = 0, cmin = 0xFFFFFFFFu, cmax = 0;
    for (u64 e = lo; e < hi; ++e) {
      const u32 c = idx[e];
      if (c >= cols) inval |= 2u;
      if (e > lo) bad |= prev >= c;
      prev = c;
      cmin = min(cmin, c); cmax = max(cmax, c);
    }
    if (hi > lo) spread = cmax - cmin;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) spread += __shfl_xor_sync(0xffffffffu, spread, d);
  if (lane == 0 && spread) atomicAdd(&cnt->spread_sum, (ull)spread);
  unsigned longmask = __ballot_sync(0xffffffffu, valid && hi - lo > 32);
  while (longmask) {
    const int src = __ffs(longmask) - 1;
    longmask &= longmask - 1;
    const u64 l = __shfl_sync(0xffffffffu, lo, src), hh = __shfl_sync(0xffffffffu, hi, src);
    for (u64 e = l + lane; e < hh; e += 32) {
      const u32 c = idx[e];
      if (c >= cols) inval |= 2u;
      if (e > l) bad |= idx[e - 1] >= c;
    }
  }
  if (bad) atomicOr(&cnt->unsorted, 1u);
  if (inval) atomicOr(&cnt->invalid, inval);
  u64 len = hi - lo;
  u32 l32 = len > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)len;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) l32 = max(l32, __shfl_xor_sync(0xffffffffu, l32, d));
  if (lane == 0 && l32) atomicMax(&cnt->max_rowlen, l32);
}

// SYMBOLIC merge: count distinct columns; also histograms the numeric bin of each row so that the
// separate k_num_bin_count pass is skipped when every row is in this bin.
template <int K, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_sym_merge(u32 n, const u32* __restrict__ perm,
                                                     const u64* __restrict__ a_ptr, const u32* __restrict__ a_col,
                                                     const u64* __restrict__ b_ptr, const u32* __restrict__ b_col,
                                                     const u32* __restrict__ flop, u32* __restrict__ row_nnz,
                                                     Counters* cnt) {
  __shared__ u32 s_hist[NBINS];
  const int tid = threadIdx.x;
  if (tid < NBINS) s_hist[tid] = 0;
  __syncthreads();
  const u32 i = blockIdx.x * BLOCK + tid;
  if (i < n) {
    const u32 row = perm ? perm[i] : i;
    const u64 alo = a_ptr[row];
    const u32 k = (u32)(a_ptr[row + 1] - alo);
    u32 pos[K], end[K], col[K];
#pragma unroll
    for (int h = 0; h < K; ++h) {
      pos[h] = 0; end[h] = 0; col[h] = INF_COL;
      if (h < k) {
        const u32 kk = a_col[alo + h];
        pos[h] = (u32)b_ptr[kk];
        end[h] = (u32)b_ptr[kk + 1];
      }
    }
#pragma unroll
    for (int h = 0; h < K; ++h)
      if (pos[h] < end[h]) col[h] = b_col[pos[h]];
    u32 z = 0;
    for (;;) {
      u32 cmin = col[0];
#pragma unroll
      for (int h = 1; h < K; ++h) cmin = min(cmin, col[h]);
      if (cmin == INF_COL) break;
      ++z;
#pragma unroll
      for (int h = 0; h < K; ++h) {
        if (col[h] == cmin) {
          ++pos[h];
          col[h] = (pos[h] < end[h]) ? b_col[pos[h]] : INF_COL;
        }
      }
    }
    row_nnz[row] = z;  // mul_hash.rs:95
    atomicAdd(&s_hist[num_bin_of(z, flop[row], k, MODE_MERGE)], 1u);
  }
  __syncthreads();
  if (tid < NBINS && s_hist[tid]) atomicAdd(&cnt->num_bins[tid], s_hist[tid]);
}

// FUSED flop count + symbolic merge (used when B's rows are sorted): one pass over A instead of two.
// Every row gets its flop count (rows_to_threads, mul_hash.rs:39-50) and its symbolic bin; rows that
// qualify for the merge bin (len(A row) <= K <= MERGE_K, flop <= MERGE_FLOP_MAX) are counted on the
// spot — the B row extents just loaded for the flop count are exactly the run heads the merge needs.
// For stencil-like matrices every row qualifies and no other symbolic kernel runs.
template <int K, int BLOCK, int PF = 0>
__global__ void __launch_bounds__(BLOCK) k_flop_sym_merge(u64 m, u64 b_rows, const u64* __restrict__ a_ptr,
                                                          const u32* __restrict__ a_col,
                                                          const u64* __restrict__ b_ptr,
                                                          const u32* __restrict__ b_col, u32* __restrict__ flop_out,
                                                          u32* __restrict__ row_nnz, Counters* cnt) {
  __shared__ u32 s_sym[NBINS], s_num[NBINS];
  __shared__ ull s_total;
  __shared__ u32 s_max;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < NBINS) { s_sym[tid] = 0; s_num[tid] = 0; }
  if (tid == 0) { s_total = 0; s_max = 0; }
  __syncthreads();
  const u64 row = (u64)blockIdx.x * BLOCK + tid;
  const bool valid = row < m;
  u64 lo = 0, hi = 0;
  if (valid) { lo = a_ptr[row]; hi = a_ptr[row + 1]; }
  const u64 len = hi - lo;
  u64 f = 0;
  bool bad = false, merged = false;
  if (valid && len <= (u64)K) {
    u32 pos[K], end[K], col[K];
#pragma unroll
    for (int h = 0; h < K; ++h) {
      pos[h] = 0; end[h] = 0; col[h] = INF_COL;
      if ((u64)h < len) {
        const u32 kk = a_col[lo + h];
        if (kk < b_rows) { pos[h] = (u32)b_ptr[kk]; end[h] = (u32)b_ptr[kk + 1]; } else bad = true;
      }
      f += end[h] - pos[h];
    }
    if (f <= MERGE_FLOP_MAX) {
      merged = true;
      u32 ncol[PF ? K : 1];
#pragma unroll
      for (int h = 0; h < K; ++h)
        if (pos[h] < end[h]) {
          col[h] = b_col[pos[h]];
          if (PF) ncol[h] = (pos[h] + 1 < end[h]) ? b_col[pos[h] + 1] : INF_COL;
        }
      u32 z = 0;
      for (;;) {
        u32 cmin = col[0];
#pragma unroll
        for (int h = 1; h < K; ++h) cmin = min(cmin, col[h]);
        if (cmin == INF_COL) break;
        ++z;
#pragma unroll
        for (int h = 0; h < K; ++h) {
          if (col[h] == cmin) {
            ++pos[h];
            if (PF) {
              col[h] = ncol[h];   // INF_COL when the run is finished
              ncol[h] = (pos[h] + 1 < end[h]) ? b_col[pos[h] + 1] : INF_COL;
            } else {
              col[h] = (pos[h] < end[h]) ? b_col[pos[h]] : INF_COL;
            }
          }
        }
      }
      row_nnz[row] = z;  // mul_hash.rs:95
      atomicAdd(&s_num[MERGE_BIN], 1u);
    }
  } else if (valid && len <= 32) {
    for (u64 e = lo; e < hi; ++e) {
      const u32 kk = a_col[e];
      if (kk < b_rows) f += b_ptr[kk + 1] - b_ptr[kk]; else bad = true;
    }
  }
  unsigned longmask = __ballot_sync(0xffffffffu, valid && len > 32);
  while (longmask) {
    const int src = __ffs(longmask) - 1;
    longmask &= longmask - 1;
    const u64 l = __shfl_sync(0xffffffffu, lo, src), hh = __shfl_sync(0xffffffffu, hi, src);
    u64 part = 0;
    for (u64 e = l + lane; e < hh; e += 32) {
      const u32 kk = a_col[e];
      if (kk < b_rows) part += b_ptr[kk + 1] - b_ptr[kk]; else bad = true;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (lane == src) f = part;
  }
  if (bad) atomicOr(&cnt->error, 1u);
  if (valid) {
    const u32 fs = f > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)f;
    flop_out[row] = fs;
    if (merged) {
      atomicAdd(&s_sym[MERGE_BIN], 1u);
    } else {
      // a row with len <= MERGE_K but len > K cannot occur: K >= min(MERGE_K, longest row of A)
      const u32 alen = len > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)len;
      atomicAdd(&s_sym[sym_bin_of(fs, alen, false)], 1u);
      atomicMax(&s_max, fs);
    }
  }
  u64 t = f;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
  if (lane == 0 && t) atomicAdd(&s_total, (ull)t);
  __syncthreads();
  if (tid < NBINS) {
    if (s_sym[tid]) atomicAdd(&cnt->sym_bins[tid], s_sym[tid]);
    if (s_num[tid]) atomicAdd(&cnt->num_bins[tid], s_num[tid]);
  }
  if (tid == 0) {
    if (s_total) atomicAdd(&cnt->total_flops, s_total);
    if (s_max) atomicMax(&cnt->max_flop, s_max);
  }
}

// The same fused pass with the block's B window staged in shared memory (see k_num_merge_win below for the
// scheme): only col_idx is needed here, MERGE_WCAP entries are 10 KB.
constexpr int MERGE_WCAP = 2560;

// Block-wide union of the K per-head entry ranges [s_lo[h], s_hi[h]) into at most K windows laid out back to back
// in a staging buffer of MERGE_WCAP entries.  Called by one thread.  Windows start and end on multiples of 4
// entries (16 bytes of col_idx, 16/32 of the values) so that they can be fetched with cp.async.bulk.  Returns the
// number of windows, 0 if they do not fit.  s_shift[h] = (first entry of the head's window) - (offset of that
// window in the buffer).
template <int K>
__device__ __forceinline__ u32 merge_window_plan(const u32* s_lo, const u32* s_hi, u32* s_shift, u32* s_wlo,
                                                 u32* s_wn, u32* s_wbase) {
  int ord[K];
  int nh = 0;
  for (int h = 0; h < K; ++h)
    if (s_lo[h] != 0xFFFFFFFFu) {
      int j = nh++;
      while (j > 0 && s_lo[ord[j - 1]] > s_lo[h]) { ord[j] = ord[j - 1]; --j; }
      ord[j] = h;
    }
  u32 nwin = 0, total = 0;
  for (int j = 0; j < nh; ++j) {
    const int h = ord[j];
    if (s_hi[h] > 0xFFFFFFF0u) return 0;
    const u32 lo = s_lo[h] & ~3u, hi = (s_hi[h] + 3u) & ~3u;
    if (hi - lo > (u32)MERGE_WCAP) return 0;
    if (nwin && lo <= s_wlo[nwin - 1] + s_wn[nwin - 1]) {
      const u32 cur_hi = s_wlo[nwin - 1] + s_wn[nwin - 1];
      if (hi > cur_hi) { total += hi - cur_hi; s_wn[nwin - 1] = hi - s_wlo[nwin - 1]; }
    } else {
      s_wlo[nwin] = lo; s_wn[nwin] = hi - lo; s_wbase[nwin] = total;
      total += hi - lo;
      ++nwin;
    }
    s_shift[h] = s_wlo[nwin - 1] - s_wbase[nwin - 1];
    if (total > (u32)MERGE_WCAP) return 0;
  }
  return nwin;
}

// Fetch the planned windows of one or two parallel arrays (col_idx, values) into the staging buffers: one thread
// issues the bulk copies, every thread copies the at most 3 entries at the very end of the arrays that a 16-byte
// copy would overrun, all wait.
template <class V, int BLOCK, bool WITH_VAL>
__device__ __forceinline__ void merge_window_fetch(u32 nwin, const u32* s_wlo, const u32* s_wn, const u32* s_wbase,
                                                   u64 b_nnz, const u32* __restrict__ b_col,
                                                   const V* __restrict__ b_val, u32* wk, V* wv, u64* bar) {
  const int tid = threadIdx.x;
  const u32 nnz4 = (u32)(b_nnz & ~3ull), nnz = (u32)b_nnz;
  constexpr u32 per = WITH_VAL ? 4u + (u32)sizeof(V) : 4u;
  if (tid == 0) {
    u32 bytes = 0;
    for (u32 w = 0; w < nwin; ++w) {
      const u32 wl = s_wlo[w], we = min(wl + s_wn[w], nnz4);
      if (we > wl) bytes += (we - wl) * per;
    }
    mbar_expect_tx(bar, bytes);
    for (u32 w = 0; w < nwin; ++w) {
      const u32 wl = s_wlo[w], we = min(wl + s_wn[w], nnz4), wb = s_wbase[w];
      if (we > wl) {
        bulk_g2s(wk + wb, b_col + wl, (we - wl) * 4u, bar);
        if (WITH_VAL) bulk_g2s(wv + wb, b_val + wl, (we - wl) * (u32)sizeof(V), bar);
      }
    }
  }
  if (nnz4 != nnz) {
    for (u32 w = 0; w < nwin; ++w) {
      const u32 wl = s_wlo[w], wb = s_wbase[w];
      const u32 e1 = min(wl + s_wn[w], nnz);
      for (u32 e = max(wl, nnz4) + tid; e < e1; e += BLOCK) {
        wk[wb + (e - wl)] = b_col[e];
        if (WITH_VAL) wv[wb + (e - wl)] = b_val[e];
      }
    }
  }
  mbar_wait(bar, 0);
  __syncthreads();
}

template <int K, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_flop_sym_merge_win(u64 m, u64 b_rows, const u64* __restrict__ a_ptr,
                                                              const u32* __restrict__ a_col,
                                                              const u64* __restrict__ b_ptr,
                                                              const u32* __restrict__ b_col, u64 b_nnz,
                                                              u32* __restrict__ flop_out, u32* __restrict__ row_nnz,
                                                              Counters* cnt) {
  __shared__ u32 s_sym[NBINS], s_num[NBINS];
  __shared__ ull s_total;
  __shared__ u32 s_max;
  __shared__ __align__(16) u32 wk[MERGE_WCAP];
  __shared__ __align__(8) u64 s_bar;
  __shared__ u32 s_lo[K], s_hi[K], s_shift[K], s_wlo[K], s_wn[K], s_wbase[K], s_nwin;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < NBINS) { s_sym[tid] = 0; s_num[tid] = 0; }
  if (tid == 0) { s_total = 0; s_max = 0; mbar_init(&s_bar, 1); }
  if (tid < K) { s_lo[tid] = 0xFFFFFFFFu; s_hi[tid] = 0; }
  __syncthreads();
  const u64 row = (u64)blockIdx.x * BLOCK + tid;
  const bool valid = row < m;
  u64 lo = 0, hi = 0;
  if (valid) { lo = a_ptr[row]; hi = a_ptr[row + 1]; }
  const u64 len = hi - lo;
  u64 f = 0;
  bool bad = false, merged = false;
  u32 pos[K], end[K], col[K];
#pragma unroll
  for (int h = 0; h < K; ++h) { pos[h] = 0; end[h] = 0; col[h] = INF_COL; }
  if (valid && len <= (u64)K) {
#pragma unroll
    for (int h = 0; h < K; ++h) {
      if ((u64)h < len) {
        const u32 kk = a_col[lo + h];
        if (kk < b_rows) { pos[h] = (u32)b_ptr[kk]; end[h] = (u32)b_ptr[kk + 1]; } else bad = true;
      }
      f += end[h] - pos[h];
    }
    merged = f <= MERGE_FLOP_MAX;
  } else if (valid && len <= 32) {
    for (u64 e = lo; e < hi; ++e) {
      const u32 kk = a_col[e];
      if (kk < b_rows) f += b_ptr[kk + 1] - b_ptr[kk]; else bad = true;
    }
  }
  unsigned longmask = __ballot_sync(0xffffffffu, valid && len > 32);
  while (longmask) {
    const int src = __ffs(longmask) - 1;
    longmask &= longmask - 1;
    const u64 l = __shfl_sync(0xffffffffu, lo, src), hh = __shfl_sync(0xffffffffu, hi, src);
    u64 part = 0;
    for (u64 e = l + lane; e < hh; e += 32) {
      const u32 kk = a_col[e];
      if (kk < b_rows) part += b_ptr[kk + 1] - b_ptr[kk]; else bad = true;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (lane == src) f = part;
  }
  // window of the rows counted here
#pragma unroll
  for (int h = 0; h < K; ++h) {
    const bool has = merged && pos[h] < end[h];
    const u32 wl = __reduce_min_sync(0xffffffffu, has ? pos[h] : 0xFFFFFFFFu);
    const u32 wh = __reduce_max_sync(0xffffffffu, has ? end[h] : 0u);
    if (lane == 0 && wl != 0xFFFFFFFFu) { atomicMin(&s_lo[h], wl); atomicMax(&s_hi[h], wh); }
  }
  __syncthreads();
  if (tid == 0) s_nwin = merge_window_plan<K>(s_lo, s_hi, s_shift, s_wlo, s_wn, s_wbase);
  __syncthreads();
  const u32 nwin = s_nwin;
  if (nwin) merge_window_fetch<u32, BLOCK, false>(nwin, s_wlo, s_wn, s_wbase, b_nnz, b_col, nullptr, wk, nullptr, &s_bar);
  if (merged) {
    u32 z = 0;
    if (nwin) {
#pragma unroll
      for (int h = 0; h < K; ++h) {
        const u32 sh = s_shift[h];
        if (pos[h] < end[h]) { pos[h] -= sh; end[h] -= sh; col[h] = wk[pos[h]]; }
      }
      for (;;) {
        u32 cmin = col[0];
#pragma unroll
        for (int h = 1; h < K; ++h) cmin = min(cmin, col[h]);
        if (cmin == INF_COL) break;
        ++z;
#pragma unroll
        for (int h = 0; h < K; ++h) {
          if (col[h] == cmin) {
            ++pos[h];
            col[h] = (pos[h] < end[h]) ? wk[pos[h]] : INF_COL;
          }
        }
      }
    } else {
#pragma unroll
      for (int h = 0; h < K; ++h)
        if (pos[h] < end[h]) col[h] = b_col[pos[h]];
      for (;;) {
        u32 cmin = col[0];
#pragma unroll
        for (int h = 1; h < K; ++h) cmin = min(cmin, col[h]);
        if (cmin == INF_COL) break;
        ++z;
#pragma unroll
        for (int h = 0; h < K; ++h) {
          if (col[h] == cmin) {
            ++pos[h];
            col[h] = (pos[h] < end[h]) ? b_col[pos[h]] : INF_COL;
          }
        }
      }
    }
    row_nnz[row] = z;  // mul_hash.rs:95
    atomicAdd(&s_num[MERGE_BIN], 1u);
  }
  if (bad) atomicOr(&cnt->error, 1u);
  if (valid) {
    const u32 fs = f > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)f;
    flop_out[row] = fs;
    if (merged) {
      atomicAdd(&s_sym[MERGE_BIN], 1u);
    } else {
      const u32 alen = len > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)len;
      atomicAdd(&s_sym[sym_bin_of(fs, alen, false)], 1u);
      atomicMax(&s_max, fs);
    }
  }
  u64 t = f;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
  if (lane == 0 && t) atomicAdd(&s_total, (ull)t);
  __syncthreads();
  if (tid < NBINS) {
    if (s_sym[tid]) atomicAdd(&cnt->sym_bins[tid], s_sym[tid]);
    if (s_num[tid]) atomicAdd(&cnt->num_bins[tid], s_num[tid]);
  }
  if (tid == 0) {
    if (s_total) atomicAdd(&cnt->total_flops, s_total);
    if (s_max) atomicMax(&cnt->max_flop, s_max);
  }
}

// NUMERIC merge.  Outputs are staged CH at a time in a small [slot][thread] shared-memory tile and
// flushed by the whole warp (8 lanes per row, 4 rows per store instruction: a row's 8 columns are one
// full 32-byte sector, its 8 values two).  Keeping the tile small matters: shared memory is carved
// out of the same 228 KB as L1, and the run heads re-read B through L1 (with 16-slot staging the L1
// hit rate fell to 19% and the kernel became L2-bandwidth bound — profiles/r01_poisson_v1_merge.txt);
// with 4-slot staging the half-sector column stores were written back early by L2 (DRAM writes 773 MB
// for a 654 MB result — profiles/r01_poisson_v2_fused.txt).
// Tile strides are padded so that the owner's writes and the flush reads are bank-conflict free.
constexpr int MERGE_CH = 8;

template <class V, int BLOCK>
struct MergeTile {
  static constexpr int STRIDE_K = BLOCK + 4;                        // u32 words: (4q + r) mod 32 distinct
  static constexpr int STRIDE_V = sizeof(V) == 8 ? BLOCK + 2 : BLOCK + 4;  // 8-byte words: (2q + r) mod 16
  static constexpr size_t bytes = (size_t)MERGE_CH * (STRIDE_V * sizeof(V) + STRIDE_K * 4);
};

// (Compiled for 12 resident blocks per SM — 40 registers instead of 54 for K = 6, a few spilled bytes — the kernel
// ran 0.479 ms instead of 0.380 ms on Poisson 2048^2: occupancy is not what limits it.)
// PF = 1: a head's value is loaded together with its column (not when it is consumed); PF = 2: also the column
// after the head is already in a register, so that the next comparison does not wait for a load.
// A grid smaller than the number of row blocks walks them with a stride (SPAM_MERGE_PERSIST: 5..12 blocks per SM measured
// 0.40..0.55 ms against 0.38 ms for one block per 128 rows: the hardware's block scheduler balances better).
template <class V, int K, int BLOCK, int PF = 0>
__global__ void __launch_bounds__(BLOCK) k_num_merge(u32 n, const u32* __restrict__ perm,
                                                     const u64* __restrict__ a_ptr, const u32* __restrict__ a_col,
                                                     const V* __restrict__ a_val, const u64* __restrict__ b_ptr,
                                                     const u32* __restrict__ b_col, const V* __restrict__ b_val,
                                                     const u64* __restrict__ c_ptr, u32* __restrict__ c_col,
                                                     V* __restrict__ c_val) {
  using Tile = MergeTile<V, BLOCK>;
  extern __shared__ __align__(16) unsigned char sm_merge[];
  V* sv = reinterpret_cast<V*>(sm_merge);                               // [MERGE_CH][STRIDE_V]
  u32* sk = reinterpret_cast<u32*>(sv + MERGE_CH * Tile::STRIDE_V);     // [MERGE_CH][STRIDE_K]
  const int tid = threadIdx.x, lane = tid & 31;
  // one pass when the grid covers all rows; a smaller (persistent) grid walks row blocks blockIdx.x, + gridDim.x, ...
  for (u32 blk = blockIdx.x; (u64)blk * BLOCK < (u64)n; blk += gridDim.x) {
  const u32 i = blk * BLOCK + tid;
  u64 c0 = 0;
  u32 z = 0, row = 0;
  if (i < n) {
    row = perm ? perm[i] : i;
    c0 = c_ptr[row];
    z = (u32)(c_ptr[row + 1] - c0);
  }
  u32 pos[K], end[K], col[K];
  u32 ncol[PF >= 2 ? K : 1];
  V av[K], bvv[PF >= 1 ? K : 1];
#pragma unroll
  for (int h = 0; h < K; ++h) { pos[h] = 0; end[h] = 0; col[h] = INF_COL; av[h] = Num<V>::zero(); }
#pragma unroll
  for (int h = 0; h < (PF >= 1 ? K : 1); ++h) bvv[h] = Num<V>::zero();
#pragma unroll
  for (int h = 0; h < (PF >= 2 ? K : 1); ++h) ncol[h] = INF_COL;
  if (z > 0) {
    const u64 alo = a_ptr[row];
    const u32 k = (u32)(a_ptr[row + 1] - alo);
#pragma unroll
    for (int h = 0; h < K; ++h) {
      if (h < k) {
        const u32 kk = a_col[alo + h];
        av[h] = a_val[alo + h];
        pos[h] = (u32)b_ptr[kk];
        end[h] = (u32)b_ptr[kk + 1];
      }
    }
#pragma unroll
    for (int h = 0; h < K; ++h)
      if (pos[h] < end[h]) {
        col[h] = b_col[pos[h]];
        if (PF >= 1) bvv[h] = b_val[pos[h]];
        if (PF >= 2) ncol[h] = (pos[h] + 1 < end[h]) ? b_col[pos[h] + 1] : INF_COL;
      }
  }
  const int wbase = tid & ~31, rsub = lane >> 3, q = lane & 7;
  for (u32 t0 = 0; __any_sync(0xffffffffu, t0 < z); t0 += MERGE_CH) {
#pragma unroll
    for (int c = 0; c < MERGE_CH; ++c) {
      if (t0 + c < z) {
        u32 cmin = col[0];
#pragma unroll
        for (int h = 1; h < K; ++h) cmin = min(cmin, col[h]);
        V acc = Num<V>::zero();
        bool first = true;
#pragma unroll
        for (int h = 0; h < K; ++h) {  // A-row storage order
          if (col[h] == cmin && cmin != INF_COL) {
            const V p = Num<V>::mul(av[h], PF >= 1 ? bvv[h] : b_val[pos[h]]);
            acc = first ? p : Num<V>::add(acc, p);  // first product stored, not added to 0
            first = false;
            ++pos[h];
            if (PF >= 2) {
              col[h] = ncol[h];
              if (pos[h] < end[h]) {
                bvv[h] = b_val[pos[h]];
                ncol[h] = (pos[h] + 1 < end[h]) ? b_col[pos[h] + 1] : INF_COL;
              }
            } else if (PF == 1) {
              if (pos[h] < end[h]) { col[h] = b_col[pos[h]]; bvv[h] = b_val[pos[h]]; } else col[h] = INF_COL;
            } else {
              col[h] = (pos[h] < end[h]) ? b_col[pos[h]] : INF_COL;
            }
          }
        }
        sk[c * Tile::STRIDE_K + tid] = cmin;
        sv[c * Tile::STRIDE_V + tid] = acc;
      }
    }
    __syncwarp();
    // flush: lane (rsub, q) stores entry t0+q of the warp's row it*4+rsub
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int src = it * 4 + rsub;
      const u32 zr = __shfl_sync(0xffffffffu, z, src);
      const u64 c0r = __shfl_sync(0xffffffffu, c0, src);
      const u32 tt = t0 + q;
      if (tt < zr) {
        c_col[c0r + tt] = sk[q * Tile::STRIDE_K + wbase + src];
        c_val[c0r + tt] = sv[q * Tile::STRIDE_V + wbase + src];
      }
    }
    __syncwarp();
  }
  }
}

template <class V, int BLOCK>
constexpr size_t num_merge_smem() { return MergeTile<V, BLOCK>::bytes; }

// NUMERIC merge with the B WINDOW of the block staged in shared memory.
//
// In k_num_merge every step of a run head is a dependent global load (col, then val): ~50 per row for the 5-point
// stencil, served by L1/L2 but each a 200-600 cycle round trip, and the kernel sits at 38% of the HBM rate waiting
// for them (profiles/r02_poisson_merge_full.txt).  For banded matrices the B rows a block of consecutive A rows
// touches are a few contiguous row ranges: head h of the block's threads walks B rows [min_h, max_h], i.e. the
// contiguous ENTRY range [b_ptr[min_h], b_ptr[max_h + 1]) of col_idx / values.  The block computes those K entry
// ranges (warp REDUX + shared atomics), thread 0 merges overlapping ones (heads i-1, i, i+1 of a stencil walk the
// same rows), and when the union fits MERGE_WCAP entries one thread fetches it into shared memory with 1-D bulk copies (cp.async.bulk,
// SASS UBLKCP, completion on an mbarrier) — B is then read from HBM/L2 once per block, and the merge loop runs on shared memory.
// Blocks whose windows do not fit (scattered rows) take the global-memory loop of k_num_merge unchanged, so the
// kernel is correct for every input; the products, their order and the rounding are the same in both branches.

template <class V, int BLOCK>
constexpr size_t num_merge_win_smem() { return MergeTile<V, BLOCK>::bytes + (size_t)MERGE_WCAP * (sizeof(V) + 4); }

template <class V, int K, int BLOCK, bool STAGED>
__device__ __forceinline__ void merge_rows_out(u32 (&pos)[K], u32 (&end)[K], u32 (&col)[K], V (&av)[K], u32 z, u64 c0,
                                               const u32* __restrict__ src_col, const V* __restrict__ src_val,
                                               V* sv, u32* sk, u32* __restrict__ c_col, V* __restrict__ c_val) {
  using Tile = MergeTile<V, BLOCK>;
  const int tid = threadIdx.x, lane = tid & 31;
  const int wbase = tid & ~31, rsub = lane >> 3, q = lane & 7;
  for (u32 t0 = 0; __any_sync(0xffffffffu, t0 < z); t0 += MERGE_CH) {
#pragma unroll
    for (int c = 0; c < MERGE_CH; ++c) {
      if (t0 + c < z) {
        u32 cmin = col[0];
#pragma unroll
        for (int h = 1; h < K; ++h) cmin = min(cmin, col[h]);
        V acc = Num<V>::zero();
        bool first = true;
#pragma unroll
        for (int h = 0; h < K; ++h) {  // A-row storage order
          if (col[h] == cmin && cmin != INF_COL) {
            const V p = Num<V>::mul(av[h], src_val[pos[h]]);
            acc = first ? p : Num<V>::add(acc, p);  // first product stored, not added to 0
            first = false;
            ++pos[h];
            col[h] = (pos[h] < end[h]) ? src_col[pos[h]] : INF_COL;
          }
        }
        sk[c * Tile::STRIDE_K + tid] = cmin;
        sv[c * Tile::STRIDE_V + tid] = acc;
      }
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int src = it * 4 + rsub;
      const u32 zr = __shfl_sync(0xffffffffu, z, src);
      const u64 c0r = __shfl_sync(0xffffffffu, c0, src);
      const u32 tt = t0 + q;
      if (tt < zr) {
        c_col[c0r + tt] = sk[q * Tile::STRIDE_K + wbase + src];
        c_val[c0r + tt] = sv[q * Tile::STRIDE_V + wbase + src];
      }
    }
    __syncwarp();
  }
}

template <class V, int K, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_num_merge_win(u32 n, const u32* __restrict__ perm,
                                                         const u64* __restrict__ a_ptr, const u32* __restrict__ a_col,
                                                         const V* __restrict__ a_val, const u64* __restrict__ b_ptr,
                                                         const u32* __restrict__ b_col, const V* __restrict__ b_val,
                                                         u64 b_nnz, const u64* __restrict__ c_ptr,
                                                         u32* __restrict__ c_col, V* __restrict__ c_val) {
  using Tile = MergeTile<V, BLOCK>;
  extern __shared__ __align__(16) unsigned char sm_merge[];
  V* sv = reinterpret_cast<V*>(sm_merge);                               // [MERGE_CH][STRIDE_V]
  u32* sk = reinterpret_cast<u32*>(sv + MERGE_CH * Tile::STRIDE_V);     // [MERGE_CH][STRIDE_K]
  V* wv = reinterpret_cast<V*>(sm_merge + Tile::bytes);                 // [MERGE_WCAP] staged B values
  u32* wk = reinterpret_cast<u32*>(wv + MERGE_WCAP);                    // [MERGE_WCAP] staged B columns
  __shared__ u32 s_lo[K], s_hi[K], s_shift[K];     // per head: entry range walked by the block; global - smem offset
  __shared__ u32 s_wlo[K], s_wn[K], s_wbase[K];    // merged windows: first entry, length, offset in wv / wk
  __shared__ u32 s_nwin;                           // 0: not staged
  __shared__ __align__(8) u64 s_bar;
  const int tid = threadIdx.x, lane = tid & 31;
  const u32 i = blockIdx.x * BLOCK + tid;
  if (tid < K) { s_lo[tid] = 0xFFFFFFFFu; s_hi[tid] = 0; }
  if (tid == 0) mbar_init(&s_bar, 1);
  u64 c0 = 0;
  u32 z = 0, row = 0;
  if (i < n) {
    row = perm ? perm[i] : i;
    c0 = c_ptr[row];
    z = (u32)(c_ptr[row + 1] - c0);
  }
  u32 pos[K], end[K], col[K];
  V av[K];
#pragma unroll
  for (int h = 0; h < K; ++h) { pos[h] = 0; end[h] = 0; col[h] = INF_COL; av[h] = Num<V>::zero(); }
  if (z > 0) {
    const u64 alo = a_ptr[row];
    const u32 k = (u32)(a_ptr[row + 1] - alo);
#pragma unroll
    for (int h = 0; h < K; ++h) {
      if (h < k) {
        const u32 kk = a_col[alo + h];
        av[h] = a_val[alo + h];
        pos[h] = (u32)b_ptr[kk];
        end[h] = (u32)b_ptr[kk + 1];
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int h = 0; h < K; ++h) {
    const bool has = pos[h] < end[h];
    const u32 wl = __reduce_min_sync(0xffffffffu, has ? pos[h] : 0xFFFFFFFFu);
    const u32 wh = __reduce_max_sync(0xffffffffu, has ? end[h] : 0u);
    if (lane == 0 && wl != 0xFFFFFFFFu) { atomicMin(&s_lo[h], wl); atomicMax(&s_hi[h], wh); }
  }
  __syncthreads();
  if (tid == 0) s_nwin = merge_window_plan<K>(s_lo, s_hi, s_shift, s_wlo, s_wn, s_wbase);
  __syncthreads();
  const u32 nwin = s_nwin;
  if (nwin) {
    merge_window_fetch<V, BLOCK, true>(nwin, s_wlo, s_wn, s_wbase, b_nnz, b_col, b_val, wk, wv, &s_bar);
#pragma unroll
    for (int h = 0; h < K; ++h) {
      const u32 sh = s_shift[h];
      if (pos[h] < end[h]) { pos[h] -= sh; end[h] -= sh; col[h] = wk[pos[h]]; }
    }
    merge_rows_out<V, K, BLOCK, true>(pos, end, col, av, z, c0, wk, wv, sv, sk, c_col, c_val);
  } else {
#pragma unroll
    for (int h = 0; h < K; ++h)
      if (pos[h] < end[h]) col[h] = b_col[pos[h]];
    merge_rows_out<V, K, BLOCK, false>(pos, end, col, av, z, c0, b_col, b_val, sv, sk, c_col, c_val);
  }
}

// ONE-PASS product for matrices whose every row is a merge row BY THE CACHED STATISTICS (longest row of A <= K,
// longest row of A x longest row of B <= MERGE_FLOP_MAX), so that no flop / binning pass is needed to know it and
// nnz(A) x longest row of B bounds nnz(C) for the allocation.  Per block of BLOCK consecutive rows:
//   1. load the run heads, count the products (flops) and the distinct columns (the symbolic merge, mul_hash.rs:84-99),
//   2. block-wide exclusive scan of the counts; the block's total goes into a decoupled look-back over the blocks
//      (tiles handed out by an atomic counter, {flag,value} in one 64-bit word as in scan.cu) and comes back as the
//      block's first position in C (the row_ptr scan, lib.rs:267-274),
//   3. write row_ptr(C), run the numeric merge of k_num_merge from the same run heads (B's rows are now in L1).
// Against the two-phase pipeline (k_flop_sym_merge, k_scan_lookback, host sync for nnz(C), allocation, k_num_merge)
// this saves one pass over A, the row_nnz / flop / row_ptr round trips, the scan launch and the mid-product host
// synchronisation; the host learns nnz(C) after the kernel.  Same products, same order: bit-identical.
template <class V, int K, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_merge_onepass(u32 m, u64 b_rows, const u64* __restrict__ a_ptr,
                                                         const u32* __restrict__ a_col, const V* __restrict__ a_val,
                                                         const u64* __restrict__ b_ptr, const u32* __restrict__ b_col,
                                                         const V* __restrict__ b_val, u64* __restrict__ c_ptr,
                                                         u32* __restrict__ c_col, V* __restrict__ c_val,
                                                         volatile u64* state, u32* tile_counter, Counters* cnt) {
  using Tile = MergeTile<V, BLOCK>;
  extern __shared__ __align__(16) unsigned char sm_merge[];
  V* sv = reinterpret_cast<V*>(sm_merge);                               // [MERGE_CH][STRIDE_V]
  u32* sk = reinterpret_cast<u32*>(sv + MERGE_CH * Tile::STRIDE_V);     // [MERGE_CH][STRIDE_K]
  __shared__ u32 s_tile, s_warp[BLOCK / 32];
  __shared__ u64 s_base;
  constexpr u64 F_AGG = 1ull << 62, F_PFX = 2ull << 62, F_MASK = (1ull << 62) - 1;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  __syncthreads();
  const u32 tile = s_tile;
  const u32 row = tile * BLOCK + tid;
  u32 pos0[K], pos[K], end[K], col[K];
  V av[K];
#pragma unroll
  for (int h = 0; h < K; ++h) { pos0[h] = 0; end[h] = 0; col[h] = INF_COL; av[h] = Num<V>::zero(); }
  bool bad = false;
  u32 f = 0;
  if (row < m) {
    const u64 alo = a_ptr[row];
    const u32 k = (u32)(a_ptr[row + 1] - alo);
#pragma unroll
    for (int h = 0; h < K; ++h) {
      if (h < k) {
        const u32 kk = a_col[alo + h];
        av[h] = a_val[alo + h];
        if (kk < b_rows) { pos0[h] = (u32)b_ptr[kk]; end[h] = (u32)b_ptr[kk + 1]; } else bad = true;
      }
      f += end[h] - pos0[h];
    }
  }
  if (bad) atomicOr(&cnt->error, 1u);
  // 1. symbolic merge
#pragma unroll
  for (int h = 0; h < K; ++h) {
    pos[h] = pos0[h];
    if (pos[h] < end[h]) col[h] = b_col[pos[h]];
  }
  u32 z = 0;
  for (;;) {
    u32 cmin = col[0];
#pragma unroll
    for (int h = 1; h < K; ++h) cmin = min(cmin, col[h]);
    if (cmin == INF_COL) break;
    ++z;
#pragma unroll
    for (int h = 0; h < K; ++h) {
      if (col[h] == cmin) {
        ++pos[h];
        col[h] = (pos[h] < end[h]) ? b_col[pos[h]] : INF_COL;
      }
    }
  }
  // 2. block scan of z, look-back for the block's first position in C
  u32 incl = z, fsum = f;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u32 y = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += y;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) fsum += __shfl_xor_sync(0xffffffffu, fsum, d);
  if (lane == 31) s_warp[wid] = incl;
  if (lane == 0 && fsum) atomicAdd(&cnt->total_flops, (ull)fsum);
  __syncthreads();
  if (wid == 0) {
    const u32 w = lane < BLOCK / 32 ? s_warp[lane] : 0;
    u32 wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const u32 y = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += y;
    }
    if (lane < BLOCK / 32) s_warp[lane] = wi - w;  // exclusive warp offsets
    const u64 block_total = __shfl_sync(0xffffffffu, wi, 31);
    if (lane == 0) state[tile] = (tile == 0 ? F_PFX : F_AGG) | block_total;
    u64 excl = 0;
    if (tile > 0) {
      long long p = (long long)tile - 1;
      for (;;) {
        const long long idx = p - lane;
        u64 s;
        if (idx >= 0) {
          do { s = state[idx]; } while ((s >> 62) == 0);
        } else {
          s = F_PFX;
        }
        const unsigned pm = __ballot_sync(0xffffffffu, (s >> 62) == 2);
        const int firstp = pm ? (__ffs(pm) - 1) : 32;
        u64 c = (lane <= firstp) ? (s & F_MASK) : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        excl += c;
        if (pm) break;
        p -= 32;
      }
      if (lane == 0) state[tile] = F_PFX | (excl + block_total);
    }
    if (lane == 0) {
      s_base = excl;
      if ((u64)(tile + 1) * BLOCK >= (u64)m) {  // last block: row_ptr[m] = nnz(C)
        c_ptr[m] = excl + block_total;
        cnt->total_nnz = excl + block_total;
      }
    }
  }
  __syncthreads();
  const u64 c0 = s_base + s_warp[wid] + (incl - z);
  if (row < m) c_ptr[row] = c0;
  // 3. numeric merge from the same run heads
#pragma unroll
  for (int h = 0; h < K; ++h) {
    pos[h] = pos0[h];
    col[h] = (pos[h] < end[h]) ? b_col[pos[h]] : INF_COL;
  }
  merge_rows_out<V, K, BLOCK, false>(pos, end, col, av, z, c0, b_col, b_val, sv, sk, c_col, c_val);
}

}  // namespace
