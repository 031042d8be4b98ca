// spgemm.cu — Gustavson CSR x CSR SpGEMM for sm_100a, the replacement for the body of
// CsrMatrix::mul_hash (spam_csr/src/mul_hash.rs:13-36):
//
//   rows_to_threads   (mul_hash.rs:38-64)   -> k_flop_count   : per-row intermediate-product count
//                                               + row binning (the GPU analogue of the thread blocks)
//   mul_hash_symbolic (mul_hash.rs:66-103)  -> k_sym_*        : distinct columns per row (linear-probe set)
//   checked_inclusive_scan (lib.rs:267-274) -> scan.cu        : row_ptr of C
//   mul_hash_numeric  (mul_hash.rs:105-201) -> k_num_*        : linear-probe map accumulate, then the
//                                               B2=true branch (:164-175): per-row sort by column
//
// Hash design = linprobe (linprobe/src/{lib,set,map}.rs): open addressing, linear probing,
// table = max(16, 2*npow2(n)) slots, u32::MAX = empty; hash = key*107 (exact) in the thread-per-row
// kernels, high bits of a Fibonacci hash elsewhere (rowhash.cuh explains why).
//
// Bins (common.cuh): k-way merge of sorted runs for short rows when B is sorted (merge.cuh),
// thread-per-row private smem tables for tiny rows of unsorted inputs (this file), one warp or a
// team of warps per row with a shared-memory table for medium rows (rowhash.cuh), and persistent
// blocks with global-memory tables for heavy power-law rows (this file).  Everything is integer/byte
// gather-scatter work bounded by HBM/L2 and shared-memory throughput; no tensor cores.
#include "common.cuh"
#include "merge.cuh"
#include "rowhash.cuh"
#include "esc.cuh"
#include "msort.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// block helpers
// ------------------------------------------------------------------------------------------
template <int T>
__device__ __forceinline__ u32 block_reduce_sum_u32(u32 v, u32* s_warp /* >= T/32 */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if (T == 32) return v;
  if (lane == 0) s_warp[wid] = v;
  __syncthreads();
  u32 r = 0;
#pragma unroll
  for (int w = 0; w < T / 32; ++w) r += s_warp[w];
  __syncthreads();
  return r;
}

// exclusive scan of one u32 per thread across the block; returns exclusive prefix, *total = sum
template <int T>
__device__ __forceinline__ u32 block_excl_scan_u32(u32 v, u32* s_warp /* >= T/32 */, u32* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  u32 x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    u32 y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  if (T == 32) {
    *total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
  }
  if (lane == 31) s_warp[wid] = x;
  __syncthreads();
  u32 woff = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < T / 32; ++w) {
    u32 s = s_warp[w];
    if (w < wid) woff += s;
    tot += s;
  }
  __syncthreads();
  *total = tot;
  return woff + x - v;
}

// ------------------------------------------------------------------------------------------
// flop count: flop_i = sum_{k in A.row i} nnz(B.row k)          (mul_hash.rs:39-50)
// thread per row; rows longer than 32 entries are swept cooperatively by the warp.
// ------------------------------------------------------------------------------------------
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_flop_count(u64 m, u64 b_rows, const u64* __restrict__ a_ptr,
                                                      const u32* __restrict__ a_col, const u64* __restrict__ b_ptr,
                                                      u32* __restrict__ flop_out, Counters* cnt, int do_bins,
                                                      int mode) {
  const bool merge_ok = (mode & MODE_MERGE) != 0;
  __shared__ u32 s_hist[NBINS];
  __shared__ ull s_total;
  __shared__ u32 s_max, s_maxalen;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < NBINS) s_hist[tid] = 0;
  if (tid == 0) { s_total = 0; s_max = 0; s_maxalen = 0; }
  __syncthreads();
  const u64 row = (u64)blockIdx.x * BLOCK + tid;
  const bool valid = row < m;
  u64 lo = 0, hi = 0;
  if (valid) { lo = a_ptr[row]; hi = a_ptr[row + 1]; }
  const u64 len = hi - lo;
  u64 f = 0;
  bool bad = false;
  if (valid && len <= 32) {
    for (u64 e = lo; e < hi; ++e) {
      const u32 k = a_col[e];
      if (k < b_rows) f += b_ptr[k + 1] - b_ptr[k]; else bad = true;
    }
  }
  unsigned longmask = __ballot_sync(0xffffffffu, valid && len > 32);
  while (longmask) {
    const int src = __ffs(longmask) - 1;
    longmask &= longmask - 1;
    const u64 l = __shfl_sync(0xffffffffu, lo, src), hh = __shfl_sync(0xffffffffu, hi, src);
    u64 part = 0;
    for (u64 e = l + lane; e < hh; e += 32) {
      const u32 k = a_col[e];
      if (k < b_rows) part += b_ptr[k + 1] - b_ptr[k]; else bad = true;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (lane == src) f = part;
  }
  if (bad) atomicOr(&cnt->error, 1u);
  if (valid) {
    const u32 fs = f > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)f;  // saturate: still lands in the heavy bin
    flop_out[row] = fs;
    if (do_bins) {
      const u32 alen = len > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)len;
      const int bin = sym_bin_of(fs, alen, merge_ok);
      atomicAdd(&s_hist[bin], 1u);
      if (merge_ok && alen <= MERGE_K) atomicMax(&s_maxalen, alen);  // bound for both merge bins
      if (bin != MERGE_BIN) atomicMax(&s_max, fs);
    }
  }
  // block total of f
  u64 t = f;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
  if (lane == 0 && t) atomicAdd(&s_total, (ull)t);
  __syncthreads();
  if (tid < NBINS && s_hist[tid]) atomicAdd(&cnt->sym_bins[tid], s_hist[tid]);
  if (tid == 0) {
    if (s_total) atomicAdd(&cnt->total_flops, s_total);
    if (s_max) atomicMax(&cnt->max_flop, s_max);
    if (s_maxalen) atomicMax(&cnt->max_alen, s_maxalen);
  }
}

// histogram of the numeric bins, from (row nnz, flop)
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_num_bin_count(u64 m, const u64* __restrict__ a_ptr,
                                                         const u32* __restrict__ row_nnz,
                                                         const u32* __restrict__ flop, Counters* cnt, int mode) {
  __shared__ u32 s_hist[NBINS];
  __shared__ u32 s_max;
  const int tid = threadIdx.x;
  if (tid < NBINS) s_hist[tid] = 0;
  if (tid == 0) s_max = 0;
  __syncthreads();
  const u64 row = (u64)blockIdx.x * BLOCK + tid;
  if (row < m) {
    const u32 z = row_nnz[row], f = flop[row];
    const u64 len = a_ptr[row + 1] - a_ptr[row];
    const u32 alen = len > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)len;
    // rows of the symbolic merge bin were already histogrammed by k_sym_merge
    if (sym_bin_of(f, alen, (mode & MODE_MERGE) != 0) != MERGE_BIN) {
      const int bin = num_bin_of(z, f, alen, mode);
      atomicAdd(&s_hist[bin], 1u);
      // longest row that can reach the global-table kernel (its own bin, or handed back by a bucket-sort bin)
      if (bin == HEAVY_BIN || bin >= ESC_BIN0) atomicMax(&s_max, z);
    }
  }
  __syncthreads();
  if (tid < NBINS && s_hist[tid]) atomicAdd(&cnt->num_bins[tid], s_hist[tid]);
  if (tid == 0 && s_max) atomicMax(&cnt->max_nnz, s_max);
}

// scatter row ids into per-bin segments of perm[] (counting sort by bin; a block's rows stay
// together inside each bin so neighbouring rows still share cache lines of A and B)
template <int BLOCK, bool NUMERIC>
__global__ void __launch_bounds__(BLOCK) k_bin_scatter(u64 m, const u64* __restrict__ a_ptr,
                                                       const u32* __restrict__ row_nnz,
                                                       const u32* __restrict__ flop, BinBase base, u32* cursors,
                                                       u32* __restrict__ perm, int mode) {
  __shared__ u32 s_cnt[NBINS];
  __shared__ u32 s_base[NBINS];
  const int tid = threadIdx.x;
  if (tid < NBINS) s_cnt[tid] = 0;
  __syncthreads();
  const u64 row = (u64)blockIdx.x * BLOCK + tid;
  int b = -1;
  u32 r = 0;
  if (row < m) {
    const u64 len = a_ptr[row + 1] - a_ptr[row];
    const u32 alen = len > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)len;
    b = NUMERIC ? num_bin_of(row_nnz[row], flop[row], alen, mode) : sym_bin_of(flop[row], alen, (mode & MODE_MERGE) != 0);
    r = atomicAdd(&s_cnt[b], 1u);
  }
  __syncthreads();
  if (tid < NBINS && s_cnt[tid]) s_base[tid] = base.v[tid] + atomicAdd(&cursors[tid], s_cnt[tid]);
  __syncthreads();
  if (b >= 0) perm[s_base[b] + r] = (u32)row;
}

// ------------------------------------------------------------------------------------------
// SYMBOLIC, tiny rows (flop <= 32): one thread per row, private 64-slot key table in shared
// memory laid out [slot][thread] so a warp's accesses never bank-conflict whatever the slots.
// Sequential insertion exactly like linprobe::HashSet::insert (set.rs:109-160), no atomics.
// ------------------------------------------------------------------------------------------
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_sym_tiny(u32 n, const u32* __restrict__ perm,
                                                    const u64* __restrict__ a_ptr, const u32* __restrict__ a_col,
                                                    const u64* __restrict__ b_ptr, const u32* __restrict__ b_col,
                                                    const u32* __restrict__ flop, u32* __restrict__ row_nnz) {
  extern __shared__ u32 sm_tab[];  // [2*SYM_TINY_MAX][BLOCK]
  const int tid = threadIdx.x;
  const u32 i = blockIdx.x * BLOCK + tid;
  if (i >= n) return;
  const u32 row = perm ? perm[i] : i;
  const u32 f = flop[row];
  if (f == 0) { row_nnz[row] = 0; return; }        // mul_hash.rs:84-86
  const u32 cap = table_size_u32(f), mask = cap - 1;  // <= 64
  u32* tab = sm_tab + tid;
  for (u32 s = 0; s < cap; ++s) tab[s * BLOCK] = EMPTY_KEY;
  u32 cnt = 0;
  const u64 lo = a_ptr[row], hi = a_ptr[row + 1];
  for (u64 e = lo; e < hi; ++e) {
    const u32 k = a_col[e];
    const u64 bl = b_ptr[k], bh = b_ptr[k + 1];
    for (u64 j = bl; j < bh; ++j) {
      const u32 key = b_col[j];
      u32 s = slot_of(key, mask);
      for (;;) {
        const u32 cur = tab[s * BLOCK];
        if (cur == key) break;
        if (cur == EMPTY_KEY) { tab[s * BLOCK] = key; ++cnt; break; }
        s = (s + 1) & mask;
      }
    }
  }
  row_nnz[row] = cnt;  // mul_hash.rs:95
}

// ------------------------------------------------------------------------------------------
// SYMBOLIC, heavy rows: persistent blocks, one global-memory key table per block, rows handed
// out through an atomic work counter.
// ------------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(T) k_sym_heavy(u32 n, const u32* __restrict__ perm, const u64* __restrict__ a_ptr,
                                                 const u32* __restrict__ a_col, const u64* __restrict__ b_ptr,
                                                 const u32* __restrict__ b_col, const u32* __restrict__ flop,
                                                 u32* __restrict__ row_nnz, u32* tables, u64 table_stride,
                                                 u32 b_cols, u32* work, int wshift) {
  __shared__ u32 s_item;
  __shared__ u32 s_warp[32];
  const int tid = threadIdx.x;
  u32* keys = tables + (u64)blockIdx.x * table_stride;
  const int W = 1 << wshift, sub = tid >> wshift, lane = tid & (W - 1), nsub = T >> wshift;
  for (;;) {
    if (tid == 0) s_item = atomicAdd(work, 1u);
    __syncthreads();
    const u32 item = s_item;
    __syncthreads();
    if (item >= n) break;
    const u32 row = perm ? perm[item] : item;
    u32 f = flop[row];
    if (f > b_cols) f = b_cols;  // at most cols(B) distinct keys
    const u64 cap = 2ull * npow2_u64(f), mask = cap - 1; const int hshift = 64 - (63 - __clzll((long long)cap));
    for (u64 s = tid; s < cap; s += T) __stcg(&keys[s], EMPTY_KEY);
    __syncthreads();
    u32 cnt = 0;
    const u64 lo = a_ptr[row], hi = a_ptr[row + 1];
    for (u64 e = lo + sub; e < hi; e += nsub) {
      const u32 k = a_col[e];
      const u64 bl = b_ptr[k], bh = b_ptr[k + 1];
      for (u64 j = bl + lane; j < bh; j += W) {
        const u32 key = b_col[j];
        u64 s = ((u64)key * 11400714819323198485ull) >> hshift;
        for (;;) {
          const u32 cur = __ldcg(&keys[s]);
          if (cur == key) break;
          if (cur == EMPTY_KEY) {
            const u32 old = atomicCAS(&keys[s], EMPTY_KEY, key);
            if (old == EMPTY_KEY) { ++cnt; break; }
            if (old == key) break;
          }
          s = (s + 1) & mask;
        }
      }
    }
    const u32 total = block_reduce_sum_u32<T>(cnt, s_warp);
    if (tid == 0) row_nnz[row] = total;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// NUMERIC, tiny rows (nnz <= 16, flop <= 128): one thread per row, private key/value table
// [slot][thread] in shared memory, table size max(16, 2*npow2(nnz)) exactly as
// HashMap::shrink_to (map.rs:49-58).  Products are visited in the reference's order
// (mul_hash.rs:145-162) and accumulated sequentially with separate mul and add, the first product
// stored — the sums are bit-identical to the reference's.  Rows are then ranked by column and
// staged through shared memory so the block writes C with coalesced stores.
// ------------------------------------------------------------------------------------------
template <class V, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_num_tiny(u32 n, const u32* __restrict__ perm,
                                                    const u64* __restrict__ a_ptr, const u32* __restrict__ a_col,
                                                    const V* __restrict__ a_val, const u64* __restrict__ b_ptr,
                                                    const u32* __restrict__ b_col, const V* __restrict__ b_val,
                                                    const u64* __restrict__ c_ptr, u32* __restrict__ c_col,
                                                    V* __restrict__ c_val) {
  constexpr int SLOTS = 2 * NUM_TINY_MAX;  // 32
  extern __shared__ __align__(16) unsigned char sm_raw[];
  V* tvals = reinterpret_cast<V*>(sm_raw);                       // [SLOTS][BLOCK]
  V* svals = tvals + SLOTS * BLOCK;                              // [NUM_TINY_MAX*BLOCK] staging
  u64* s_cptr = reinterpret_cast<u64*>(svals + NUM_TINY_MAX * BLOCK);  // [BLOCK]
  u32* tkeys = reinterpret_cast<u32*>(s_cptr + BLOCK);           // [SLOTS][BLOCK]
  u32* skeys = tkeys + SLOTS * BLOCK;                            // [NUM_TINY_MAX*BLOCK]
  u32* s_off = skeys + NUM_TINY_MAX * BLOCK;                     // [BLOCK+1]
  u32* s_warp = s_off + BLOCK + 1;                               // [32]

  const int tid = threadIdx.x;
  const u32 i = blockIdx.x * BLOCK + tid;
  const bool valid = i < n;
  const u32 row = valid ? (perm ? perm[i] : i) : 0;
  u64 c0 = 0;
  u32 z = 0;
  if (valid) { c0 = c_ptr[row]; z = (u32)(c_ptr[row + 1] - c0); }
  u32 total;
  const u32 off = block_excl_scan_u32<BLOCK>(z, s_warp, &total);
  s_off[tid] = off;
  s_cptr[tid] = c0;
  if (tid == 0) s_off[BLOCK] = total;

  if (z > 0) {
    const u32 cap = table_size_u32(z), mask = cap - 1;
    u32* keys = tkeys + tid;
    V* vals = tvals + tid;
    for (u32 s = 0; s < cap; ++s) keys[s * BLOCK] = EMPTY_KEY;
    const u64 lo = a_ptr[row], hi = a_ptr[row + 1];
    for (u64 e = lo; e < hi; ++e) {
      const u32 k = a_col[e];
      const V av = a_val[e];
      const u64 bl = b_ptr[k], bh = b_ptr[k + 1];
      for (u64 j = bl; j < bh; ++j) {
        const u32 key = b_col[j];
        const V prod = Num<V>::mul(av, b_val[j]);
        u32 s = slot_of(key, mask);
        for (;;) {
          const u32 cur = keys[s * BLOCK];
          if (cur == key) { vals[s * BLOCK] = Num<V>::add(vals[s * BLOCK], prod); break; }
          if (cur == EMPTY_KEY) { keys[s * BLOCK] = key; vals[s * BLOCK] = prod; break; }
          s = (s + 1) & mask;
        }
      }
    }
    // drain in slot order (map.rs:59-63), compacting in place (write index never passes read index)
    u32 cntz = 0;
    for (u32 s = 0; s < cap; ++s) {
      const u32 kk = keys[s * BLOCK];
      if (kk != EMPTY_KEY) {
        const V vv = vals[s * BLOCK];
        keys[cntz * BLOCK] = kk;
        vals[cntz * BLOCK] = vv;
        ++cntz;
      }
    }
    // rank by column (keys are distinct) == sort_unstable_by_key (mul_hash.rs:166)
    for (u32 x = 0; x < cntz; ++x) {
      const u32 kx = keys[x * BLOCK];
      u32 r = 0;
      for (u32 y = 0; y < cntz; ++y) r += (keys[y * BLOCK] < kx) ? 1u : 0u;
      skeys[off + r] = kx;
      svals[off + r] = vals[x * BLOCK];
    }
  }
  __syncthreads();
  // coalesced copy-out: staged element q belongs to the last local row r with s_off[r] <= q
  for (u32 q = tid; q < total; q += BLOCK) {
    int lo_r = 0, hi_r = BLOCK;  // invariant: s_off[lo_r] <= q < s_off[hi_r]
    while (hi_r - lo_r > 1) {
      const int mid = (lo_r + hi_r) >> 1;
      if (s_off[mid] <= q) lo_r = mid; else hi_r = mid;
    }
    const u64 dst = s_cptr[lo_r] + (q - s_off[lo_r]);
    c_col[dst] = skeys[q];
    c_val[dst] = svals[q];
  }
}

template <class V, int BLOCK>
constexpr size_t num_tiny_smem() {
  return (size_t)(2 * NUM_TINY_MAX) * BLOCK * (sizeof(V) + 4) + (size_t)NUM_TINY_MAX * BLOCK * (sizeof(V) + 4) +
         (size_t)BLOCK * 8 + (size_t)(BLOCK + 1 + 32) * 4;
}

// ------------------------------------------------------------------------------------------
// bitonic sort of n2 (power of two) key/value pairs held in shared or global memory, by key
// ascending; EMPTY_KEY pads sort to the end.  One thread block.
// ------------------------------------------------------------------------------------------
template <int T, class V, class KP, class VP>
__device__ __forceinline__ void block_bitonic_sort(KP keys, VP vals, u32 n2) {
  const int tid = threadIdx.x;
  for (u32 k = 2; k <= n2; k <<= 1) {
    for (u32 j = k >> 1; j > 0; j >>= 1) {
      for (u32 p = tid; p < (n2 >> 1); p += T) {
        const u32 i = 2 * p - (p & (j - 1));
        const u32 l = i + j;
        const bool up = (i & k) == 0;
        const u32 ki = keys[i], kl = keys[l];
        if ((ki > kl) == up && ki != kl) {
          keys[i] = kl; keys[l] = ki;
          const V vi = vals[i], vl = vals[l];
          vals[i] = vl; vals[l] = vi;
        }
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------
// NUMERIC, heavy rows: persistent blocks, global-memory key+value tables (ld.cg/st.cg: the
// tables are updated by L2 atomics, so never read them through L1).
// ------------------------------------------------------------------------------------------
template <class V, int T>
__global__ void __launch_bounds__(T) k_num_heavy(u32 n, const u32* __restrict__ perm, const u64* __restrict__ a_ptr,
                                                 const u32* __restrict__ a_col, const V* __restrict__ a_val,
                                                 const u64* __restrict__ b_ptr, const u32* __restrict__ b_col,
                                                 const V* __restrict__ b_val, const u64* __restrict__ c_ptr,
                                                 u32* __restrict__ c_col, V* __restrict__ c_val, u32* key_tables,
                                                 V* val_tables, u64 table_stride, u32* cnt_tables, u32* ord_tables,
                                                 u32 b_cols, u32* work, Counters* cnt_dev, const u32* n_dev) {
  // n_dev != nullptr: the row list was filled on the device (rows handed back by the bucket-sort bins)
  if (n_dev) n = *n_dev;
  constexpr int ITEMS = 4;
  constexpr u32 HEAVY_RANK_MAX = 2048;  // longest bucket ranked by counting; beyond: bitonic fallback
  __shared__ u32 s_item, s_maxcnt;
  __shared__ u32 s_warp[32];
  const int tid = threadIdx.x, wid = tid >> 5, ln = tid & 31;
  u32* keys = key_tables + (u64)blockIdx.x * table_stride;
  V* vals = val_tables + (u64)blockIdx.x * table_stride;
  u32* cnt = cnt_tables + (u64)blockIdx.x * (table_stride / 2 + 4);  // [npow2(z) + 1] bucket counters, 16 B aligned
  u32* okey = ord_tables + (u64)blockIdx.x * table_stride;           // [z] keys in bucket order
  u32* ord = okey + table_stride / 2;                                // [z] their table slots
  for (;;) {
    if (tid == 0) s_item = atomicAdd(work, 1u);
    __syncthreads();
    const u32 item = s_item;
    __syncthreads();
    if (item >= n) break;
    const u32 row = perm ? perm[item] : item;
    const u64 c0 = c_ptr[row];
    const u32 z = (u32)(c_ptr[row + 1] - c0);
    if (z == 0) continue;
    const u64 cap = 2ull * npow2_u64(z), mask = cap - 1; const int hshift = 64 - (63 - __clzll((long long)cap));
    const u32 NB = max(npow2_u32(z), 4u * T);  // every thread scans a multiple of four counters below
    for (u64 s = tid; s < cap; s += T) { __stcg(&keys[s], EMPTY_KEY); __stcg(&vals[s], Num<V>::zero()); }
    for (u32 b = tid; b <= NB; b += T) __stcg(&cnt[b], 0u);
    if (tid == 0) s_maxcnt = 0;
    __syncthreads();
    // FLAT enumeration (rowhash.cuh): the products of 32 A entries are spread evenly over the block's lanes,
    // whatever the B row lengths (a sub-warp per A entry left a quarter of the kernel waiting at the
    // barrier below for the warps that drew the hub rows: profiles/r01_rmat20_v3_team_drain.txt)
    const u64 lo = a_ptr[row], hi = a_ptr[row + 1];
    for (u64 ec = lo; ec < hi; ec += 32) {
      const AChunk<V> c = load_chunk<V, true, true>(ec, hi, ln, a_col, a_val, b_ptr);
      for (u32 p0 = 32u * wid; p0 < c.total; p0 += T) {
        u64 addr;
        V av;
        locate<V, true>(c, p0 + ln, addr, av);
        if (p0 + ln < c.total) {
          const u32 key = b_col[addr];
          const V prod = Num<V>::mul(av, b_val[addr]);
          u64 s = ((u64)key * 11400714819323198485ull) >> hshift;
          for (;;) {
            const u32 cur = __ldcg(&keys[s]);
            if (cur != key) {
              if (cur != EMPTY_KEY) { s = (s + 1) & mask; continue; }
              const u32 old = atomicCAS(&keys[s], EMPTY_KEY, key);
              if (old != EMPTY_KEY && old != key) { s = (s + 1) & mask; continue; }
            }
            Num<V>::atomic_add(&vals[s], prod);
            break;
          }
        }
      }
    }
    __syncthreads();
    // drain, sorted by column: count the occupied slots into NB = npow2(z) order-preserving buckets over the
    // column range, scan, scatter (key, slot) bucket by bucket into scratch, then every entry ranks itself
    // inside its bucket and stores (key, value) at its final place in C: one coalesced pass over C.
    // Falls back to compaction + bitonic network when some bucket is far longer than average.
    {
      const int lgnb = 31 - __clz(NB);
      const int rbits = b_cols > 1 ? 32 - __clz(b_cols - 1) : 0;
      const int bshift = rbits > lgnb ? rbits - lgnb : 0;
      // every loop below keeps U independent L2 requests in flight per thread: with one request at a time a
      // heavy row costs ~0.35 ms of pure L2 latency, a fixed price every shard of a partitioned product pays
      constexpr int U = 4;
      for (u64 s0 = tid; s0 < cap; s0 += (u64)U * T) {
        u32 kk[U];
#pragma unroll
        for (int i = 0; i < U; ++i) { const u64 s = s0 + (u64)i * T; kk[i] = s < cap ? __ldcg(&keys[s]) : EMPTY_KEY; }
#pragma unroll
        for (int i = 0; i < U; ++i) if (kk[i] != EMPTY_KEY) atomicAdd(&cnt[kk[i] >> bshift], 1u);
      }
      __syncthreads();
      // NB >= 4096: each thread scans a multiple of four consecutive counters
      const u32 chunk = NB / T;
      const u32 b0 = tid * chunk, b1 = b0 + chunk;
      u32 sum = 0, mx = 0;
      for (u32 b = b0; b < b1; b += 4) {
        const uint4 c = __ldcg(reinterpret_cast<const uint4*>(&cnt[b]));
        sum += c.x + c.y + c.z + c.w;
        mx = max(max(mx, max(c.x, c.y)), max(c.z, c.w));
      }
      u32 total;
      u32 runb = block_excl_scan_u32<T>(sum, s_warp, &total);
      for (u32 b = b0; b < b1; b += 4) {
        const uint4 c = __ldcg(reinterpret_cast<const uint4*>(&cnt[b]));
        uint4 o;
        o.x = runb; o.y = o.x + c.x; o.z = o.y + c.y; o.w = o.z + c.z;
        runb = o.w + c.w;
        __stcg(reinterpret_cast<uint4*>(&cnt[b]), o);
      }
      if (mx) atomicMax(&s_maxcnt, mx);
      __syncthreads();
      if (s_maxcnt <= HEAVY_RANK_MAX) {
        for (u64 s0 = tid; s0 < cap; s0 += (u64)U * T) {
          u32 kk[U], pos[U];
#pragma unroll
          for (int i = 0; i < U; ++i) { const u64 s = s0 + (u64)i * T; kk[i] = s < cap ? __ldcg(&keys[s]) : EMPTY_KEY; }
#pragma unroll
          for (int i = 0; i < U; ++i)
            if (kk[i] != EMPTY_KEY) pos[i] = atomicAdd(&cnt[kk[i] >> bshift], 1u);  // afterwards cnt[b] = end of bucket b
#pragma unroll
          for (int i = 0; i < U; ++i)
            if (kk[i] != EMPTY_KEY) { __stcg(&okey[pos[i]], kk[i]); __stcg(&ord[pos[i]], (u32)(s0 + (u64)i * T)); }
        }
        __syncthreads();
        for (u32 p0 = tid; p0 < z; p0 += U * T) {
          u32 kk[U], lo_b[U], hi_b[U], slot[U];
          V val[U];
#pragma unroll
          for (int i = 0; i < U; ++i) {
            const u32 p = p0 + i * T;
            kk[i] = 0; slot[i] = 0;
            if (p < z) { kk[i] = __ldcg(&okey[p]); slot[i] = __ldcg(&ord[p]); }
          }
#pragma unroll
          for (int i = 0; i < U; ++i) {
            const u32 bkt = kk[i] >> bshift;
            lo_b[i] = 0; hi_b[i] = 0; val[i] = Num<V>::zero();
            if (p0 + i * T < z) {
              lo_b[i] = bkt ? __ldcg(&cnt[bkt - 1]) : 0u;
              hi_b[i] = __ldcg(&cnt[bkt]);
              val[i] = __ldcg(&vals[slot[i]]);
            }
          }
#pragma unroll
          for (int i = 0; i < U; ++i) {
            if (p0 + i * T < z) {
              u32 rank = 0;
              if (hi_b[i] - lo_b[i] > 1) {
#pragma unroll 4
                for (u32 j = lo_b[i]; j < hi_b[i]; ++j) rank += (__ldcg(&okey[j]) < kk[i]) ? 1u : 0u;
              }
              c_col[c0 + lo_b[i] + rank] = kk[i];
              c_val[c0 + lo_b[i] + rank] = val[i];
            }
          }
        }
        __syncthreads();
        continue;
      }
    }
    // fallback: in-place chunked compaction: chunk c is read into registers, barrier, then written at
    // the running output offset, which never passes the start of the next unread chunk.
    if (tid == 0) atomicAdd(&cnt_dev->fb_heavy_bitonic, 1u);
    u32 run = 0;
    for (u64 cbase = 0; cbase < cap; cbase += (u64)T * ITEMS) {
      u32 rk[ITEMS];
      V rv[ITEMS];
      u32 mine = 0;
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        const u64 s = cbase + (u64)it * T + tid;
        rk[it] = EMPTY_KEY;
        rv[it] = Num<V>::zero();
        if (s < cap) { rk[it] = __ldcg(&keys[s]); rv[it] = __ldcg(&vals[s]); }
        mine += (rk[it] != EMPTY_KEY) ? 1u : 0u;
      }
      u32 total;
      u32 pos = run + block_excl_scan_u32<T>(mine, s_warp, &total);  // barriers inside end the reads
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        if (rk[it] != EMPTY_KEY) { __stcg(&keys[pos], rk[it]); __stcg(&vals[pos], rv[it]); ++pos; }
      }
      run += total;
      __syncthreads();
    }
    const u32 n2 = npow2_u32(z);
    for (u32 s = z + tid; s < n2; s += T) __stcg(&keys[s], EMPTY_KEY);
    __syncthreads();
    // bitonic sort in global memory (volatile accesses: L2-coherent within the block)
    block_bitonic_sort<T, V>((volatile u32*)keys, (volatile V*)vals, n2);
    for (u32 s = tid; s < z; s += T) { c_col[c0 + s] = __ldcg(&keys[s]); c_val[c0 + s] = __ldcg(&vals[s]); }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------
// One B row per warp batch (DIRECT) pays off when B rows are about a warp long and not skewed (stencils);
// otherwise the products are flattened over the lanes (FLAT).  Uses the per-matrix cached stats.
bool direct_enumeration(const spam_dcsr* b) {
  // DIRECT updates values with a plain read-modify-write per batch, which needs distinct columns inside a B
  // row: known only when the rows are strictly increasing (a duplicate column in an unsorted row is folded
  // correctly by the FLAT path's match-any)
  if (b->rows_sorted != 1) return false;
  const double mean = b->rows ? (double)b->nnz / (double)b->rows : 0.0;
  return mean >= 12.0 && (double)b->max_row_len <= 3.0 * mean + 8.0;
}

int wshift_for(const spam_dcsr* b, int tmax_shift) {
  const double avg = b->rows ? (double)b->nnz / (double)b->rows : 1.0;
  int ws = 2;  // at least 4 lanes per A entry
  while ((1 << ws) < avg && ws < 5) ++ws;
  if (ws > tmax_shift) ws = tmax_shift;
  return ws;
}

template <class K>
int set_smem(spam_handle* h, K kernel, size_t bytes) {
  if (bytes > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return SPAM_OK;
}

static inline u32 n_esc_rows(const u32* count) {
  u32 n = 0;
  for (int bin = ESC_BIN0; bin <= ESC_HEAVY_BIN; ++bin) n += count[bin];
  return n;
}

// SPAM_L2_PERSIST (experiment): an access-policy window on the stream marks one array of B as persisting in L2 while a
// merge kernel streams A and C through it.  Returns true when a window was set (clear it after the launch).
static inline bool l2_window_set(spam_handle* h, const void* base, size_t bytes) {
  if (!h->l2_persist || !h->l2_persist_max || !bytes) return false;
  cudaStreamAttrValue v = {};
  v.accessPolicyWindow.base_ptr = const_cast<void*>(base);
  v.accessPolicyWindow.num_bytes = bytes < h->l2_window_max ? bytes : h->l2_window_max;
  const double r = (double)h->l2_persist_max / (double)v.accessPolicyWindow.num_bytes;
  v.accessPolicyWindow.hitRatio = (float)(r > 1.0 ? 1.0 : r);
  v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  v.accessPolicyWindow.missProp = h->l2_persist >= 10 ? cudaAccessPropertyNormal : cudaAccessPropertyStreaming;
  if (cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &v) == cudaSuccess) return true;
  cudaGetLastError();
  return false;
}
static inline void l2_window_clear(spam_handle* h) {
  cudaStreamAttrValue v = {};
  v.accessPolicyWindow.num_bytes = 0;
  cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &v);
}

struct Bins {
  u32 count[NBINS];
  u32 base[NBINS];
  u32* perm;  // device, null when one bin holds every row (identity)
};

int build_perm(spam_handle* h, u64 m, const u32* counts, bool numeric, const u64* d_aptr, const u32* d_row_nnz,
               const u32* d_flop, int mode, Bins* out) {
  u32 acc = 0;
  bool identity = false;
  for (int b = 0; b < NBINS; ++b) {
    out->count[b] = counts[b];
    out->base[b] = acc;
    acc += counts[b];
    if (counts[b] == m) identity = true;
  }
  out->perm = nullptr;
  if (identity || m == 0) return SPAM_OK;
  CKS(dev_alloc_t(h, &out->perm, m));
  BinBase bb;
  for (int b = 0; b < NBINS; ++b) bb.v[b] = out->base[b];
  u32* cursors = numeric ? h->d_cnt->num_cursor : h->d_cnt->sym_cursor;
  const unsigned grid = (unsigned)((m + 255) / 256);
  if (numeric)
    k_bin_scatter<256, true><<<grid, 256, 0, h->stream>>>(m, d_aptr, d_row_nnz, d_flop, bb, cursors, out->perm, mode);
  else
    k_bin_scatter<256, false><<<grid, 256, 0, h->stream>>>(m, d_aptr, d_row_nnz, d_flop, bb, cursors, out->perm, mode);
  count_launch(h);
  CK(cudaGetLastError());
  return SPAM_OK;
}

}  // namespace

struct SpgemmPending {
  const spam_dcsr* a;
  const spam_dcsr* b;        // the right-hand side the kernels read (a cached sorted copy when B's rows are unsorted)
  const spam_dcsr* b_orig;   // the caller's right-hand side: its row order defines the B2 = false output order
  u32* d_flop;
  u32* d_row_nnz;
  u64* d_cptr;
  u64 nnz;
  u32 num_counts[NBINS];
  u32 max_nnz;
  u32 max_alen;   // longest A row in the merge bins
  int merge_ok;   // B's rows are sorted and nnz(B) < 2^32: the merge bin is in use
  int mode;       // MODE_MERGE | MODE_ESC: the optional bins this product uses
};

namespace {

// Which merge kernels: when the entries of A's short rows are far apart, the B rows a block walks do not stay in
// L1/L2 and every step of a run head is a real memory round trip — the variants that keep the next column (and the
// head's value) in registers win (C5 A*A^T: 0.97 -> 0.85 ms); for banded matrices the run heads hit L1 and the extra
// registers only cost residency (Poisson 2048^2: 0.567 -> 0.590 ms).  Footprint estimate: mean column spread of an
// A row x mean B row in bytes, against 8 MB.
int merge_prefetch_mode(const spam_handle* h, const spam_dcsr* a, const spam_dcsr* b) {
  if (h->merge_pf >= 0) return h->merge_pf;
  if (!a->rows || !b->rows) return 0;
  const double foot = (double)a->spread_sum / (double)a->rows * ((double)b->nnz / (double)b->rows) * 12.0;
  return foot > 8e6 ? 6 : 0;
}

// Cached per matrix: are all rows strictly increasing by column?  One pass over col_idx the first
// time a matrix is used as a right-hand side (device matrices are immutable through this API).
int ensure_rows_sorted(spam_handle* h, const spam_dcsr* b) {
  spam_dcsr* mb = const_cast<spam_dcsr*>(b);
  if (b->rows_sorted < 0) {
    if (b->rows == 0) { mb->rows_sorted = 1; mb->max_row_len = 0; mb->invalid = b->nnz ? 1 : 0; }
    else {
      CK(cudaMemsetAsync(&h->d_cnt->unsorted, 0, 3 * sizeof(u32), h->stream));  // unsorted, max_rowlen, invalid
      CK(cudaMemsetAsync(&h->d_cnt->spread_sum, 0, sizeof(ull), h->stream));
      k_rows_sorted<256><<<(unsigned)((b->rows + 255) / 256), 256, 0, h->stream>>>(b->rows, b->nnz, b->cols, b->ptr, b->idx, h->d_cnt);
      count_launch(h);
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(&h->h_cnt->unsorted, &h->d_cnt->unsorted, 3 * sizeof(u32), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaMemcpyAsync(&h->h_cnt->spread_sum, &h->d_cnt->spread_sum, sizeof(ull), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      mb->spread_sum = h->h_cnt->spread_sum;
      mb->rows_sorted = h->h_cnt->unsorted ? 0 : 1;
      mb->max_row_len = h->h_cnt->max_rowlen;
      mb->invalid = (int)h->h_cnt->invalid;
    }
  }
  if (b->invalid & 1) return spam_fail(h, SPAM_EINVAL, "row_ptr is not a monotone sequence from 0 to nnz (invariants 3, 4, 7)");
  if (b->invalid & 2) return spam_fail(h, SPAM_EINDEX, "a column index is >= cols (invariant 5)");
  return SPAM_OK;
}

}  // namespace

u64 spgemm_pending_nnz(const SpgemmPending* p) { return p->nnz; }
const u64* spgemm_pending_cptr(const SpgemmPending* p) { return p->d_cptr; }

void spgemm_pending_free(spam_handle* h, SpgemmPending* p) {
  if (!p) return;
  dev_free(h, p->d_flop);
  dev_free(h, p->d_row_nnz);
  dev_free(h, p->d_cptr);
  delete p;
}

int ensure_matrix_stats(spam_handle* h, const spam_dcsr* m) { return ensure_rows_sorted(h, m); }

int flop_count_dev(spam_handle* h, const spam_dcsr* a, const spam_dcsr* b, u32* d_flop, bool do_bins, int mode) {
  const u64 m = a->rows;
  if (m == 0) return SPAM_OK;
  k_flop_count<256><<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(m, b->rows, a->ptr, a->idx, b->ptr, d_flop,
                                                                         h->d_cnt, do_bins ? 1 : 0, mode);
  count_launch(h);
  CK(cudaGetLastError());
  return SPAM_OK;
}

// Phase 1: flop count + binning, symbolic per bin, scan -> row_ptr(C), nnz(C), numeric bin histogram.
int spgemm_symbolic_dev(spam_handle* h, const spam_dcsr* a, const spam_dcsr* b, SpgemmPending** out) {
  *out = nullptr;
  if (a->dtype != b->dtype) return spam_fail(h, SPAM_EDTYPE, "operand dtypes differ");
  if (a->cols != b->rows) return spam_fail(h, SPAM_EDIM, "A.cols != B.rows");
  const u64 m = a->rows;
  if (m >= 0xFFFFFFFFull || b->cols >= 0xFFFFFFFFull || a->cols >= 0xFFFFFFFFull)
    return spam_fail(h, SPAM_ECOLS, "dimension >= 2^32-1 (u32::MAX is the empty-slot sentinel)");

  h->stats = spam_stats{};
  CKS(ensure_rows_sorted(h, b));  // one cached pass per matrix: sortedness, longest row, CSR invariants
  if (a != b) CKS(ensure_rows_sorted(h, a));
  // B with unsorted rows (CsrMatrix<T, false>): multiply by its cached copy with sorted rows instead.  Every
  // entry of C sums one product per entry of A's row, in A's order, whatever the order inside B's rows, so the
  // result is the same bit for bit — and the merge bin and the DIRECT enumeration apply (Poisson with shuffled
  // rows: 3.1 ms through the thread-per-row hash bin, 0.6 ms like this).
  const spam_dcsr* b_orig = b;
  if (b->rows_sorted == 0 && b->nnz < 0xFFFFFFFFull && h->sort_b) {
    const spam_dcsr* sb = nullptr;
    CKS(sorted_rows_of(h, b, &sb));
    h->stats = spam_stats{};
    b = sb;
  }
  const int merge_ok = (b->rows_sorted == 1 && b->nnz < 0xFFFFFFFFull) ? 1 : 0;
  SpgemmPending* p = new SpgemmPending();
  p->a = a; p->b = b; p->b_orig = b_orig; p->d_flop = nullptr; p->d_row_nnz = nullptr; p->d_cptr = nullptr; p->nnz = 0; p->max_nnz = 0;
  // use_esc = 3: the merge-tree bins (msort.cuh) need B's rows sorted, like the merge bin
  const bool esc_on = h->use_esc == 3 ? merge_ok != 0 : h->use_esc != 0;
  const int mode = (merge_ok ? MODE_MERGE : 0) | (esc_on ? MODE_ESC : 0) | (h->use_esc == 2 ? MODE_ESC_HEAVY : 0);
  p->max_alen = 0; p->merge_ok = merge_ok; p->mode = mode;
  *out = p;
#define FAIL_FREE(expr) do { int _s = (expr); if (_s != SPAM_OK) { spgemm_pending_free(h, p); *out = nullptr; return _s; } } while (0)
#define CK_FREE(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { spgemm_pending_free(h, p); *out = nullptr; return spam_fail(h, SPAM_ECUDA, #call, _e); } } while (0)

  FAIL_FREE(dev_alloc_t(h, &p->d_flop, m));
  FAIL_FREE(dev_alloc_t(h, &p->d_row_nnz, m));
  FAIL_FREE(dev_alloc_t(h, &p->d_cptr, m + 1));
  CK_FREE(cudaMemsetAsync(h->d_cnt, 0, sizeof(Counters), h->stream));
  timing_begin_product(h);
  if (h->timing) CK_FREE(cudaEventRecord(h->ev[0], h->stream));
  bool fused = false;
  if (merge_ok && m) {
    // B sorted: flop count and the merge-bin symbolic pass in ONE kernel, then a speculative scan —
    // when every row is in the merge bin (stencils) that scan is final and the product needs a
    // single host sync before the numeric pass.
    const u64 amax = a->max_row_len;
    const u32 kk = amax <= 4 ? 4 : amax <= 6 ? 6 : 8;
    p->max_alen = kk;
    const unsigned grid = (unsigned)((m + 127) / 128);
    const bool l2w = l2_window_set(h, b->idx, (size_t)b->nnz * 4);
    if ((h->merge_win & 2) && ((uintptr_t)b->idx & 15) == 0) {
      if (kk == 4)
        k_flop_sym_merge_win<4, 128><<<grid, 128, 0, h->stream>>>(m, b->rows, a->ptr, a->idx, b->ptr, b->idx, b->nnz, p->d_flop, p->d_row_nnz, h->d_cnt);
      else if (kk == 6)
        k_flop_sym_merge_win<6, 128><<<grid, 128, 0, h->stream>>>(m, b->rows, a->ptr, a->idx, b->ptr, b->idx, b->nnz, p->d_flop, p->d_row_nnz, h->d_cnt);
      else
        k_flop_sym_merge_win<8, 128><<<grid, 128, 0, h->stream>>>(m, b->rows, a->ptr, a->idx, b->ptr, b->idx, b->nnz, p->d_flop, p->d_row_nnz, h->d_cnt);
    } else if (merge_prefetch_mode(h, a, b) & 4) {
      if (kk == 4)
        k_flop_sym_merge<4, 128, 1><<<grid, 128, 0, h->stream>>>(m, b->rows, a->ptr, a->idx, b->ptr, b->idx, p->d_flop, p->d_row_nnz, h->d_cnt);
      else if (kk == 6)
        k_flop_sym_merge<6, 128, 1><<<grid, 128, 0, h->stream>>>(m, b->rows, a->ptr, a->idx, b->ptr, b->idx, p->d_flop, p->d_row_nnz, h->d_cnt);
      else
        k_flop_sym_merge<8, 128, 1><<<grid, 128, 0, h->stream>>>(m, b->rows, a->ptr, a->idx, b->ptr, b->idx, p->d_flop, p->d_row_nnz, h->d_cnt);
    } else if (kk == 4)
      k_flop_sym_merge<4, 128><<<grid, 128, 0, h->stream>>>(m, b->rows, a->ptr, a->idx, b->ptr, b->idx, p->d_flop, p->d_row_nnz, h->d_cnt);
    else if (kk == 6)
      k_flop_sym_merge<6, 128><<<grid, 128, 0, h->stream>>>(m, b->rows, a->ptr, a->idx, b->ptr, b->idx, p->d_flop, p->d_row_nnz, h->d_cnt);
    else
      k_flop_sym_merge<8, 128><<<grid, 128, 0, h->stream>>>(m, b->rows, a->ptr, a->idx, b->ptr, b->idx, p->d_flop, p->d_row_nnz, h->d_cnt);
    count_launch(h);
    if (l2w) l2_window_clear(h);
    CK_FREE(cudaGetLastError());
    if (h->timing) { CK_FREE(cudaEventRecord(h->ev[1], h->stream)); CK_FREE(cudaEventRecord(h->ev[2], h->stream)); }
    FAIL_FREE(scan_u32_to_u64(h, p->d_row_nnz, p->d_cptr, m, &h->d_cnt->total_nnz));
    fused = true;
  } else {
    FAIL_FREE(flop_count_dev(h, a, b, p->d_flop, true, mode));
    if (h->timing) CK_FREE(cudaEventRecord(h->ev[1], h->stream));
  }
  CK_FREE(cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  if (fused && h->timing) CK_FREE(cudaEventRecord(h->ev[3], h->stream));
  CK_FREE(cudaStreamSynchronize(h->stream));
  // (the previous product finished before this sync: its phase events are read back at the end of the numeric
  // phase, while the GPU is busy, not here where it would wait for the host)
  const Counters c1 = *h->h_cnt;
  if (c1.error & 1u) {
    spgemm_pending_free(h, p); *out = nullptr;
    return spam_fail(h, SPAM_EINDEX, "a column index of A is >= rows(B)");
  }
  h->stats.flops = c1.total_flops;
  for (int i = 0; i < NBINS; ++i) h->stats.sym_bin_rows[i] = c1.sym_bins[i];
  if (fused && c1.sym_bins[MERGE_BIN] == m) {  // every row merged: the speculative scan stands
    p->nnz = c1.total_nnz;
    p->max_nnz = 0;
    for (int i = 0; i < NBINS; ++i) { p->num_counts[i] = c1.num_bins[i]; h->stats.num_bin_rows[i] = c1.num_bins[i]; }
    h->stats.nnz_c = p->nnz;
    return SPAM_OK;
  }

  // ---- symbolic per bin ----
  Bins sb;
  FAIL_FREE(build_perm(h, m, c1.sym_bins, false, a->ptr, nullptr, p->d_flop, mode, &sb));
  if (!fused) p->max_alen = c1.max_alen;
  u32* heavy_tab = nullptr;
  {
    const u64* ap = a->ptr; const u32* ac = a->idx; const u64* bp = b->ptr; const u32* bc = b->idx;
    const u32* fl = p->d_flop; u32* rz = p->d_row_nnz;
    auto seg = [&](int bin) -> const u32* { return sb.perm ? sb.perm + sb.base[bin] : nullptr; };
    // bins 4.. run on the side lanes (disjoint rows); the heavy bin runs alone after the join
    const int ws = wshift_for(b, 5);
    bool side = false;
    for (int bin = 4; bin <= NHASH; ++bin) side = side || sb.count[bin] != 0;
    u64 heavy_nblk = 0, heavy_stride = 0;
    if (sb.count[HEAVY_BIN]) {
      u32 fmax = c1.max_flop;
      if (fmax > (u32)b->cols) fmax = (u32)b->cols;
      heavy_stride = 2ull * npow2_u64(fmax);
      heavy_nblk = (u64)h->num_sms * 2;
      if (heavy_nblk > sb.count[HEAVY_BIN]) heavy_nblk = sb.count[HEAVY_BIN];
      const u64 budget = 8ull << 30;
      while (heavy_nblk > 1 && heavy_nblk * heavy_stride * sizeof(u32) > budget) heavy_nblk /= 2;
      FAIL_FREE(dev_alloc_t(h, &heavy_tab, heavy_nblk * heavy_stride));
    }
    if (side) CK_FREE(lanes_fork(h));
#define SYM_MODE 0
#define LAUNCH_SYM_ROW(BIN, NW, CAP, DIRECT)                                                             \
    if (sb.count[BIN]) {                                                                                 \
      constexpr size_t smem = sym_row_smem<NW, CAP>();                                                   \
      FAIL_FREE(set_smem(h, k_sym_row<NW, CAP, DIRECT>, smem));                                          \
      const unsigned grid = NW == 1 ? (sb.count[BIN] + ROWS_PER_BLOCK_W1 - 1) / ROWS_PER_BLOCK_W1 : sb.count[BIN]; \
      k_sym_row<NW, CAP, DIRECT><<<grid, NW == 1 ? 32 * ROWS_PER_BLOCK_W1 : 32 * NW, smem, lane_of(h, BIN)>>>( \
          sb.count[BIN], seg(BIN), ap, ac, bp, bc, fl, rz, SYM_MODE);                                    \
      count_launch(h);                                                                                   \
    }
    // symbolic hash bin b: f <= 64 << b, table 128 << b keys (4 B each); expensive bins first
    LAUNCH_SYM_ROW(8, 32, 32768, false)
    LAUNCH_SYM_ROW(7, 16, 16384, false)
    LAUNCH_SYM_ROW(6, 8, 8192, false)
    LAUNCH_SYM_ROW(5, 4, 4096, false)
    if (direct_enumeration(b)) {
      LAUNCH_SYM_ROW(1, 1, 256, true)
      LAUNCH_SYM_ROW(2, 1, 512, true)
      LAUNCH_SYM_ROW(3, 1, 1024, true)
      // f in (512, 1024]: the table f asks for (8 KB per row) leaves 28 warps per SM.  Regular matrices compress
      // well (27-point stencil: 729 products, 125 columns): try a 512-key table first, redo the rows that
      // pass 256 distinct columns with the full-size one (same stream, so the order holds).
#undef SYM_MODE
#define SYM_MODE 1
      LAUNCH_SYM_ROW(4, 1, 512, true)
#undef SYM_MODE
#define SYM_MODE 2
      LAUNCH_SYM_ROW(4, 1, 2048, true)
#undef SYM_MODE
#define SYM_MODE 0
    } else {
      LAUNCH_SYM_ROW(1, 1, 256, false)
      LAUNCH_SYM_ROW(2, 1, 512, false)
      LAUNCH_SYM_ROW(3, 1, 1024, false)
      LAUNCH_SYM_ROW(4, 4, 2048, false)  // 4 rows x 8 KB per block left 28 warps per SM; a team per row runs 64
    }
#undef LAUNCH_SYM_ROW
#undef SYM_MODE
    if (sb.count[MERGE_BIN] && !fused) {
      constexpr int BL = 128;
      const u32 nm = sb.count[MERGE_BIN];
      const unsigned grid = (nm + BL - 1) / BL;
      if (c1.max_alen <= 4)
        k_sym_merge<4, BL><<<grid, BL, 0, h->stream>>>(nm, seg(MERGE_BIN), ap, ac, bp, bc, fl, rz, h->d_cnt);
      else if (c1.max_alen <= 6)
        k_sym_merge<6, BL><<<grid, BL, 0, h->stream>>>(nm, seg(MERGE_BIN), ap, ac, bp, bc, fl, rz, h->d_cnt);
      else
        k_sym_merge<8, BL><<<grid, BL, 0, h->stream>>>(nm, seg(MERGE_BIN), ap, ac, bp, bc, fl, rz, h->d_cnt);
      count_launch(h);
    }
    if (sb.count[0]) {
      constexpr int BL = 128;
      const size_t smem = (size_t)2 * SYM_TINY_MAX * BL * sizeof(u32);
      k_sym_tiny<BL><<<(sb.count[0] + BL - 1) / BL, BL, smem, h->stream>>>(sb.count[0], seg(0), ap, ac, bp, bc, fl, rz);
      count_launch(h);
    }
    CK_FREE(cudaGetLastError());
    if (side) CK_FREE(lanes_join(h));
    if (sb.count[HEAVY_BIN]) {
      k_sym_heavy<1024><<<(unsigned)heavy_nblk, 1024, 0, h->stream>>>(sb.count[HEAVY_BIN], seg(HEAVY_BIN), ap, ac, bp, bc, fl,
                                                                                 rz, heavy_tab, heavy_stride, (u32)b->cols,
                                                                                 &h->d_cnt->work_a, ws);
      count_launch(h);
      CK_FREE(cudaGetLastError());
    }
  }
  if (h->timing) CK_FREE(cudaEventRecord(h->ev[2], h->stream));
  // ---- numeric bin histogram + row_ptr scan ----
  if (m && sb.count[MERGE_BIN] != m) {  // the merge kernel histograms its own rows
    k_num_bin_count<256><<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(m, a->ptr, p->d_row_nnz, p->d_flop, h->d_cnt,
                                                                            mode);
    count_launch(h);
    CK_FREE(cudaGetLastError());
  }
  FAIL_FREE(scan_u32_to_u64(h, p->d_row_nnz, p->d_cptr, m, &h->d_cnt->total_nnz));
  CK_FREE(cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  if (h->timing) CK_FREE(cudaEventRecord(h->ev[3], h->stream));
  if (heavy_tab) FAIL_FREE(dev_free(h, heavy_tab));
  if (sb.perm) FAIL_FREE(dev_free(h, sb.perm));
  CK_FREE(cudaStreamSynchronize(h->stream));
  const Counters c2 = *h->h_cnt;
  p->nnz = c2.total_nnz;
  p->max_nnz = c2.max_nnz;
  for (int i = 0; i < NBINS; ++i) { p->num_counts[i] = c2.num_bins[i]; h->stats.num_bin_rows[i] = c2.num_bins[i]; }
  h->stats.nnz_c = p->nnz;
  return SPAM_OK;
#undef FAIL_FREE
#undef CK_FREE
}

namespace {

template <class V>
int numeric_typed(spam_handle* h, SpgemmPending* p, spam_dcsr* c) {
  const spam_dcsr* a = p->a;
  const spam_dcsr* b = p->b;
  const u64 m = a->rows;
  Bins nb;
  // the symbolic phase's memset zeroed these, but another product may have run on this handle between the two
  // host phases (spam_spgemm_symbolic ... spam_spgemm_dev ... spam_spgemm_numeric)
  CK(cudaMemsetAsync(h->d_cnt->num_cursor, 0, sizeof(u32) * NBINS, h->stream));
  CK(cudaMemsetAsync(&h->d_cnt->work_a, 0, 3 * sizeof(u32), h->stream));  // work_a, work_b, work_c: row queues
  CK(cudaMemsetAsync(&h->d_cnt->fb_esc, 0, 2 * sizeof(u32), h->stream));   // fb_esc, fb_list_n
  DevGuard g(h);
  CKS(build_perm(h, m, p->num_counts, true, a->ptr, p->d_row_nnz, p->d_flop, p->mode, &nb));
  if (nb.perm) g.owned.push_back(nb.perm);
  const u64* ap = a->ptr; const u32* ac = a->idx; const V* av = (const V*)a->val;
  const u64* bp = b->ptr; const u32* bc = b->idx; const V* bv = (const V*)b->val;
  const u64* cp = c->ptr; u32* cc = c->idx; V* cv = (V*)c->val;
  auto seg = [&](int bin) -> const u32* { return nb.perm ? nb.perm + nb.base[bin] : nullptr; };
#define LAUNCH_NUM_ROW(BIN, NW, CAP, DIRECT)                                                             \
  if (nb.count[BIN]) {                                                                                   \
    constexpr size_t smem = num_row_smem<V, NW, CAP>();                                                  \
    CKS(set_smem(h, k_num_row<V, NW, CAP, DIRECT>, smem));                                               \
    const unsigned grid = NW == 1 ? (nb.count[BIN] + ROWS_PER_BLOCK_W1 - 1) / ROWS_PER_BLOCK_W1 : nb.count[BIN]; \
    /* NW = 1 sorts (column << log2(CAP/2)) | index packed in 32 bits when the columns are narrow enough */ \
    const int pack_ok = b->cols < (1ull << (32 - (31 - __builtin_clz((unsigned)(CAP) / 2)))) ? 1 : 0;   \
    k_num_row<V, NW, CAP, DIRECT><<<grid, NW == 1 ? 32 * ROWS_PER_BLOCK_W1 : 32 * NW, smem, lane_of(h, BIN)>>>( \
        nb.count[BIN], seg(BIN), ap, ac, av, bp, bc, bv, cp, cc, cv, pack_ok, h->d_cnt);                 \
    count_launch(h);                                                                                     \
  }
  // The bins touch disjoint rows of C.  The team bins (largest first) go to the side lanes so that the tail of
  // one (a few blocks still on their last rows) overlaps the next; the heavy bin runs alone at the end.
  bool side = n_esc_rows(nb.count) != 0;
  for (int bin = 4; bin <= NHASH; ++bin) side = side || nb.count[bin] != 0;
  u32 *hk = nullptr, *hc = nullptr, *ho = nullptr;
  V* hv = nullptr;
  u64 heavy_nblk = 0, heavy_stride = 0;
  u32 n_esc = 0;  // rows in the bucket-sort bins: any of them may be handed back to the global-table kernel
  for (int bin = ESC_BIN0; bin <= ESC_HEAVY_BIN; ++bin) n_esc += nb.count[bin];
  u32* fb_list = nullptr;
  if (n_esc) CKS(g.alloc(&fb_list, n_esc));
  if (nb.count[HEAVY_BIN] || n_esc) {
    u32 zmax = p->max_nnz;
    heavy_stride = 2ull * npow2_u64(zmax);
    if (heavy_stride < 8192) heavy_stride = 8192;  // k_num_heavy's bucket counters: at least 4096 + 4
    heavy_nblk = (u64)h->num_sms * 2;
    if (heavy_nblk > (u64)nb.count[HEAVY_BIN] + n_esc) heavy_nblk = (u64)nb.count[HEAVY_BIN] + n_esc;
    const u64 budget = 16ull << 30;
    // per block: keys + values (stride each), bucket counters (stride/2), bucket-ordered keys + slots (stride)
    while (heavy_nblk > 1 && heavy_nblk * heavy_stride * (2 * sizeof(u32) + sizeof(V) + 2) > budget) heavy_nblk /= 2;
    CKS(g.alloc(&hk, heavy_nblk * heavy_stride));
    CKS(g.alloc(&hv, heavy_nblk * heavy_stride));
    CKS(g.alloc(&hc, heavy_nblk * (heavy_stride / 2 + 4)));
    CKS(g.alloc(&ho, heavy_nblk * heavy_stride));
  }
  if (side) CK(lanes_fork(h));
  // bucket-sort bins (esc.cuh): rows that do not compress.  The column-range kernel for the longest rows goes
  // first: its persistent blocks take a whole SM each and its rows are the longest-running items of the product.
  if (nb.count[ESC_HEAVY_BIN]) {
    constexpr size_t smem = num_esc_heavy_smem<V>();
    CKS(set_smem(h, k_num_esc_heavy<V>, smem));
    unsigned grid = (unsigned)h->num_sms;
    if (grid > nb.count[ESC_HEAVY_BIN]) grid = nb.count[ESC_HEAVY_BIN];
    k_num_esc_heavy<V><<<grid, ESCH_T, smem, lane_of(h, ESC_HEAVY_BIN)>>>(nb.count[ESC_HEAVY_BIN], seg(ESC_HEAVY_BIN), ap, ac, av, bp, bc,
                                                                        bv, cp, cc, cv, (u32)b->cols, &h->d_cnt->work_a,
                                                                        h->d_cnt, fb_list);
    count_launch(h);
  }
#define LAUNCH_ESC(BIN, NW)                                                                              \
  if (nb.count[BIN]) {                                                                                   \
    constexpr size_t smem = num_esc_smem<V, NW>();                                                       \
    CKS(set_smem(h, k_num_esc<V, NW>, smem));                                                            \
    k_num_esc<V, NW><<<nb.count[BIN], 32 * NW, smem, lane_of(h, BIN)>>>(nb.count[BIN], seg(BIN), ap, ac, av, bp, bc, bv, \
                                                                       cp, cc, cv, h->d_cnt, fb_list);   \
    count_launch(h);                                                                                     \
  }
#define LAUNCH_MSORT(BIN, NW)                                                                            \
  if (nb.count[BIN]) {                                                                                   \
    constexpr size_t smem = num_msort_smem<V, NW>();                                                     \
    CKS(set_smem(h, k_num_msort<V, NW>, smem));                                                          \
    k_num_msort<V, NW><<<nb.count[BIN], 32 * NW, smem, lane_of(h, BIN)>>>(nb.count[BIN], seg(BIN), ap, ac, av, bp, bc, \
                                                                         bv, cp, cc, cv);                \
    count_launch(h);                                                                                     \
  }
  if (h->use_esc == 3) {
    LAUNCH_MSORT(ESC_BIN0 + 3, 32)
    LAUNCH_MSORT(ESC_BIN0 + 2, 16)
    LAUNCH_MSORT(ESC_BIN0 + 1, 8)
    LAUNCH_MSORT(ESC_BIN0, 4)
  } else {
    LAUNCH_ESC(ESC_BIN0 + 3, 32)
    LAUNCH_ESC(ESC_BIN0 + 2, 16)
    LAUNCH_ESC(ESC_BIN0 + 1, 8)
    LAUNCH_ESC(ESC_BIN0, 4)
  }
#undef LAUNCH_MSORT
#undef LAUNCH_ESC
  // numeric hash bin b: z <= 32 << b, table 64 << b (key, value) slots.  Team sizes: these kernels are
  // latency-bound (dependent shared-memory and shuffle chains), so the big-table bins get many warps per row
  LAUNCH_NUM_ROW(8, 32, 16384, false)
  LAUNCH_NUM_ROW(7, 24, 8192, false)
  LAUNCH_NUM_ROW(6, 16, 4096, false)
  LAUNCH_NUM_ROW(5, 8, 2048, false)
  if (direct_enumeration(b)) {
    LAUNCH_NUM_ROW(4, 1, 1024, true)
    LAUNCH_NUM_ROW(3, 1, 512, true)
    LAUNCH_NUM_ROW(2, 1, 256, true)
    LAUNCH_NUM_ROW(1, 1, 128, true)
  } else {
    // z in (256, 512]: one warp per row leaves 16 warps per SM (4 rows x 12 KB per block); a 4-warp team with the
    // bucket drain runs 48
    LAUNCH_NUM_ROW(4, 4, 1024, false)
    LAUNCH_NUM_ROW(3, 1, 512, false)
    LAUNCH_NUM_ROW(2, 1, 256, false)
    LAUNCH_NUM_ROW(1, 1, 128, false)
  }
#undef LAUNCH_NUM_ROW
  // SPAM_L2_PERSIST = 1: B's col_idx, 2: B's values persisting in L2 during the merge kernel
  const bool l2_window = nb.count[MERGE_BIN] &&
                         (h->l2_persist % 10 == 2 ? l2_window_set(h, bv, (size_t)b->nnz * sizeof(V)) : l2_window_set(h, bc, (size_t)b->nnz * 4));
  if (nb.count[MERGE_BIN]) {
    constexpr int BL = 128;
    constexpr size_t smem = num_merge_smem<V, BL>();
    const u32 nm = nb.count[MERGE_BIN];
    unsigned grid = (nm + BL - 1) / BL;
    const u32 kmax = p->max_alen;  // longest A row among all rows short enough for a merge bin
    if (h->merge_persist > 0) {  // persistent grid: merge_persist blocks per SM walk the row blocks
      const unsigned pg = (unsigned)h->num_sms * (unsigned)h->merge_persist;
      if (pg < grid) grid = pg;
    }
    const int mpf = merge_prefetch_mode(h, a, b);
    if ((h->merge_win & 1) && (((uintptr_t)bc | (uintptr_t)bv) & 15) == 0) {
      constexpr size_t wsmem = num_merge_win_smem<V, BL>();
      if (kmax <= 4) {
        CKS(set_smem(h, k_num_merge_win<V, 4, BL>, wsmem));
        k_num_merge_win<V, 4, BL><<<grid, BL, wsmem, h->stream>>>(nm, seg(MERGE_BIN), ap, ac, av, bp, bc, bv, b->nnz, cp, cc, cv);
      } else if (kmax <= 6) {
        CKS(set_smem(h, k_num_merge_win<V, 6, BL>, wsmem));
        k_num_merge_win<V, 6, BL><<<grid, BL, wsmem, h->stream>>>(nm, seg(MERGE_BIN), ap, ac, av, bp, bc, bv, b->nnz, cp, cc, cv);
      } else {
        CKS(set_smem(h, k_num_merge_win<V, 8, BL>, wsmem));
        k_num_merge_win<V, 8, BL><<<grid, BL, wsmem, h->stream>>>(nm, seg(MERGE_BIN), ap, ac, av, bp, bc, bv, b->nnz, cp, cc, cv);
      }
    } else if ((mpf & 3) == 1) {
      if (kmax <= 4)
        k_num_merge<V, 4, BL, 1><<<grid, BL, smem, h->stream>>>(nm, seg(MERGE_BIN), ap, ac, av, bp, bc, bv, cp, cc, cv);
      else if (kmax <= 6)
        k_num_merge<V, 6, BL, 1><<<grid, BL, smem, h->stream>>>(nm, seg(MERGE_BIN), ap, ac, av, bp, bc, bv, cp, cc, cv);
      else
        k_num_merge<V, 8, BL, 1><<<grid, BL, smem, h->stream>>>(nm, seg(MERGE_BIN), ap, ac, av, bp, bc, bv, cp, cc, cv);
    } else if ((mpf & 3) == 2) {
      if (kmax <= 4)
        k_num_merge<V, 4, BL, 2><<<grid, BL, smem, h->stream>>>(nm, seg(MERGE_BIN), ap, ac, av, bp, bc, bv, cp, cc, cv);
      else if (kmax <= 6)
        k_num_merge<V, 6, BL, 2><<<grid, BL, smem, h->stream>>>(nm, seg(MERGE_BIN), ap, ac, av, bp, bc, bv, cp, cc, cv);
      else
        k_num_merge<V, 8, BL, 2><<<grid, BL, smem, h->stream>>>(nm, seg(MERGE_BIN), ap, ac, av, bp, bc, bv, cp, cc, cv);
    } else if (kmax <= 4)
      k_num_merge<V, 4, BL><<<grid, BL, smem, h->stream>>>(nm, seg(MERGE_BIN), ap, ac, av, bp, bc, bv, cp, cc, cv);
    else if (kmax <= 6)
      k_num_merge<V, 6, BL><<<grid, BL, smem, h->stream>>>(nm, seg(MERGE_BIN), ap, ac, av, bp, bc, bv, cp, cc, cv);
    else
      k_num_merge<V, 8, BL><<<grid, BL, smem, h->stream>>>(nm, seg(MERGE_BIN), ap, ac, av, bp, bc, bv, cp, cc, cv);
    count_launch(h);
    if (l2_window) l2_window_clear(h);
  }
  if (nb.count[0]) {
    constexpr int BL = 128;
    constexpr size_t smem = num_tiny_smem<V, BL>();
    CKS(set_smem(h, k_num_tiny<V, BL>, smem));
    k_num_tiny<V, BL><<<(nb.count[0] + BL - 1) / BL, BL, smem, h->stream>>>(nb.count[0], seg(0), ap, ac, av, bp, bc, bv, cp, cc, cv);
    count_launch(h);
  }
  CK(cudaGetLastError());
  if (side) CK(lanes_join(h));
  if (nb.count[HEAVY_BIN]) {
    k_num_heavy<V, 1024><<<(unsigned)heavy_nblk, 1024, 0, h->stream>>>(nb.count[HEAVY_BIN], seg(HEAVY_BIN), ap, ac, av, bp, bc, bv,
                                                                                  cp, cc, cv, hk, hv, heavy_stride, hc, ho,
                                                                                  (u32)b->cols, &h->d_cnt->work_b, h->d_cnt, nullptr);
    count_launch(h);
    CK(cudaGetLastError());
  }
  if (n_esc) {
    // rows the bucket-sort bins handed back (crowded buckets): usually none, then the blocks exit at once
    k_num_heavy<V, 1024><<<(unsigned)heavy_nblk, 1024, 0, h->stream>>>(0u, fb_list, ap, ac, av, bp, bc, bv, cp, cc, cv, hk, hv,
                                                                                  heavy_stride, hc, ho, (u32)b->cols,
                                                                                  &h->d_cnt->work_c, h->d_cnt, &h->d_cnt->fb_list_n);
    count_launch(h);
    CK(cudaGetLastError());
  }
  return SPAM_OK;  // the guard frees the tables and the permutation (stream-ordered: after the kernels above)
}

}  // namespace

namespace {
int numeric_dispatch(spam_handle* h, SpgemmPending* p, spam_dcsr* c) {
  switch (c->dtype) {
    case SPAM_F32: return numeric_typed<float>(h, p, c);
    case SPAM_F64: return numeric_typed<double>(h, p, c);
    case SPAM_I32: return numeric_typed<int32_t>(h, p, c);
    case SPAM_I64: return numeric_typed<int64_t>(h, p, c);
    default: return spam_fail(h, SPAM_EINVAL, "bad dtype");
  }
}
int numeric_timing(spam_handle* h) {
  if (!h->timing) return SPAM_OK;
  cudaError_t e = cudaEventRecord(h->ev[4], h->stream);
  if (e != cudaSuccess) return spam_fail(h, SPAM_ECUDA, "cudaEventRecord", e);
  h->ev_pending[h->ev_cur] = true;
  timing_harvest(h, h->ev_cur ^ 1);  // previous product: complete since this product's symbolic sync
  return SPAM_OK;
}
}  // namespace

// Phase 2: allocate C (exact nnz, like Vec::with_capacity(nnz), mul_hash.rs:119), numeric per bin.
// Consumes the pending state.  On success *cout owns col_idx/val and takes over row_ptr.
int spgemm_numeric_dev(spam_handle* h, SpgemmPending* p, spam_dcsr** cout, int sorted) {
  *cout = nullptr;
  spam_dcsr* c = new spam_dcsr();
  c->dtype = p->a->dtype; c->rows = p->a->rows; c->cols = p->b->cols; c->nnz = p->nnz;
  c->ptr = p->d_cptr; c->idx = nullptr; c->val = nullptr; c->owning = true; c->rows_sorted = -1; c->max_row_len = 0;  // stats are taken lazily if C becomes a right-hand side
  int st = dev_alloc_t(h, &c->idx, p->nnz ? p->nnz : 1);
  if (st == SPAM_OK) st = dev_alloc(h, &c->val, (p->nnz ? p->nnz : 1) * dtype_size(c->dtype));
  if (st == SPAM_OK && p->nnz) st = numeric_dispatch(h, p, c);
  if (st == SPAM_OK && !sorted) st = slot_order_dev(h, p->a, p->b_orig, c);  // B2 = false: the reference's slot order
  if (st == SPAM_OK) st = numeric_timing(h);
  if (st != SPAM_OK) {
    dev_free(h, c->idx); dev_free(h, c->val);
    delete c;
    spgemm_pending_free(h, p);
    return st;
  }
  p->d_cptr = nullptr;  // ownership moved to C
  spgemm_pending_free(h, p);
  *cout = c;
  return SPAM_OK;
}

// The whole product in ONE kernel (k_merge_onepass, merge.cuh) when the cached per-matrix statistics alone prove
// that every row is a merge row: longest row of A <= MERGE_K and longest row of A x longest row of B <=
// MERGE_FLOP_MAX (the binning rule of the two-phase pipeline, applied to the bound instead of the count), B's rows
// sorted.  C's arrays are allocated for the bound nnz(A) x longest row of B (stencils: about 2 x nnz(C)) — the
// price for not knowing nnz(C) before the kernel runs; products over 8 GB of bound take the two-phase pipeline.
// *cout stays null when the operands do not qualify (not an error).
// Measured on B200 (Poisson 2048^2 A*A f64): the kernel takes 0.558 ms against 0.132 + 0.386 ms for the two kernels it
// replaces — the merge work is the same (the columns are still merged twice), what it saves (a pass over A, the scan,
// the mid-product host sync: ~0.05 ms) goes into blocks waiting for their predecessors' totals in the look-back — step
// 0.594 ms against 0.566 ms.  Opt-in (SPAM_ONEPASS=1), kept with its parity test.
template <class V>
static int onepass_launch(spam_handle* h, const spam_dcsr* a, const spam_dcsr* b, spam_dcsr* c, u64* state, u32* counter) {
  constexpr int BL = 128;
  constexpr size_t smem = num_merge_smem<V, BL>();
  const u32 m = (u32)a->rows;
  const unsigned grid = (m + BL - 1) / BL;
  const u64 amax = a->max_row_len;
#define ONEPASS(KK)                                                                                                  \
  k_merge_onepass<V, KK, BL><<<grid, BL, smem, h->stream>>>(m, b->rows, a->ptr, a->idx, (const V*)a->val, b->ptr, b->idx, \
                                                            (const V*)b->val, c->ptr, c->idx, (V*)c->val, state, counter, h->d_cnt)
  if (amax <= 4) ONEPASS(4);
  else if (amax <= 6) ONEPASS(6);
  else ONEPASS(8);
#undef ONEPASS
  count_launch(h);
  CK(cudaGetLastError());
  return SPAM_OK;
}

int spgemm_onepass_dev(spam_handle* h, const spam_dcsr* a, const spam_dcsr* b, spam_dcsr** cout) {
  *cout = nullptr;
  if (!h->onepass || a->dtype != b->dtype || a->cols != b->rows) return SPAM_OK;
  const u64 m = a->rows;
  if (m == 0 || m >= 0xFFFFFFFFull || b->cols >= 0xFFFFFFFFull || a->cols >= 0xFFFFFFFFull) return SPAM_OK;
  CKS(ensure_rows_sorted(h, b));
  if (a != b) CKS(ensure_rows_sorted(h, a));
  if (b->rows_sorted == 0 && b->nnz < 0xFFFFFFFFull && h->sort_b) {
    const spam_dcsr* sb = nullptr;
    CKS(sorted_rows_of(h, b, &sb));
    b = sb;
  }
  if (b->rows_sorted != 1 || b->nnz >= 0xFFFFFFFFull) return SPAM_OK;
  const u64 amax = a->max_row_len, bmax = b->max_row_len;
  if (amax > MERGE_K || amax * bmax > MERGE_FLOP_MAX) return SPAM_OK;
  u64 cap = a->nnz * bmax;
  if (cap == 0) return SPAM_OK;
  const size_t vs = dtype_size(a->dtype);
  if (cap * (4 + vs) > (8ull << 30)) return SPAM_OK;

  h->stats = spam_stats{};
  spam_dcsr* c = new spam_dcsr();
  c->dtype = a->dtype; c->rows = m; c->cols = b->cols; c->nnz = 0;
  c->ptr = nullptr; c->idx = nullptr; c->val = nullptr; c->owning = true; c->rows_sorted = -1; c->max_row_len = 0;
  auto fail = [&](int st) {
    dev_free(h, c->ptr); dev_free(h, c->idx); dev_free(h, c->val);
    delete c;
    return st;
  };
  int st = dev_alloc_t(h, &c->ptr, m + 1);
  if (st == SPAM_OK) st = dev_alloc_t(h, &c->idx, cap);
  if (st == SPAM_OK) st = dev_alloc(h, &c->val, cap * vs);
  if (st != SPAM_OK) return fail(st);
  cudaError_t e = cudaMemsetAsync(h->d_cnt, 0, sizeof(Counters), h->stream);
  if (e != cudaSuccess) return fail(spam_fail(h, SPAM_ECUDA, "cudaMemsetAsync", e));
  timing_begin_product(h);
  if (h->timing)
    for (int i = 0; i < 4; ++i) cudaEventRecord(h->ev[i], h->stream);  // no separate flop / symbolic / scan phases
  u64* state = nullptr;
  u32* counter = nullptr;
  st = lookback_workspace(h, (m + 127) / 128, &state, &counter);
  if (st != SPAM_OK) return fail(st);
  switch (a->dtype) {
    case SPAM_F32: st = onepass_launch<float>(h, a, b, c, state, counter); break;
    case SPAM_F64: st = onepass_launch<double>(h, a, b, c, state, counter); break;
    case SPAM_I32: st = onepass_launch<int32_t>(h, a, b, c, state, counter); break;
    case SPAM_I64: st = onepass_launch<int64_t>(h, a, b, c, state, counter); break;
    default: st = spam_fail(h, SPAM_EINVAL, "bad dtype");
  }
  if (st == SPAM_OK) st = numeric_timing(h);
  if (st != SPAM_OK) return fail(st);
  e = cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return fail(spam_fail(h, SPAM_ECUDA, "one-pass product", e));
  const Counters cn = *h->h_cnt;
  if (cn.error & 1u) return fail(spam_fail(h, SPAM_EINDEX, "a column index of A is >= rows(B)"));
  c->nnz = cn.total_nnz;
  h->stats.flops = cn.total_flops;
  h->stats.nnz_c = c->nnz;
  h->stats.sym_bin_rows[MERGE_BIN] = (u32)m;
  h->stats.num_bin_rows[MERGE_BIN] = (u32)m;
  h->stats.fallbacks[5] = 1;  // the one-pass kernel ran
  *cout = c;
  return SPAM_OK;
}

// Phase 2 into arrays the caller owns: c_ptr has rows + 1 entries whose VALUES are positions in c_idx / c_val
// (a rank of a row-sharded product passes its offset-fixed row_ptr and the arrays of the whole C, so its rows
// land where the gathered result wants them: disjoint slices of one output, mul_hash.rs:121-128).  Consumes the
// pending state.
int spgemm_numeric_into(spam_handle* h, SpgemmPending* p, const u64* c_ptr, u32* c_idx, void* c_val) {
  spam_dcsr c = {};
  c.dtype = p->a->dtype; c.rows = p->a->rows; c.cols = p->b->cols; c.nnz = p->nnz;
  c.ptr = const_cast<u64*>(c_ptr); c.idx = c_idx; c.val = c_val; c.owning = false; c.rows_sorted = -1;
  int st = SPAM_OK;
  if (p->nnz) st = numeric_dispatch(h, p, &c);
  if (st == SPAM_OK) st = numeric_timing(h);
  spgemm_pending_free(h, p);
  return st;
}
