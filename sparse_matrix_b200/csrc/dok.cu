// dok.cu — DOK -> sorted CSR on the device, as a stable sort plus a segmented "last write wins".
//
// Replaces `impl From<DokMatrix<T>> for CsrMatrix<T, true>` (spam_csr/src/lib.rs:315-334) fed by
// a stream of DokMatrix::set_element calls (spam_dok/src/lib.rs:167-176): inserting a non-zero
// replaces the entry, inserting zero removes it, so the final map holds, per (row, col) key, the
// LAST write of the stream unless that write is zero.  BTreeMap iteration order is (row, col)
// lexicographic (spam_dok/src/lib.rs:234-242); empty rows get repeated offsets (lib.rs:321,325).
//
// The default is the BUCKET PATH of bucket.cuh (one partition pass into buckets of consecutive rows, one build pass per
// bucket in shared memory, one host sync; dok_bucket / transpose_bucket below).  It hands shapes it does not take, crowded
// buckets and very long rows to the two paths of this file, chosen after a histogram of the stream by row (one small host
// sync):
//
//  COUNTING PATH (no row holds more than SEG_MAX = 32 triplets — C5: 8 per row): CSR is a counting sort by row,
//    1. k_dok_hist     raw count per row (L2 atomics), index validation
//    2. look-back scan (scan.cu) -> start of every row's segment
//    3. k_dok_scatter  (col, stream position, value) into the row's segment, order inside it arbitrary
//    4. k_dok_seg      one thread per row: an entry survives iff no entry of the segment has the same column and a
//                      later stream position, and its value is non-zero; survivors as a 32-bit mask + their count
//    5. look-back scan -> row_ptr, nnz            [host sync: the caller allocates exactly nnz]
//    6. k_dok_emit     one thread per row: every survivor ranks itself by column among the survivors
//   The stream is read twice (16 B + 16 B + value), the segments written once and read twice: about 1.7x the
//   algorithmic bytes, against 5.5x for six radix passes.
//
//  RADIX PATH (some row is longer): the general stable sort,
//    1. key = row << cbits | col (only the bits the shape needs), payload = stream position
//    2. hand-written LSD radix sort, 8 bits per pass, stable => equal keys stay in stream order
//    3. keep entry i iff it ends its key run and its value is non-zero; count kept entries per row
//    4. look-back scans give output positions and row_ptr; scatter col_idx / val
//
// All integer work, HBM/L2-bound.
#include "common.cuh"
#include "bucket.cuh"

namespace {

constexpr int RS_T = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_T * RS_ITEMS;
constexpr int RADIX = 256;

__global__ void __launch_bounds__(256) k_make_keys(u64 n, u64 rows, u64 cols, int cbits, const u64* __restrict__ r,
                                                   const u64* __restrict__ c, u64* __restrict__ keys,
                                                   u32* __restrict__ pay, Counters* cnt) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  bool bad = false;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const u64 ri = r[i], ci = c[i];
    u64 key = 0;
    if (ri < rows && ci < cols) key = (ri << cbits) | ci; else bad = true;  // IndexError (spam_dok lib.rs:168-170)
    keys[i] = key;
    pay[i] = (u32)i;
  }
  if (bad) atomicOr(&cnt->error, 2u);
}

__global__ void __launch_bounds__(RS_T) k_radix_hist(const u64* __restrict__ keys, u64 n, int shift,
                                                     u32* __restrict__ hist, u32 nblocks) {
  __shared__ u32 s_hist[RADIX];
  const int tid = threadIdx.x;
  s_hist[tid] = 0;
  __syncthreads();
  const u64 base = (u64)blockIdx.x * RS_TILE;
#pragma unroll 4
  for (int r = 0; r < RS_ITEMS; ++r) {
    const u64 i = base + (u64)r * RS_T + tid;
    if (i < n) atomicAdd(&s_hist[(keys[i] >> shift) & (RADIX - 1)], 1u);
  }
  __syncthreads();
  hist[(u64)tid * nblocks + blockIdx.x] = s_hist[tid];
}

// Stable scatter: element order inside a tile is (round r, warp, lane); ranks follow it.
__global__ void __launch_bounds__(RS_T) k_radix_scatter(const u64* __restrict__ keys_in, const u32* __restrict__ pay_in,
                                                        u64* __restrict__ keys_out, u32* __restrict__ pay_out, u64 n,
                                                        int shift, const u64* __restrict__ offs, u32 nblocks) {
  // Each warp owns a contiguous 512-key slice of the tile and ranks it alone: per-warp digit counters in
  // shared memory, __match_any_sync for the rank inside a 32-key round, no block barrier inside the
  // loop (the first version synchronised the block three times per round and ran 70x off the roofline).
  // Two block barriers in total: one before the per-digit prefix over warps, one after.
  __shared__ u32 s_cnt[RS_T / 32][RADIX];
  __shared__ u64 s_off[RADIX];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int w = 0; w < RS_T / 32; ++w) s_cnt[w][tid] = 0;
  __syncthreads();
  const u64 wbase = (u64)blockIdx.x * RS_TILE + (u64)wid * (32 * RS_ITEMS);
  u64 key[RS_ITEMS];
  u32 pay[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const u64 i = wbase + (u64)r * 32 + lane;
    const bool valid = i < n;
    key[r] = 0; pay[r] = 0;
    if (valid) { key[r] = keys_in[i]; pay[r] = pay_in[i]; }
  }
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const bool valid = wbase + (u64)r * 32 + lane < n;
    const u32 d = valid ? (u32)((key[r] >> shift) & (RADIX - 1)) : (u32)RADIX;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const u32 before = valid ? s_cnt[wid][d] : 0u;
    rank[r] = before + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
    if (valid && (__ffs(peers) - 1) == lane) s_cnt[wid][d] = before + __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  {
    u32 run = 0;  // thread `tid` = digit: exclusive prefix of the warps' counts
#pragma unroll
    for (int w = 0; w < RS_T / 32; ++w) { const u32 t = s_cnt[w][tid]; s_cnt[w][tid] = run; run += t; }
    s_off[tid] = offs[(u64)tid * nblocks + blockIdx.x];  // global base of digit `tid` for this block
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    if (wbase + (u64)r * 32 + lane < n) {
      const u32 d = (u32)((key[r] >> shift) & (RADIX - 1));
      const u64 pos = s_off[d] + s_cnt[wid][d] + rank[r];
      keys_out[pos] = key[r];
      pay_out[pos] = pay[r];
    }
  }
}

template <class V>
__global__ void __launch_bounds__(256) k_mark_last(u64 n, int cbits, const u64* __restrict__ keys,
                                                   const u32* __restrict__ pay, const V* __restrict__ vals,
                                                   u32* __restrict__ flags, u32* __restrict__ row_cnt) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const u64 k = keys[i];
    const bool last = (i + 1 == n) || (keys[i + 1] != k);
    // num_traits::Zero::is_zero: t == 0 (so -0.0 is zero, NaN is not)
    const bool keep = last && !(vals[pay[i]] == (V)0);
    flags[i] = keep ? 1u : 0u;
    if (keep) atomicAdd(&row_cnt[k >> cbits], 1u);
  }
}

template <class V>
__global__ void __launch_bounds__(256) k_emit(u64 n, int cbits, const u64* __restrict__ keys, const u32* __restrict__ pay,
                                              const V* __restrict__ vals, const u32* __restrict__ flags,
                                              const u64* __restrict__ pos, u32* __restrict__ c_idx, V* __restrict__ c_val) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  const u64 cmask = (1ull << cbits) - 1;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (flags[i]) {
      const u64 p = pos[i];
      c_idx[p] = (u32)(keys[i] & cmask);
      c_val[p] = vals[pay[i]];
    }
  }
}

// ---- counting path ----------------------------------------------------------------------------------
constexpr u32 SEG_MAX = 32;  // longest row / column segment the counting paths sort with one thread

// rank of every triplet inside its row = the row counter's value when the triplet arrived (an atomic WITH return):
// the scatter pass then needs no atomics.  Ranks saturate at 255: a row that long takes the radix path anyway.
__global__ void __launch_bounds__(256) k_dok_hist(u64 n, u64 rows, const u64* __restrict__ r, u32* __restrict__ raw_cnt,
                                                  unsigned char* __restrict__ rank8, Counters* cnt) {
  constexpr int U = 4;  // independent atomics in flight per thread
  const u64 stride = (u64)gridDim.x * blockDim.x;
  bool bad = false;
  for (u64 i0 = (u64)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += U * stride) {
    u64 ri[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const u64 i = i0 + u * stride;
      ri[u] = 0; ok[u] = false;
      if (i < n) {
        ri[u] = r[i];
        ok[u] = ri[u] < rows;   // the column is checked by the scatter pass, which reads it anyway
        bad = bad || !ok[u];    // IndexError (spam_dok lib.rs:168-170)
      }
    }
    u32 rk[U];
#pragma unroll
    for (int u = 0; u < U; ++u) rk[u] = ok[u] ? atomicAdd(&raw_cnt[ri[u]], 1u) : 0u;
#pragma unroll
    for (int u = 0; u < U; ++u) if (ok[u]) rank8[i0 + u * stride] = (unsigned char)(rk[u] > 255u ? 255u : rk[u]);
  }
  if (bad) atomicOr(&cnt->error, 2u);
}

// one 16-byte record per triplet: (column, stream position, value bits) — a single store in the scatter, a single
// load per entry afterwards (scattered 8-byte stores cost a whole 32-byte sector each at the L2: the scatter is
// bound by sector transactions, not bytes)
template <class V>
__device__ __forceinline__ uint4 dok_pack(u32 col, u32 pos, V v) {
  unsigned long long bits = 0;
  memcpy(&bits, &v, sizeof(V));
  return make_uint4(col, pos, (u32)bits, (u32)(bits >> 32));
}
template <class V>
__device__ __forceinline__ V dok_val(const uint4& e) {
  const unsigned long long bits = (unsigned long long)e.z | ((unsigned long long)e.w << 32);
  V v;
  memcpy(&v, &bits, sizeof(V));
  return v;
}

template <class V>
__global__ void __launch_bounds__(256) k_dok_scatter(u64 n, u64 cols, const u64* __restrict__ r, const u64* __restrict__ c,
                                                     const V* __restrict__ v, const unsigned char* __restrict__ rank8,
                                                     const u64* __restrict__ seg_ptr, uint4* __restrict__ ent, Counters* cnt) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  bool bad = false;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const u64 ci = c[i];
    bad = bad || ci >= cols;  // IndexError: reported after the second scan, before anything is returned
    ent[seg_ptr[r[i]] + rank8[i]] = dok_pack<V>((u32)ci, (u32)i, v[i]);
  }
  if (bad) atomicOr(&cnt->error, 2u);
}

constexpr int SEG_REGS = 16;  // segments up to this length are held in registers (C5: 8-9 triplets per row)

// One thread per row.  DokMatrix::set_element semantics over the row's triplets (spam_dok lib.rs:167-176): per
// column the LAST write of the stream decides; a zero deletes (num_traits::Zero::is_zero: -0.0 is zero, NaN is not).
// Output: the row's surviving entries, sorted by column, compacted to the FRONT of its own segment (in place), and
// their count.  Short segments are loaded once into registers (every load instruction of a warp touches 32
// different sectors: the nested loops of the first version re-read the segment len times and were bound by the
// load/store unit).
template <class V>
__global__ void __launch_bounds__(128) k_dok_seg(u64 rows, const u64* __restrict__ seg_ptr, uint4* __restrict__ ent,
                                                 u32* __restrict__ row_cnt) {
  const u64 row = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const u64 lo = seg_ptr[row];
  const u32 len = (u32)(seg_ptr[row + 1] - lo);  // <= SEG_MAX (checked on the host before this path is taken)
  if (len == 0) { row_cnt[row] = 0; return; }
  if (len <= (u32)SEG_REGS) {
    uint4 e[SEG_REGS];
#pragma unroll
    for (int a = 0; a < SEG_REGS; ++a) {
      e[a] = make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
      if ((u32)a < len) e[a] = ent[lo + a];
    }
    u32 mask = 0;
#pragma unroll
    for (int a = 0; a < SEG_REGS; ++a) {
      bool last = (u32)a < len;
#pragma unroll
      for (int b = 0; b < SEG_REGS; ++b) last = last && !((u32)b < len && e[b].x == e[a].x && e[b].y > e[a].y);
      if (last && !(dok_val<V>(e[a]) == (V)0)) mask |= 1u << a;
    }
    // every survivor's rank by column among the survivors, then store it at the front of the segment
#pragma unroll
    for (int a = 0; a < SEG_REGS; ++a) {
      u32 rk = 0;
#pragma unroll
      for (int b = 0; b < SEG_REGS; ++b) rk += ((mask >> b) & 1u) && e[b].x < e[a].x ? 1u : 0u;
      if ((mask >> a) & 1u) ent[lo + rk] = e[a];
    }
    row_cnt[row] = __popc(mask);
    return;
  }
  // 17..32 triplets: the same with the segment re-read from L1
  u32 mask = 0;
  for (u32 a = 0; a < len; ++a) {
    const uint4 ea = ent[lo + a];
    bool last = true;
    for (u32 b = 0; b < len; ++b) {
      const uint4 eb = ent[lo + b];
      last = last && !(eb.x == ea.x && eb.y > ea.y);
    }
    if (last && !(dok_val<V>(ea) == (V)0)) mask |= 1u << a;
  }
  const u32 nk = __popc(mask);
  u32 t = 0;
  for (u32 a = 0; a < len; ++a) {  // compact survivors to the front (stable)
    if ((mask >> a) & 1u) {
      if (a != t) ent[lo + t] = ent[lo + a];
      ++t;
    }
  }
  for (u32 i = 1; i < nk; ++i) {  // insertion sort by column
    const uint4 ke = ent[lo + i];
    u32 j = i;
    while (j > 0 && ent[lo + j - 1].x > ke.x) { ent[lo + j] = ent[lo + j - 1]; --j; }
    ent[lo + j] = ke;
  }
  row_cnt[row] = nk;
}

// copy every row's survivors (front of its segment, already in column order) to their place in C
template <class V>
__global__ void __launch_bounds__(128) k_dok_emit(u64 rows, const u64* __restrict__ seg_ptr, const uint4* __restrict__ ent,
                                                  const u64* __restrict__ c_ptr, u32* __restrict__ c_idx,
                                                  V* __restrict__ c_val) {
  const u64 row = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const u64 o = c_ptr[row];
  const u32 nk = (u32)(c_ptr[row + 1] - o);
  const u64 lo = seg_ptr[row];
  for (u32 a = 0; a < nk; ++a) {
    const uint4 e = ent[lo + a];
    c_idx[o + a] = e.x;
    c_val[o + a] = dok_val<V>(e);
  }
}

// ---- transpose, counting path: histogram by column (each entry keeps its arrival rank), scan, scatter (row, value)
// into the column's segment, then one thread per column puts its segment in increasing row order (rows are
// distinct inside a column) ----
__global__ void __launch_bounds__(256) k_tr_hist(u64 nnz, u64 cols, const u32* __restrict__ idx, u32* __restrict__ col_cnt,
                                                 unsigned char* __restrict__ rank8, Counters* cnt) {
  constexpr int U = 4;
  const u64 stride = (u64)gridDim.x * blockDim.x;
  bool bad = false;
  for (u64 e0 = (u64)blockIdx.x * blockDim.x + threadIdx.x; e0 < nnz; e0 += U * stride) {
    u32 c[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const u64 e = e0 + u * stride;
      c[u] = 0; ok[u] = false;
      if (e < nnz) { c[u] = idx[e]; ok[u] = c[u] < cols; bad = bad || !ok[u]; }
    }
    u32 rk[U];
#pragma unroll
    for (int u = 0; u < U; ++u) rk[u] = ok[u] ? atomicAdd(&col_cnt[c[u]], 1u) : 0u;
#pragma unroll
    for (int u = 0; u < U; ++u) if (ok[u]) rank8[e0 + u * stride] = (unsigned char)(rk[u] > 255u ? 255u : rk[u]);
  }
  if (bad) atomicOr(&cnt->error, 2u);
}

// scatter: one 16-byte record (row, value bits) per entry into its column's segment of a temporary array (a single
// store per entry: scattered stores cost one L2 sector transaction each whatever their size)
template <class W>
__global__ void __launch_bounds__(256) k_tr_scatter(u64 m, const u64* __restrict__ ptr, const u32* __restrict__ idx,
                                                    const W* __restrict__ val, const unsigned char* __restrict__ rank8,
                                                    const u64* __restrict__ t_ptr, uint4* __restrict__ rec) {
  const int lane = threadIdx.x & 31;
  const u64 row = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = row < m;
  u64 lo = 0, hi = 0;
  if (valid) { lo = ptr[row]; hi = ptr[row + 1]; }
  if (valid && hi - lo <= 32) {
    for (u64 e = lo; e < hi; ++e) {
      const unsigned long long bits = (unsigned long long)val[e];
      rec[t_ptr[idx[e]] + rank8[e]] = make_uint4((u32)row, 0u, (u32)bits, (u32)(bits >> 32));
    }
  }
  unsigned longmask = __ballot_sync(0xffffffffu, valid && hi - lo > 32);  // long rows: the whole warp helps
  while (longmask) {
    const int src = __ffs(longmask) - 1;
    longmask &= longmask - 1;
    const u64 l = __shfl_sync(0xffffffffu, lo, src), hh = __shfl_sync(0xffffffffu, hi, src);
    const u32 rr = (u32)__shfl_sync(0xffffffffu, row, src);
    for (u64 e = l + lane; e < hh; e += 32) {
      const unsigned long long bits = (unsigned long long)val[e];
      rec[t_ptr[idx[e]] + rank8[e]] = make_uint4(rr, 0u, (u32)bits, (u32)(bits >> 32));
    }
  }
}

// one thread per column: its records in increasing row order (rows are distinct inside a column; ties — an invalid
// matrix with a repeated column in a row — keep their arrival order) into the final arrays
template <class W>
__global__ void __launch_bounds__(128) k_tr_segsort(u64 cols, const u64* __restrict__ t_ptr, const uint4* __restrict__ rec,
                                                    u32* __restrict__ t_idx, W* __restrict__ t_val) {
  const u64 col = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= cols) return;
  const u64 lo = t_ptr[col], hi = t_ptr[col + 1];  // hi - lo <= SEG_MAX
  const u32 len = (u32)(hi - lo);
  if (len == 0) return;
  auto val_of = [](const uint4& r) { return (W)((unsigned long long)r.z | ((unsigned long long)r.w << 32)); };
  if (len <= (u32)SEG_REGS) {
    uint4 r[SEG_REGS];
#pragma unroll
    for (int a = 0; a < SEG_REGS; ++a) {
      r[a] = make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
      if ((u32)a < len) r[a] = rec[lo + a];
    }
#pragma unroll
    for (int a = 0; a < SEG_REGS; ++a) {
      u32 rk = 0;
#pragma unroll
      for (int b = 0; b < SEG_REGS; ++b) rk += (r[b].x < r[a].x || (r[b].x == r[a].x && b < a)) ? 1u : 0u;  // stable
      if ((u32)a < len) { t_idx[lo + rk] = r[a].x; t_val[lo + rk] = val_of(r[a]); }
    }
    return;
  }
  for (u32 a = 0; a < len; ++a) {  // 17..32 records: rank by re-reading the segment (L1)
    const uint4 ra = rec[lo + a];
    u32 rk = 0;
    for (u32 b = 0; b < len; ++b) { const u32 rb = rec[lo + b].x; rk += (rb < ra.x || (rb == ra.x && b < a)) ? 1u : 0u; }
    t_idx[lo + rk] = ra.x; t_val[lo + rk] = val_of(ra);
  }
}

int bits_for(u64 x) {  // bits needed to represent values in [0, x)
  int b = 0;
  while (b < 63 && (1ull << b) < x) ++b;
  return b;
}

unsigned grid_for(const spam_handle* h, u64 n) {
  u64 g = (n + 255) / 256;
  const u64 cap = (u64)h->num_sms * 16;
  if (g > cap) g = cap;
  if (g == 0) g = 1;
  return (unsigned)g;
}

// stable LSD radix sort of (key, payload) on `keybits` low bits; result may end in either buffer
int radix_sort_pairs(spam_handle* h, u64 n, int keybits, u64*& k0, u32*& p0, u64*& k1, u32*& p1, u32* hist,
                     u64* offs) {
  if (n == 0) return SPAM_OK;
  const u32 nblocks = (u32)((n + RS_TILE - 1) / RS_TILE);
  for (int shift = 0; shift < keybits; shift += 8) {
    k_radix_hist<<<nblocks, RS_T, 0, h->stream>>>(k0, n, shift, hist, nblocks);
    count_launch(h);
    CK(cudaGetLastError());
    CKS(scan_u32_to_u64(h, hist, offs, (u64)RADIX * nblocks, nullptr));
    k_radix_scatter<<<nblocks, RS_T, 0, h->stream>>>(k0, p0, k1, p1, n, shift, offs, nblocks);
    count_launch(h);
    CK(cudaGetLastError());
    u64* tk = k0; k0 = k1; k1 = tk;
    u32* tp = p0; p0 = p1; p1 = tp;
  }
  return SPAM_OK;
}

// the general path: stable LSD radix sort of (row, col) keys, last write of every key run wins
template <class V>
int dok_radix(spam_handle* h, u64 rows, u64 cols, u64 n, const u64* d_r, const u64* d_c, const V* d_v, spam_dcsr* out) {
  const int cbits = bits_for(cols), rbits = bits_for(rows);
  // one workspace allocation, carved up (allocator calls were a visible part of this routine)
  const u64 nblocks = (n + RS_TILE - 1) / RS_TILE;
  auto al = [](u64 bytes) { return (bytes + 255) & ~255ull; };
  const u64 sz_k = al(n * 8), sz_pos = al((n + 1) * 8), sz_p = al(n * 4), sz_rc = al(rows * 4),
            sz_hist = al((u64)RADIX * nblocks * 4), sz_offs = al(((u64)RADIX * nblocks + 1) * 8);
  DevGuard g(h);
  char* ws = nullptr;
  CKS(g.alloc(&ws, 2 * sz_k + sz_pos + 3 * sz_p + sz_rc + sz_hist + sz_offs));
  char* cur = ws;
  u64* k0 = (u64*)cur; cur += sz_k;
  u64* k1 = (u64*)cur; cur += sz_k;
  u64* pos = (u64*)cur; cur += sz_pos;
  u64* offs = (u64*)cur; cur += sz_offs;
  u32* p0 = (u32*)cur; cur += sz_p;
  u32* p1 = (u32*)cur; cur += sz_p;
  u32* flags = (u32*)cur; cur += sz_p;
  u32* row_cnt = (u32*)cur; cur += sz_rc;
  u32* hist = (u32*)cur;
  CK(cudaMemsetAsync(h->d_cnt, 0, sizeof(Counters), h->stream));
  CK(cudaMemsetAsync(row_cnt, 0, rows * sizeof(u32), h->stream));
  if (n) {
    k_make_keys<<<grid_for(h, n), 256, 0, h->stream>>>(n, rows, cols, cbits, d_r, d_c, k0, p0, h->d_cnt);
    count_launch(h);
    CK(cudaGetLastError());
    CKS(radix_sort_pairs(h, n, cbits + rbits, k0, p0, k1, p1, hist, offs));
    k_mark_last<V><<<grid_for(h, n), 256, 0, h->stream>>>(n, cbits, k0, p0, d_v, flags, row_cnt);
    count_launch(h);
    CK(cudaGetLastError());
  }
  CKS(scan_u32_to_u64(h, flags, pos, n, &h->d_cnt->total_nnz));
  CKS(scan_u32_to_u64(h, row_cnt, out->ptr, rows, nullptr));
  CK(cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->h_cnt->error & 2u) return spam_fail(h, SPAM_EINDEX, "triplet index out of range");
  out->nnz = h->h_cnt->total_nnz;
  CKS(dev_alloc_t(h, &out->idx, out->nnz));   // owned by `out`: the caller frees it on failure
  CKS(dev_alloc(h, &out->val, out->nnz * sizeof(V)));
  if (n) {
    k_emit<V><<<grid_for(h, n), 256, 0, h->stream>>>(n, cbits, k0, p0, d_v, flags, pos, out->idx, (V*)out->val);
    count_launch(h);
    CK(cudaGetLastError());
  }
  return SPAM_OK;
}


// ---- bucket path (bucket.cuh): one partition pass, one build pass, one host sync ----
template <class K>
int bk_set_smem(spam_handle* h, K kernel, size_t bytes) {
  if (bytes > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return SPAM_OK;
}

// look-back states, ticket and bucket cursors of one call: carved from the handle's grow-only scan workspace, zeroed by
// its one memset (state[nb], ticket, then nb cursors one 32-byte sector apart)
static int bk_control(spam_handle* h, u32 nb, u64** state, u32** ticket, u32** cursor) {
  const u64 words = (u64)nb + 2 + ((u64)nb * BK_CUR_STRIDE * sizeof(u32) + 7) / 8;
  u32* unused = nullptr;
  CKS(lookback_workspace(h, words, state, &unused));  // zeroes words + 1 u64
  *ticket = reinterpret_cast<u32*>(*state + nb);
  *cursor = reinterpret_cast<u32*>(*state + nb + 2);
  return SPAM_OK;
}

// *done = false with SPAM_OK: the path does not apply (shape) or raised a flag (a bucket overflowed, a row is too
// long) — the caller takes the counting / radix path.  The result arrays are sized for n entries (nnz <= n).
template <class V>
int dok_bucket(spam_handle* h, u64 rows, u64 cols, u64 n, const u64* d_r, const u64* d_c, const V* d_v, spam_dcsr* out,
               bool* done) {
  *done = false;
  BkPlan pl;
  if (!h->dok_bucket || !bk_plan(rows, cols, n, &pl)) return SPAM_OK;
  DevGuard g(h);
  u32 *cursor = nullptr, *idx = nullptr;
  uint4* part = nullptr;
  V* val = nullptr;
  CKS(g.alloc(&part, (u64)pl.nb * BK_CAP));
  CKS(g.alloc(&idx, n));
  CKS(g.alloc(&val, n));
  u64* state = nullptr;
  u32* ticket = nullptr;
  CKS(bk_control(h, pl.nb, &state, &ticket, &cursor));
  CK(cudaMemsetAsync(h->d_cnt, 0, sizeof(Counters), h->stream));
  const size_t psmem = bk_part_smem(pl.nb);
  CKS(bk_set_smem(h, k_bk_part_dok<V>, psmem));
  CKS(bk_set_smem(h, k_bk_build<V, true>, bk_build_smem(BK_SHIFT_MAX)));
  k_bk_part_dok<V><<<(unsigned)((n + BK_PTILE - 1) / BK_PTILE), BK_PT, psmem, h->stream>>>(
      n, rows, cols, pl.shift, pl.mbits, pl.nb, d_r, d_c, d_v, cursor, part, h->d_cnt);
  k_bk_build<V, true><<<pl.nb, BK_BT, bk_build_smem(pl.shift), h->stream>>>(rows, pl.shift, pl.mbits, pl.nb, cursor, part, state,
                                                                       ticket, out->ptr, idx, val, h->d_cnt);
  count_launch(h, 2);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->h_cnt->error & 2u) return spam_fail(h, SPAM_EINDEX, "triplet index out of range");
  if (h->h_cnt->bk_flags) return SPAM_OK;
  out->nnz = h->h_cnt->total_nnz;
  out->idx = idx; out->val = val;
  g.release(idx); g.release(val);
  h->stats.fallbacks[4] = 3;
  *done = true;
  return SPAM_OK;
}

template <class V>
int dok_typed(spam_handle* h, u64 rows, u64 cols, u64 n, const u64* d_r, const u64* d_c, const V* d_v, spam_dcsr* out) {
  CKS(dev_alloc_t(h, &out->ptr, rows + 1));
  {
    bool done = false;
    CKS(dok_bucket<V>(h, rows, cols, n, d_r, d_c, d_v, out, &done));
    if (done) return SPAM_OK;
  }
  DevGuard g(h);
  u32* raw_cnt = nullptr;
  unsigned char* rank8 = nullptr;
  u64* seg_ptr = nullptr;
  CKS(g.alloc(&raw_cnt, rows));
  CKS(g.alloc(&seg_ptr, rows + 1));
  CKS(g.alloc(&rank8, n ? n : 1));
  CK(cudaMemsetAsync(h->d_cnt, 0, sizeof(Counters), h->stream));
  CK(cudaMemsetAsync(raw_cnt, 0, rows * sizeof(u32), h->stream));
  if (n) {
    k_dok_hist<<<grid_for(h, (n + 3) / 4), 256, 0, h->stream>>>(n, rows, d_r, raw_cnt, rank8, h->d_cnt);
    count_launch(h);
    CK(cudaGetLastError());
  }
  CKS(scan_u32_to_u64(h, raw_cnt, seg_ptr, rows, nullptr, &h->d_cnt->max_flop));  // also: the longest row
  CK(cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->h_cnt->error & 2u) return spam_fail(h, SPAM_EINDEX, "triplet index out of range");
  if (h->h_cnt->max_flop > SEG_MAX) {
    h->stats.fallbacks[4] = 2;
    return dok_radix<V>(h, rows, cols, n, d_r, d_c, d_v, out);
  }
  h->stats.fallbacks[4] = 1;
  uint4* ent = nullptr;
  CKS(g.alloc(&ent, n ? n : 1));
  const unsigned rgrid = (unsigned)((rows + 127) / 128);   // rows >= 1 (NonZeroUsize in the reference)
  if (n) {
    k_dok_scatter<V><<<grid_for(h, n), 256, 0, h->stream>>>(n, cols, d_r, d_c, d_v, rank8, seg_ptr, ent, h->d_cnt);
    count_launch(h);
  }
  k_dok_seg<V><<<rgrid, 128, 0, h->stream>>>(rows, seg_ptr, ent, raw_cnt);  // raw_cnt reused: survivors per row
  count_launch(h);
  CK(cudaGetLastError());
  CKS(scan_u32_to_u64(h, raw_cnt, out->ptr, rows, &h->d_cnt->total_nnz));
  CK(cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->h_cnt->error & 2u) return spam_fail(h, SPAM_EINDEX, "triplet index out of range");
  out->nnz = h->h_cnt->total_nnz;
  CKS(dev_alloc_t(h, &out->idx, out->nnz));
  CKS(dev_alloc(h, &out->val, out->nnz * sizeof(V)));
  k_dok_emit<V><<<rgrid, 128, 0, h->stream>>>(rows, seg_ptr, ent, out->ptr, out->idx, (V*)out->val);
  count_launch(h);
  CK(cudaGetLastError());
  return SPAM_OK;
}

// ---- CSR transpose ------------------------------------------------------------
// Reference: `fn transpose` spam_csr/src/lib.rs:256-264 walks (j, i) over cols x rows and moves every stored
// entry (i, j) — explicit zeros included: CsrMatrix::set_element stores a zero like any value,
// lib.rs:213-253 — to (j, i) of the new matrix, appending in increasing i.  So row j of the result lists
// the rows i of A that hold column j in increasing i, whatever the order inside A's rows: a STABLE sort of
// the entries by column.  Here: key = column, payload = entry position, the same LSD radix sort as above
// on the column bits only, then one gather.
__global__ void __launch_bounds__(256) k_transpose_keys(u64 m, u64 cols, const u64* __restrict__ ptr,
                                                        const u32* __restrict__ idx, u64* __restrict__ keys,
                                                        u32* __restrict__ pay, u32* __restrict__ erow,
                                                        u32* __restrict__ col_cnt, Counters* cnt) {
  const int lane = threadIdx.x & 31;
  const u64 row = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = row < m;
  u64 lo = 0, hi = 0;
  if (valid) { lo = ptr[row]; hi = ptr[row + 1]; }
  bool bad = false;
  if (valid && hi - lo <= 32) {
    for (u64 e = lo; e < hi; ++e) {
      const u32 c = idx[e];
      if (c < cols) atomicAdd(&col_cnt[c], 1u); else bad = true;
      keys[e] = c < cols ? c : 0; pay[e] = (u32)e; erow[e] = (u32)row;
    }
  }
  unsigned longmask = __ballot_sync(0xffffffffu, valid && hi - lo > 32);  // long rows: the whole warp helps
  while (longmask) {
    const int src = __ffs(longmask) - 1;
    longmask &= longmask - 1;
    const u64 l = __shfl_sync(0xffffffffu, lo, src), hh = __shfl_sync(0xffffffffu, hi, src);
    const u32 rr = (u32)__shfl_sync(0xffffffffu, row, src);
    for (u64 e = l + lane; e < hh; e += 32) {
      const u32 c = idx[e];
      if (c < cols) atomicAdd(&col_cnt[c], 1u); else bad = true;
      keys[e] = c < cols ? c : 0; pay[e] = (u32)e; erow[e] = rr;
    }
  }
  if (bad) atomicOr(&cnt->error, 2u);
}

template <class W>  // W = uint32_t / uint64_t: values are moved, never interpreted
__global__ void __launch_bounds__(256) k_transpose_emit(u64 n, const u32* __restrict__ pay, const u32* __restrict__ erow,
                                                        const W* __restrict__ val, u32* __restrict__ t_idx,
                                                        W* __restrict__ t_val) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const u32 e = pay[i];
    t_idx[i] = erow[e];
    t_val[i] = val[e];
  }
}

}  // namespace

static int transpose_radix(spam_handle* h, const spam_dcsr* a, spam_dcsr** out) {
  *out = nullptr;
  const u64 m = a->rows, n = a->nnz, tc = a->cols;
  if (n >= 0xFFFFFFFFull) return spam_fail(h, SPAM_EOVERFLOW, "more than 2^32-1 entries");
  h->stats.fallbacks[4] = 2;
  spam_dcsr* t = new spam_dcsr();
  t->dtype = a->dtype; t->rows = a->cols; t->cols = a->rows; t->nnz = n; t->owning = true;
  t->rows_sorted = -1; t->max_row_len = 0;  // sorted by construction; the cached stats (longest row) are taken lazily
  t->ptr = nullptr; t->idx = nullptr; t->val = nullptr;
  const size_t es = dtype_size(a->dtype);
  const u64 nblocks = (n + RS_TILE - 1) / RS_TILE;
  auto al = [](u64 bytes) { return (bytes + 255) & ~255ull; };
  const u64 sz_k = al(n * 8), sz_p = al(n * 4), sz_cc = al(tc * 4), sz_hist = al((u64)RADIX * nblocks * 4),
            sz_offs = al(((u64)RADIX * nblocks + 1) * 8);
  char* ws = nullptr;
  int st = dev_alloc(h, (void**)&ws, 2 * sz_k + 3 * sz_p + sz_cc + sz_hist + sz_offs);
  if (st == SPAM_OK) st = dev_alloc_t(h, &t->ptr, tc + 1);
  if (st == SPAM_OK) st = dev_alloc_t(h, &t->idx, n);
  if (st == SPAM_OK) st = dev_alloc(h, &t->val, n * es);
  auto fail = [&](int s) {
    dev_free(h, ws); dev_free(h, t->ptr); dev_free(h, t->idx); dev_free(h, t->val);
    delete t;
    return s;
  };
  if (st != SPAM_OK) return fail(st);
  char* cur = ws;
  u64* k0 = (u64*)cur; cur += sz_k;
  u64* k1 = (u64*)cur; cur += sz_k;
  u64* offs = (u64*)cur; cur += sz_offs;
  u32* p0 = (u32*)cur; cur += sz_p;
  u32* p1 = (u32*)cur; cur += sz_p;
  u32* erow = (u32*)cur; cur += sz_p;
  u32* col_cnt = (u32*)cur; cur += sz_cc;
  u32* hist = (u32*)cur;
  cudaError_t e = cudaMemsetAsync(h->d_cnt, 0, sizeof(Counters), h->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(col_cnt, 0, tc * sizeof(u32), h->stream);
  if (e != cudaSuccess) return fail(spam_fail(h, SPAM_ECUDA, "cudaMemsetAsync", e));
  if (m) {
    k_transpose_keys<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(m, tc, a->ptr, a->idx, k0, p0, erow, col_cnt, h->d_cnt);
    count_launch(h);
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(spam_fail(h, SPAM_ECUDA, "k_transpose_keys", e));
  }
  st = radix_sort_pairs(h, n, bits_for(tc), k0, p0, k1, p1, hist, offs);
  if (st == SPAM_OK) st = scan_u32_to_u64(h, col_cnt, t->ptr, tc, nullptr);
  if (st != SPAM_OK) return fail(st);
  if (n) {
    if (es == 4)
      k_transpose_emit<uint32_t><<<grid_for(h, n), 256, 0, h->stream>>>(n, p0, erow, (const uint32_t*)a->val, t->idx, (uint32_t*)t->val);
    else
      k_transpose_emit<uint64_t><<<grid_for(h, n), 256, 0, h->stream>>>(n, p0, erow, (const uint64_t*)a->val, t->idx, (uint64_t*)t->val);
    count_launch(h);
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(spam_fail(h, SPAM_ECUDA, "k_transpose_emit", e));
  }
  e = cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return fail(spam_fail(h, SPAM_ECUDA, "transpose sync", e));
  if (h->h_cnt->error & 2u) return fail(spam_fail(h, SPAM_EINDEX, "a column index is >= cols"));
  dev_free(h, ws);
  *out = t;
  return SPAM_OK;
}

// Transpose by the bucket path: buckets of columns, records (column_local << bits(rows) | row, entry position, value).
static int transpose_bucket(spam_handle* h, const spam_dcsr* a, spam_dcsr** out, bool* done) {
  *done = false;
  const u64 m = a->rows, n = a->nnz, tc = a->cols;
  BkPlan pl;
  if (!h->dok_bucket || !bk_plan(tc, m, n, &pl)) return SPAM_OK;
  const size_t es = dtype_size(a->dtype);
  DevGuard g(h);
  u32 *cursor = nullptr, *t_idx = nullptr;
  uint4* part = nullptr;
  u64* t_ptr = nullptr;
  void* t_val = nullptr;
  CKS(g.alloc(&part, (u64)pl.nb * BK_CAP));
  CKS(g.alloc(&t_ptr, tc + 1));
  CKS(g.alloc(&t_idx, n));
  CKS(g.alloc_bytes(&t_val, n * es));
  u64* state = nullptr;
  u32* ticket = nullptr;
  CKS(bk_control(h, pl.nb, &state, &ticket, &cursor));
  CK(cudaMemsetAsync(h->d_cnt, 0, sizeof(Counters), h->stream));
  const size_t psmem = bk_part_smem(pl.nb);
  const unsigned pgrid = (unsigned)((n + BK_PTILE - 1) / BK_PTILE);
  if (es == 4) {
    CKS(bk_set_smem(h, k_bk_part_csr<uint32_t>, psmem));
    CKS(bk_set_smem(h, k_bk_build<uint32_t, false>, bk_build_smem(BK_SHIFT_MAX)));
    k_bk_part_csr<uint32_t><<<pgrid, BK_PT, psmem, h->stream>>>(m, n, tc, pl.shift, pl.mbits, pl.nb, a->ptr, a->idx,
                                                               (const uint32_t*)a->val, cursor, part, h->d_cnt);
    k_bk_build<uint32_t, false><<<pl.nb, BK_BT, bk_build_smem(pl.shift), h->stream>>>(
        tc, pl.shift, pl.mbits, pl.nb, cursor, part, state, ticket, t_ptr, t_idx, (uint32_t*)t_val, h->d_cnt);
  } else {
    CKS(bk_set_smem(h, k_bk_part_csr<uint64_t>, psmem));
    CKS(bk_set_smem(h, k_bk_build<uint64_t, false>, bk_build_smem(BK_SHIFT_MAX)));
    k_bk_part_csr<uint64_t><<<pgrid, BK_PT, psmem, h->stream>>>(m, n, tc, pl.shift, pl.mbits, pl.nb, a->ptr, a->idx,
                                                               (const uint64_t*)a->val, cursor, part, h->d_cnt);
    k_bk_build<uint64_t, false><<<pl.nb, BK_BT, bk_build_smem(pl.shift), h->stream>>>(
        tc, pl.shift, pl.mbits, pl.nb, cursor, part, state, ticket, t_ptr, t_idx, (uint64_t*)t_val, h->d_cnt);
  }
  count_launch(h, 2);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->h_cnt->error & 2u) return spam_fail(h, SPAM_EINDEX, "a column index is >= cols");
  if (h->h_cnt->bk_flags) return SPAM_OK;
  spam_dcsr* t = new spam_dcsr();
  t->dtype = a->dtype; t->rows = a->cols; t->cols = a->rows; t->nnz = n; t->owning = true;
  t->rows_sorted = -1; t->max_row_len = 0;
  t->ptr = t_ptr; t->idx = t_idx; t->val = t_val;
  g.release(t_ptr); g.release(t_idx); g.release(t_val);
  h->stats.fallbacks[4] = 3;
  *out = t;
  *done = true;
  return SPAM_OK;
}

int transpose_dev(spam_handle* h, const spam_dcsr* a, spam_dcsr** out) {
  *out = nullptr;
  const u64 m = a->rows, n = a->nnz, tc = a->cols;
  if (n >= 0xFFFFFFFFull) return spam_fail(h, SPAM_EOVERFLOW, "more than 2^32-1 entries");
  if (m >= 0xFFFFFFFFull || tc >= 0xFFFFFFFFull) return spam_fail(h, SPAM_ECOLS, "dimension >= 2^32-1");
  CKS(ensure_matrix_stats(h, a));  // cached per matrix: the CSR invariants every path below relies on (row_ptr monotone, idx < cols)
  h->stats = spam_stats{};
  if (m == 0 || tc == 0) return transpose_radix(h, a, out);
  {
    bool done = false;
    CKS(transpose_bucket(h, a, out, &done));
    if (done) return SPAM_OK;
  }
  // Counting path: histogram by column -> scan -> scatter -> per-column order by row.  Taken when no column holds
  // more than SEG_MAX entries (one thread sorts a column); otherwise the stable radix sort above.
  const size_t es = dtype_size(a->dtype);
  DevGuard g(h);
  u32 *col_cnt = nullptr, *t_idx = nullptr;
  u64* t_ptr = nullptr;
  void* t_val = nullptr;
  unsigned char* rank8 = nullptr;
  CKS(g.alloc(&col_cnt, tc));
  CKS(g.alloc(&t_ptr, tc + 1));
  CKS(g.alloc(&rank8, n ? n : 1));
  CK(cudaMemsetAsync(h->d_cnt, 0, sizeof(Counters), h->stream));
  CK(cudaMemsetAsync(col_cnt, 0, tc * sizeof(u32), h->stream));
  if (n) {
    k_tr_hist<<<grid_for(h, (n + 3) / 4), 256, 0, h->stream>>>(n, tc, a->idx, col_cnt, rank8, h->d_cnt);
    count_launch(h);
    CK(cudaGetLastError());
  }
  CKS(scan_u32_to_u64(h, col_cnt, t_ptr, tc, nullptr, &h->d_cnt->max_flop));  // also: the longest column
  CK(cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->h_cnt->error & 2u) return spam_fail(h, SPAM_EINDEX, "a column index is >= cols");
  if (h->h_cnt->max_flop > SEG_MAX) return transpose_radix(h, a, out);
  h->stats.fallbacks[4] = 1;
  CKS(g.alloc(&t_idx, n));
  CKS(g.alloc_bytes(&t_val, n * es));
  if (n) {
    uint4* rec = nullptr;
    CKS(g.alloc(&rec, n));
    const unsigned rgrid = (unsigned)((m + 255) / 256), cgrid = (unsigned)((tc + 127) / 128);
    if (es == 4) {
      k_tr_scatter<uint32_t><<<rgrid, 256, 0, h->stream>>>(m, a->ptr, a->idx, (const uint32_t*)a->val, rank8, t_ptr, rec);
      k_tr_segsort<uint32_t><<<cgrid, 128, 0, h->stream>>>(tc, t_ptr, rec, t_idx, (uint32_t*)t_val);
    } else {
      k_tr_scatter<uint64_t><<<rgrid, 256, 0, h->stream>>>(m, a->ptr, a->idx, (const uint64_t*)a->val, rank8, t_ptr, rec);
      k_tr_segsort<uint64_t><<<cgrid, 128, 0, h->stream>>>(tc, t_ptr, rec, t_idx, (uint64_t*)t_val);
    }
    count_launch(h, 2);
    CK(cudaGetLastError());
  }
  spam_dcsr* t = new spam_dcsr();
  t->dtype = a->dtype; t->rows = a->cols; t->cols = a->rows; t->nnz = n; t->owning = true;
  t->rows_sorted = -1; t->max_row_len = 0;  // sorted by construction; the cached stats (longest row) are taken lazily
  t->ptr = t_ptr; t->idx = t_idx; t->val = t_val;
  g.release(t_ptr); g.release(t_idx); g.release(t_val);
  *out = t;
  return SPAM_OK;
}

// ---- multi-GPU DOK -> CSR, local half: stable partition of a slice of the triplet stream by the rank that owns the
// row (rows_per consecutive rows per rank), rows rebased to the owner's block ----
namespace {
__global__ void __launch_bounds__(256) k_dest_keys(u64 n, u64 rows, u64 cols, u64 rows_per, const u64* __restrict__ r,
                                                   const u64* __restrict__ c, u64* __restrict__ keys, u32* __restrict__ pay,
                                                   ull* __restrict__ counts, Counters* cnt) {
  __shared__ u32 s_cnt[32];
  if (threadIdx.x < 32) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const u64 stride = (u64)gridDim.x * blockDim.x;
  bool bad = false;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const u64 ri = r[i];
    u64 d = 0;
    if (ri < rows && c[i] < cols) { d = ri / rows_per; atomicAdd(&s_cnt[d & 31], 1u); } else bad = true;
    keys[i] = d;
    pay[i] = (u32)i;
  }
  if (bad) atomicOr(&cnt->error, 2u);
  __syncthreads();
  if (threadIdx.x < 32 && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (ull)s_cnt[threadIdx.x]);
}
template <class V>
__global__ void __launch_bounds__(256) k_dest_gather(u64 n, u64 rows_per, const u64* __restrict__ keys, const u32* __restrict__ pay,
                                                     const u64* __restrict__ r, const u64* __restrict__ c,
                                                     const V* __restrict__ v, u64* __restrict__ o_r, u64* __restrict__ o_c,
                                                     V* __restrict__ o_v) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const u32 i = pay[j];
    o_r[j] = r[i] - keys[j] * rows_per;
    o_c[j] = c[i];
    o_v[j] = v[i];
  }
}
}  // namespace

// counts_host[d] = triplets of this slice owned by rank d; o_r / o_c / o_v = the slice grouped by owner (stream order
// kept inside every group: the radix pass is stable).  world <= 32.
int dok_partition_dev(spam_handle* h, int dtype, u64 rows, u64 cols, u64 rows_per, int world, u64 n, const u64* d_r,
                      const u64* d_c, const void* d_v, u64* o_r, u64* o_c, void* o_v, u64* counts_host) {
  for (int d = 0; d < world; ++d) counts_host[d] = 0;
  if (n == 0) return SPAM_OK;
  if (n >= 0xFFFFFFFFull) return spam_fail(h, SPAM_EOVERFLOW, "more than 2^32-1 triplets in one slice");
  const u64 nblocks = (n + RS_TILE - 1) / RS_TILE;
  DevGuard g(h);
  u64 *k0 = nullptr, *k1 = nullptr, *offs = nullptr;
  u32 *p0 = nullptr, *p1 = nullptr, *hist = nullptr;
  ull* counts = nullptr;
  CKS(g.alloc(&k0, n)); CKS(g.alloc(&k1, n)); CKS(g.alloc(&p0, n)); CKS(g.alloc(&p1, n));
  CKS(g.alloc(&hist, (u64)RADIX * nblocks)); CKS(g.alloc(&offs, (u64)RADIX * nblocks + 1));
  CKS(g.alloc(&counts, 32));
  CK(cudaMemsetAsync(h->d_cnt, 0, sizeof(Counters), h->stream));
  CK(cudaMemsetAsync(counts, 0, 32 * sizeof(ull), h->stream));
  k_dest_keys<<<grid_for(h, n), 256, 0, h->stream>>>(n, rows, cols, rows_per, d_r, d_c, k0, p0, counts, h->d_cnt);
  count_launch(h);
  CK(cudaGetLastError());
  CKS(radix_sort_pairs(h, n, 8, k0, p0, k1, p1, hist, offs));  // one stable 8-bit pass: at most 32 owners
  const unsigned grid = grid_for(h, n);
  switch (dtype) {
    case SPAM_F32: case SPAM_I32:
      k_dest_gather<uint32_t><<<grid, 256, 0, h->stream>>>(n, rows_per, k0, p0, d_r, d_c, (const uint32_t*)d_v, o_r, o_c, (uint32_t*)o_v); break;
    default:
      k_dest_gather<uint64_t><<<grid, 256, 0, h->stream>>>(n, rows_per, k0, p0, d_r, d_c, (const uint64_t*)d_v, o_r, o_c, (uint64_t*)o_v); break;
  }
  count_launch(h);
  CK(cudaGetLastError());
  ull hc[32];
  CK(cudaMemcpyAsync(hc, counts, sizeof(hc), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->h_cnt->error & 2u) return spam_fail(h, SPAM_EINDEX, "triplet index out of range");
  for (int d = 0; d < world; ++d) counts_host[d] = hc[d];
  return SPAM_OK;
}

// Rows in column order: a stable sort of the entries by column, twice (A -> A^T -> A), each a counting or radix
// sort above.  Cached with the matrix: device matrices are immutable through this API.
int sorted_rows_of(spam_handle* h, const spam_dcsr* m, const spam_dcsr** view) {
  *view = m;
  CKS(ensure_matrix_stats(h, m));
  if (m->rows_sorted == 1) return SPAM_OK;
  if (!m->sorted_copy) {
    const spam_stats keep = h->stats;
    spam_dcsr* t = nullptr;
    CKS(transpose_dev(h, m, &t));
    spam_dcsr* tt = nullptr;
    const int st = transpose_dev(h, t, &tt);
    free_dcsr_tree(h, t);
    h->stats = keep;
    if (st != SPAM_OK) return st;
    tt->rows_sorted = 1;
    tt->max_row_len = m->max_row_len;
    tt->spread_sum = m->spread_sum;
    tt->invalid = 0;
    const_cast<spam_dcsr*>(m)->sorted_copy = tt;
  }
  *view = m->sorted_copy;
  return SPAM_OK;
}

int dok_to_csr_dev(spam_handle* h, int dtype, u64 rows, u64 cols, u64 n, const u64* d_r, const u64* d_c,
                   const void* d_v, spam_dcsr** out) {
  if (n >= 0xFFFFFFFFull) return spam_fail(h, SPAM_EOVERFLOW, "more than 2^32-1 triplets");
  h->stats = spam_stats{};
  spam_dcsr* m = new spam_dcsr();
  m->dtype = dtype; m->rows = rows; m->cols = cols; m->nnz = 0; m->owning = true; m->rows_sorted = -1; m->max_row_len = 0;
  m->ptr = nullptr; m->idx = nullptr; m->val = nullptr;
  int st;
  switch (dtype) {
    case SPAM_F32: st = dok_typed<float>(h, rows, cols, n, d_r, d_c, (const float*)d_v, m); break;
    case SPAM_F64: st = dok_typed<double>(h, rows, cols, n, d_r, d_c, (const double*)d_v, m); break;
    case SPAM_I32: st = dok_typed<int32_t>(h, rows, cols, n, d_r, d_c, (const int32_t*)d_v, m); break;
    case SPAM_I64: st = dok_typed<int64_t>(h, rows, cols, n, d_r, d_c, (const int64_t*)d_v, m); break;
    default: st = spam_fail(h, SPAM_EINVAL, "bad dtype");
  }
  if (st != SPAM_OK) {
    dev_free(h, m->ptr); dev_free(h, m->idx); dev_free(h, m->val);
    delete m;
    return st;
  }
  *out = m;
  return SPAM_OK;
}
