// bucket.cuh — DOK -> CSR and CSR transpose as ONE partition pass + ONE build pass (round 2).
//
// Both operations are "group entries by a major index, order each group by a minor index":
//   DOK -> CSR  (spam_csr/src/lib.rs:315-334 over spam_dok/src/lib.rs:167-176): major = row, minor = column, the last
//               write of a (row, column) key wins and a zero deletes;
//   transpose   (spam_csr/src/lib.rs:256-264): major = column of A, minor = row of A, nothing is dropped.
// The counting paths of dok.cu scatter one record per entry to its major's segment: 8 M scattered 16-byte stores over
// a 128 MB array are bound by DRAM page activations (242 us of the 531 us of a C5 build).  Here the majors are cut
// into buckets of 2^shift consecutive majors, sized so that a bucket's entries fit in shared memory:
//   k_bk_part_*   one pass over the input: a block ranks its tile's entries per bucket in shared memory, takes a slice
//                 of every bucket with ONE global atomic per (block, bucket), puts the tile in bucket order in shared
//                 memory and stores 16-byte records (major_local << mbits | minor, stream position, value), consecutive
//                 threads storing consecutive records of a bucket (whole sectors).  A bucket's region is filled front
//                 to back by all blocks, so the partially written lines — one per bucket — live in L2 and DRAM sees
//                 whole lines.  Buckets have a fixed capacity (BK_CAP records): no counting pass, no scan.
//   k_bk_build    one block per bucket (in major order, handed out by a ticket): counting sort by major inside shared
//                 memory, then every ENTRY decides for itself — it walks its own segment (neighbouring threads read
//                 the same words: broadcasts, no bank conflicts, no idle lanes) to find out whether a later write of
//                 its key exists, and ranks itself by minor among the survivors — a block scan of the survivors per
//                 major, a decoupled look-back over the buckets for the position in the result, and the result
//                 (offsets, indices, values) is written once, nearly coalesced.
// Bytes: the stream is read once (24 B per triplet), the records written and read once (16 B, the second read of the
// build pass hits L2), the result written once: 24 + 32 + 12 per triplet against 37 B algorithmic.
// A bucket that overflows (clustered majors) or a segment longer than BK_SEG_MAX raises a flag; the caller then
// takes the counting / radix paths of dok.cu, which stay the general case.
#pragma once
#include "common.cuh"

namespace {

constexpr u32 BK_CAP = 5040;      // records per bucket: 10 bytes each in k_bk_build; 3 blocks of 512-major buckets fit the 164 KB carve-out, 92 KB stay L1
constexpr u32 BK_NB_MAX = 8192;   // buckets: the partition kernel keeps two u32 per bucket in shared memory
constexpr int BK_SHIFT_MAX = 11;  // at most 2048 majors per bucket (bk_block_scan: 4 per thread)
constexpr int BK_PT = 512, BK_PITEMS = 16, BK_PTILE = BK_PT * BK_PITEMS;  // partition tile: 8192 entries
constexpr int BK_BT = 512;        // threads of the build kernel
constexpr u32 BK_SEG_MAX = 512;   // longest segment the all-pairs loops are allowed to take (quadratic per segment)
constexpr u32 BK_DROPPED = 0xFFFFu;

struct BkPlan { int shift; u32 nb; int mbits; };

// Largest bucket width (fewest buckets) whose average fill leaves 15 % of BK_CAP for the spread of the bucket sizes.
static bool bk_plan(u64 majors, u64 minors, u64 n, BkPlan* p) {
  if (majors == 0 || n == 0) return false;
  int mbits = 0;
  while (mbits < 63 && (1ull << mbits) < minors) ++mbits;
  if (mbits > 32) return false;
  const int smax = 32 - mbits < BK_SHIFT_MAX ? 32 - mbits : BK_SHIFT_MAX;
  for (int s = smax; s >= 0; --s) {
    const u64 nb = (majors + (1ull << s) - 1) >> s;
    if (nb > BK_NB_MAX) return false;
    if (n * 100 <= (u64)BK_CAP * 85 * nb) { p->shift = s; p->nb = (u32)nb; p->mbits = mbits; return true; }
  }
  return false;
}

template <class V>
__device__ __forceinline__ uint4 bk_pack(u32 key, u32 pos, V v) {
  unsigned long long bits = 0;
  memcpy(&bits, &v, sizeof(V));
  return make_uint4(key, pos, (u32)bits, (u32)(bits >> 32));
}
template <class V>
__device__ __forceinline__ V bk_val(const uint4& e) {
  const unsigned long long bits = (unsigned long long)e.z | ((unsigned long long)e.w << 32);
  V v;
  memcpy(&v, &bits, sizeof(V));
  return v;
}

__device__ __forceinline__ void bk_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// shared memory of the partition kernels: counts / starts and slice offsets per bucket, the staged tile (8 bytes per entry;
// the transpose's row marks — BK_PTILE + BK_PTILE / 32 words — use the same space before)
static size_t bk_part_smem(u32 nb) { return ((size_t)2 * (nb + (nb & 1)) + 2 * BK_PTILE) * sizeof(u32); }

constexpr int BK_CUR_STRIDE = 8;  // one bucket cursor per 32-byte sector: a warp's 32 atomics spread over 8 lines

// Second half of both partition kernels.  br[it] = bucket << 16 | rank inside (block, bucket), or ~0 for an entry that
// is not placed; entry `it` of thread `tid` is element t0 + it * BK_PT + tid of the input.
// The tile's (key, place, bucket) triples are first put in BUCKET ORDER in shared memory (s_stage, 8 bytes each), so that
// consecutive threads store consecutive records of a bucket: the stores of a run leave the SM as whole 32-byte sectors.
// (Scattered 16-byte stores straight from the registers — the first version — are half-sector writes the L2 has to
// merge or fill: 168 us against 85 us for the same bytes stored contiguously and 73 us with no stores at all,
// profiles/r02_bucket_dok_transpose.txt.)  One global atomic per (block, bucket) reserves the block's slice of the
// bucket; four are in flight per thread.  s_hist: counts -> starts inside the tile; s_base: slice start - tile start.
template <class V>
__device__ __forceinline__ void bk_part_tail(u32 tid, u32 nb, u64 t0, const u32 (&key)[BK_PITEMS], const u32 (&br)[BK_PITEMS],
                                             const V* __restrict__ vals, u32* s_hist, u32* s_base, uint2* s_stage,
                                             u32* __restrict__ cursor, uint4* __restrict__ part, Counters* cnt) {
  __shared__ u32 s_tw[BK_PT / 32 + 1];
  const u32 lane = tid & 31, wid = tid >> 5;
  __syncthreads();
  // exclusive scan of the counts: thread tid owns buckets [tid * C, tid * C + C)
  const u32 C = (nb + BK_PT - 1) / BK_PT;
  u32 sum = 0;
  for (u32 k = 0; k < C; ++k) { const u32 d = tid * C + k; if (d < nb) sum += s_hist[d]; }
  u32 incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u32 y = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= (u32)d) incl += y;
  }
  if (lane == 31) s_tw[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    const u32 w = lane < BK_PT / 32 ? s_tw[lane] : 0u;
    u32 wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const u32 y = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= (u32)d) wi += y;
    }
    if (lane < BK_PT / 32) s_tw[lane] = wi - w;
    if (lane == BK_PT / 32 - 1) s_tw[BK_PT / 32] = wi;
  }
  __syncthreads();
  u32 run = s_tw[wid] + incl - sum;
  const u32 placed = s_tw[BK_PT / 32];
  for (u32 k0 = 0; k0 < C; k0 += 4) {
    u32 hc[4], bs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { const u32 d = tid * C + k0 + k; hc[k] = (k0 + k < C && d < nb) ? s_hist[d] : 0u; }
#pragma unroll
    for (int k = 0; k < 4; ++k) bs[k] = hc[k] ? atomicAdd(&cursor[(u64)(tid * C + k0 + k) * BK_CUR_STRIDE], hc[k]) : 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const u32 d = tid * C + k0 + k;
      if (k0 + k < C && d < nb) { s_hist[d] = run; s_base[d] = bs[k] - run; run += hc[k]; }
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < BK_PITEMS; ++it) {
    if (br[it] != 0xFFFFFFFFu) {
      const u32 b = br[it] >> 16;
      s_stage[s_hist[b] + (br[it] & 0xFFFFu)] = make_uint2(key[it], (b << 16) | (u32)(it * BK_PT + tid));
    }
  }
  __syncthreads();
  bool over = false;
  for (u32 q0 = tid; q0 < placed; q0 += 4 * BK_PT) {
    uint2 e[4];
    V val[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const u32 q = q0 + k * BK_PT;
      e[k] = q < placed ? s_stage[q] : make_uint2(0u, 0u);
      if (q < placed) val[k] = vals[t0 + (e[k].y & 0xFFFFu)];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const u32 q = q0 + k * BK_PT;
      if (q < placed) {
        const u32 b = e[k].y >> 16;
        const u32 p = s_base[b] + q;
        if (p < BK_CAP) part[(u64)b * BK_CAP + p] = bk_pack(e[k].x, (u32)(t0 + (e[k].y & 0xFFFFu)), val[k]);
        else over = true;
      }
    }
  }
  if (over) atomicOr(&cnt->bk_flags, 1u);
}

// Triplet stream -> bucket records.  key = (row & (2^shift - 1)) << mbits | col.
template <class V>
__global__ void __launch_bounds__(BK_PT, 2) k_bk_part_dok(u64 n, u64 rows, u64 cols, int shift, int mbits, u32 nb,
                                                          const u64* __restrict__ r, const u64* __restrict__ c,
                                                          const V* __restrict__ v, u32* __restrict__ cursor,
                                                          uint4* __restrict__ part, Counters* cnt) {
  extern __shared__ u32 s_bk[];
  u32* s_hist = s_bk;
  u32* s_base = s_bk + nb;
  uint2* s_stage = reinterpret_cast<uint2*>(s_bk + 2 * (nb + (nb & 1)));  // [BK_PTILE], 8-byte aligned
  const u32 tid = threadIdx.x;
  for (u32 d = tid; d < nb; d += BK_PT) s_hist[d] = 0;
  __syncthreads();
  const u64 t0 = (u64)blockIdx.x * BK_PTILE;
  const u64 rmask = (1ull << shift) - 1;
  {  // the values are read in the second half, after two barriers: have them in L2 by then (one prefetch per line)
    const u64 i = t0 + (u64)tid * (128 / sizeof(V));
    if (i < n && i < t0 + BK_PTILE) bk_prefetch_l2(v + i);
  }
  u32 key[BK_PITEMS], br[BK_PITEMS];
  bool bad = false;
#pragma unroll
  for (int it = 0; it < BK_PITEMS; ++it) {
    const u64 i = t0 + (u64)it * BK_PT + tid;
    br[it] = 0xFFFFFFFFu; key[it] = 0;
    if (i < n) {
      const u64 ri = r[i], ci = c[i];
      if (ri < rows && ci < cols) {  // else IndexError (spam_dok lib.rs:168-170): the whole build fails
        const u32 b = (u32)(ri >> shift);
        key[it] = (u32)(((ri & rmask) << mbits) | ci);
        br[it] = (b << 16) | atomicAdd(&s_hist[b], 1u);
      } else {
        bad = true;
      }
    }
  }
  if (bad) atomicOr(&cnt->error, 2u);
  bk_part_tail<V>(tid, nb, t0, key, br, v, s_hist, s_base, s_stage, cursor, part, cnt);
}

// CSR entries -> bucket records for the transpose.  key = (col & (2^shift - 1)) << mbits | row.  The row of every entry
// of the tile: each row that starts inside the tile marks its first entry with (row - r0) in shared memory (empty rows
// share a start: the largest wins), an inclusive max-scan over the tile's 8192 slots carries the marks forward.
__device__ __forceinline__ u32 bk_warp_max_scan(u32 x, u32 lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u32 y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= (u32)d) x = max(x, y);
  }
  return x;
}

template <class W>
__global__ void __launch_bounds__(BK_PT, 2) k_bk_part_csr(u64 m, u64 nnz, u64 tcols, int shift, int mbits, u32 nb,
                                                          const u64* __restrict__ ptr, const u32* __restrict__ idx,
                                                          const W* __restrict__ val, u32* __restrict__ cursor,
                                                          uint4* __restrict__ part, Counters* cnt) {
  extern __shared__ u32 s_bk[];
  u32* s_hist = s_bk;
  u32* s_base = s_bk + nb;
  uint2* s_stage = reinterpret_cast<uint2*>(s_bk + 2 * (nb + (nb & 1)));  // [BK_PTILE], 8-byte aligned
  u32* s_mark = reinterpret_cast<u32*>(s_stage);  // [BK_PTILE + BK_PTILE / 32] while the rows are looked up (the stage is filled later); slot x at x + x / 32: both access patterns below are conflict-free
  auto MK = [](u32 x) { return x + (x >> 5); };
  __shared__ u64 s_rows[2];
  __shared__ u32 s_wmax[BK_PT / 32];
  const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (u32 d = tid; d < nb; d += BK_PT) s_hist[d] = 0;
  for (u32 d = tid; d < (u32)(BK_PTILE + BK_PTILE / 32); d += BK_PT) s_mark[d] = 0;
  const u64 t0 = (u64)blockIdx.x * BK_PTILE;
  const u64 t1 = (t0 + BK_PTILE < nnz ? t0 + BK_PTILE : nnz) - 1;  // last entry of the tile
  if (wid < 2) {  // last row whose start is <= the entry (rows are [ptr[r], ptr[r+1]); empty rows share a start):
    // a 32-ary search by one warp — 4 round trips to row_ptr for a million rows where a binary search by one thread
    // took 20, with the whole block waiting behind it
    const u64 e = wid == 0 ? t0 : t1;
    u64 lo = 0, len = m;  // candidates [lo, lo + len); ptr[lo] <= e throughout (ptr[0] = 0)
    while (len > 1) {
      const u64 step = (len + 31) / 32;
      const u64 probe = lo + (u64)lane * step;
      const bool ok = probe < lo + len && ptr[probe] <= e;  // true for a prefix of the lanes, lane 0 included
      const unsigned mk = __ballot_sync(0xffffffffu, ok);
      const u64 adv = (u64)(31 - __clz(mk)) * step;
      len = (len - adv) < step ? (len - adv) : step;
      lo += adv;
    }
    if (lane == 0) s_rows[wid] = lo;
  }
  __syncthreads();
  const u64 r0 = s_rows[0], r1 = s_rows[1];
  for (u64 rr = r0 + 1 + tid; rr <= r1; rr += BK_PT) atomicMax(&s_mark[MK((u32)(ptr[rr] - t0))], (u32)(rr - r0));  // t0 < ptr[rr] <= t1
  __syncthreads();
  {  // inclusive max-scan of s_mark: thread tid owns slots [16 tid, 16 tid + 16)
    u32 mx = 0;
    u32 loc[BK_PITEMS];
#pragma unroll
    for (int k = 0; k < BK_PITEMS; ++k) { mx = max(mx, s_mark[MK(tid * BK_PITEMS + k)]); loc[k] = mx; }
    const u32 incl = bk_warp_max_scan(mx, lane);
    if (lane == 31) s_wmax[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const u32 w = lane < BK_PT / 32 ? s_wmax[lane] : 0u;
      const u32 wi = bk_warp_max_scan(w, lane);
      if (lane < BK_PT / 32) s_wmax[lane] = wi;
    }
    __syncthreads();
    u32 before = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) before = 0;
    if (wid) before = max(before, s_wmax[wid - 1]);
#pragma unroll
    for (int k = 0; k < BK_PITEMS; ++k) s_mark[MK(tid * BK_PITEMS + k)] = max(loc[k], before);
  }
  __syncthreads();
  const u64 cmask = (1ull << shift) - 1;
  u32 key[BK_PITEMS], br[BK_PITEMS];
  bool bad = false;
#pragma unroll
  for (int it = 0; it < BK_PITEMS; ++it) {
    const u64 i = t0 + (u64)it * BK_PT + tid;
    br[it] = 0xFFFFFFFFu; key[it] = 0;
    if (i < nnz) {
      const u32 ci = idx[i];
      if (ci < tcols) {
        const u32 b = ci >> shift;
        key[it] = (u32)((((u64)ci & cmask) << mbits) | (r0 + s_mark[MK(it * BK_PT + tid)]));
        br[it] = (b << 16) | atomicAdd(&s_hist[b], 1u);
      } else {
        bad = true;
      }
    }
  }
  if (bad) atomicOr(&cnt->error, 2u);
  bk_part_tail<W>(tid, nb, t0, key, br, val, s_hist, s_base, s_stage, cursor, part, cnt);
}

// exclusive scan of s_in[0 .. r) (r <= 4 * BK_BT) into s_out[0 .. r], s_out[r] = total; returns the total to every
// thread.  s_in and s_out may be the same array.  Ends with a barrier.
__device__ __forceinline__ u32 bk_block_scan(const u32* s_in, u32* s_out, u32 r, u32* s_warp) {
  const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  u32 a[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) a[k] = 4 * tid + k < r ? s_in[4 * tid + k] : 0u;
  const u32 mine = a[0] + a[1] + a[2] + a[3];
  u32 incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u32 y = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= (u32)d) incl += y;
  }
  __syncthreads();  // every thread has read its inputs (in-place use) and s_warp is free
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    const u32 w = lane < BK_BT / 32 ? s_warp[lane] : 0u;
    u32 wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const u32 y = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= (u32)d) wi += y;
    }
    if (lane < BK_BT / 32) s_warp[lane] = wi - w;
    if (lane == BK_BT / 32 - 1) s_warp[BK_BT / 32] = wi;
  }
  __syncthreads();
  u32 excl = s_warp[wid] + incl - mine;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (4 * tid + k < r) s_out[4 * tid + k] = excl;
    excl += a[k];
  }
  const u32 total = s_warp[BK_BT / 32];
  if (tid == 0) s_out[r] = total;
  __syncthreads();
  return total;
}

// shared memory of k_bk_build: 10 bytes per record + two counters per major of the bucket (56 KB for 512 majors, 66 KB for
// 2048: three blocks per SM either way, and what is not claimed stays L1 for the second pass over the records)
static size_t bk_build_smem(int shift) {
  return (size_t)BK_CAP * (4 + 2 + 2 + 2) + (size_t)(2 * ((1u << shift) + 2)) * 4 + 64 * 4;
}

// One block per bucket.  out_ptr has majors + 1 entries; out_idx / out_val hold the result (sized for every entry).
//
// Shared memory holds, per entry in major order, only the minor index, the record's place in the bucket (+ a flag: its
// value is zero), its major and its rank; stream positions and values stay in the bucket's records (L2): positions are
// looked at only when two entries of a segment share their minor, values once, when the result is written.
// One walk over its segment tells an entry how many minors are smaller (its rank) and how many are equal.
// DEDUPE (DOK): an equal minor means a rewritten key — the later stream position wins; that rank is final unless the
// segment lost an entry (a rewritten key, a zero): only those segments — about one in ten for C5's 1 % rewrites — are
// walked a second time.
template <class V, bool DEDUPE>
__global__ void __launch_bounds__(BK_BT, 3) k_bk_build(u64 majors, int shift, int mbits, u32 nb, const u32* __restrict__ cursor,
                                                       const uint4* __restrict__ part, volatile u64* state, u32* ticket,
                                                       u64* __restrict__ out_ptr, u32* __restrict__ out_idx,
                                                       V* __restrict__ out_val, Counters* cnt) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  u32* s_minor = reinterpret_cast<u32*>(s_raw);                                // [BK_CAP] minor index, ~0 once dropped
  unsigned short* s_j = reinterpret_cast<unsigned short*>(s_minor + BK_CAP);   // [BK_CAP] record of the entry | 0x8000: zero value
  unsigned short* s_seg = s_j + BK_CAP;                                        // [BK_CAP] major_local of a sorted entry
  unsigned short* s_rk = s_seg + BK_CAP;                                       // [BK_CAP] rank among the survivors / BK_DROPPED
  const u32 R = 1u << shift;
  u32* s_cnt = reinterpret_cast<u32*>(s_rk + BK_CAP);                          // [R + 2]
  u32* s_off = s_cnt + R + 2;                                                  // [R + 2]
  u32* s_warp = s_off + R + 2;                                                 // [64]
  __shared__ u32 s_b, s_ndrop;
  __shared__ u64 s_basepos;
  constexpr u64 F_AGG = 1ull << 62, F_PFX = 2ull << 62, VMASK = (1ull << 62) - 1;
  constexpr u32 JMASK = 0x7FFFu;
  constexpr int U = 4;  // global loads in flight per thread in the two passes over the records
  const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) { s_b = atomicAdd(ticket, 1u); s_ndrop = 0; }
  for (u32 t = tid; t <= R; t += BK_BT) s_cnt[t] = 0;
  __syncthreads();
  const u32 b = s_b;
  u32 n = cursor[(u64)b * BK_CUR_STRIDE];
  if (n > BK_CAP) n = BK_CAP;  // overflow: the partition kernel raised the flag, the result is discarded
  if (!DEDUPE && tid == 0) state[b] = (b == 0 ? F_PFX : F_AGG) | (u64)n;  // nothing is dropped: the total is known now
  const uint4* rec = part + (u64)b * BK_CAP;
  const u32 mmask = mbits >= 32 ? 0xFFFFFFFFu : ((1u << mbits) - 1u);
  // 1. entries per major
  for (u32 j0 = tid; j0 < n; j0 += U * BK_BT) {
    u32 key[U];
#pragma unroll
    for (int k = 0; k < U; ++k) { const u32 j = j0 + k * BK_BT; key[k] = j < n ? rec[j].x : 0u; }
#pragma unroll
    for (int k = 0; k < U; ++k)
      if (j0 + k * BK_BT < n) atomicAdd(&s_cnt[mbits >= 32 ? 0u : (key[k] >> mbits)], 1u);
  }
  __syncthreads();
  u32 longest = 0;
  for (u32 t = tid; t < R; t += BK_BT) longest = max(longest, s_cnt[t]);
  const bool too_long = __syncthreads_or(longest > BK_SEG_MAX);
  if (too_long && tid == 0) atomicOr(&cnt->bk_flags, 2u);
  bk_block_scan(s_cnt, s_off, R, s_warp);
  for (u32 t = tid; t < R; t += BK_BT) s_cnt[t] = s_off[t];  // cursors of the placement pass
  __syncthreads();
  // 2. place (the records come from L2 this time)
  for (u32 j0 = tid; j0 < n; j0 += U * BK_BT) {
    uint4 e[U];
#pragma unroll
    for (int k = 0; k < U; ++k) { const u32 j = j0 + k * BK_BT; e[k] = j < n ? rec[j] : make_uint4(0u, 0u, 0u, 0u); }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const u32 j = j0 + k * BK_BT;
      if (j < n) {
        const u32 seg = mbits >= 32 ? 0u : (e[k].x >> mbits);
        const u32 p = atomicAdd(&s_cnt[seg], 1u);
        // num_traits::Zero::is_zero: -0.0 is zero, NaN is not (spam_dok lib.rs:167-176)
        const bool zero = DEDUPE && bk_val<V>(e[k]) == (V)0;
        s_minor[p] = e[k].x & mmask;
        s_j[p] = (unsigned short)(j | (zero ? 0x8000u : 0u));
        s_seg[p] = (unsigned short)seg;
      }
    }
  }
  __syncthreads();
  if (DEDUPE) {
    for (u32 t = tid; t < R; t += BK_BT) s_cnt[t] = 0;  // now: entries a segment loses
    __syncthreads();
  }
  // 3. one walk over the segment: rank by minor; equal minors are settled by the stream position (from the records)
  u32 ndrop = 0;
  for (u32 p = tid; p < n; p += BK_BT) {
    const u32 seg = s_seg[p];
    const u32 lo = s_off[seg], hi = s_off[seg + 1];
    const u32 mh = s_minor[p];
    u32 lt = 0, eq = 0;
    if (!too_long)
      for (u32 q = lo; q < hi; ++q) {
        const u32 o = s_minor[q];
        lt += o < mh ? 1u : 0u;
        eq += o == mh ? 1u : 0u;
      }
    bool later = false;
    if (eq > 1) {  // a rewritten key (DOK) / a repeated column in a row of an invalid matrix (transpose)
      const u32 mypos = rec[s_j[p] & JMASK].y;
      for (u32 q = lo; q < hi; ++q) {
        if (q != p && s_minor[q] == mh) {
          const u32 opos = rec[s_j[q] & JMASK].y;
          if (DEDUPE) later = later || opos > mypos;  // last write wins
          else lt += opos < mypos ? 1u : 0u;          // stable
        }
      }
    }
    if (DEDUPE) {
      const bool drop = later || too_long || (s_j[p] & 0x8000u);
      s_rk[p] = drop ? (unsigned short)BK_DROPPED : (unsigned short)lt;
      if (drop) { atomicAdd(&s_cnt[seg], 1u); ++ndrop; }
    } else {
      s_rk[p] = (unsigned short)lt;
    }
  }
  u32 total = n;
  if (DEDUPE) {
    if (ndrop) atomicAdd(&s_ndrop, ndrop);
    __syncthreads();
    total = n - s_ndrop;
    if (tid == 0) state[b] = (b == 0 ? F_PFX : F_AGG) | (u64)total;  // published early: the look-back below rarely waits
    for (u32 p = tid; p < n; p += BK_BT)
      if (s_rk[p] == BK_DROPPED) s_minor[p] = 0xFFFFFFFFu;  // never smaller than a survivor
    __syncthreads();
    // 4. segments that lost an entry: rank again among the survivors (their minors are distinct)
    for (u32 p = tid; p < n; p += BK_BT) {
      const u32 seg = s_seg[p];
      if (s_rk[p] == BK_DROPPED || s_cnt[seg] == 0) continue;
      const u32 lo = s_off[seg], hi = s_off[seg + 1];
      const u32 mh = s_minor[p];
      u32 rk = 0;
      for (u32 q = lo; q < hi; ++q) rk += s_minor[q] < mh ? 1u : 0u;
      s_rk[p] = (unsigned short)rk;
    }
    __syncthreads();
    for (u32 t = tid; t < R; t += BK_BT) s_cnt[t] = s_off[t + 1] - s_off[t] - s_cnt[t];  // survivors per major
    __syncthreads();
    bk_block_scan(s_cnt, s_off, R, s_warp);  // s_off: survivors before each major of the bucket
  } else {
    __syncthreads();
  }
  // 5. position of the bucket in the result: decoupled look-back over the buckets (ticket order)
  if (wid == 0) {
    u64 excl = 0;
    if (b > 0) {
      long long pidx = (long long)b - 1;
      for (;;) {
        const long long i = pidx - lane;
        u64 st;
        if (i >= 0) {
          do { st = state[i]; } while ((st >> 62) == 0);
        } else {
          st = F_PFX;
        }
        const unsigned pm = __ballot_sync(0xffffffffu, (st >> 62) == 2);
        const int first = pm ? (__ffs(pm) - 1) : 32;
        u64 cc = (lane <= first) ? (st & VMASK) : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cc += __shfl_xor_sync(0xffffffffu, cc, d);
        excl += cc;
        if (pm) break;
        pidx -= 32;
      }
      if (lane == 0) state[b] = F_PFX | (excl + total);
    }
    if (lane == 0) s_basepos = excl;
  }
  __syncthreads();
  const u64 base = s_basepos;
  // 6. the result
  const u64 major0 = (u64)b << shift;
  for (u32 t = tid; t < R; t += BK_BT)
    if (major0 + t < majors) out_ptr[major0 + t] = base + s_off[t];
  if (b == nb - 1 && tid == 0) {
    out_ptr[majors] = base + total;
    cnt->total_nnz = base + total;
  }
  if (too_long) return;  // flagged: the caller discards this result
  for (u32 p = tid; p < n; p += BK_BT) {
    const u32 rk = s_rk[p];
    if (DEDUPE && rk == BK_DROPPED) continue;
    const u64 o = base + s_off[s_seg[p]] + rk;
    const uint2 vb = reinterpret_cast<const uint2*>(rec + (s_j[p] & JMASK))[1];  // the record's value (L2)
    unsigned long long bits = (unsigned long long)vb.x | ((unsigned long long)vb.y << 32);
    V v;
    memcpy(&v, &bits, sizeof(V));
    out_idx[o] = s_minor[p];
    out_val[o] = v;
  }
}

}  // namespace
