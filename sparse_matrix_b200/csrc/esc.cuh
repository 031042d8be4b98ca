// esc.cuh — the bucket-sort numeric bins ("expand, sort, compress") for rows of C that do not compress.
//
// The hash bins (rowhash.cuh) pay one probe + atomicCAS + atomicAdd(value) per intermediate product and two more
// passes over a table of 2-4 slots per output entry in the drain.  That is the right tool when many products fold
// into one entry (27-point stencil: 729 products, 125 columns).  On power-law matrices almost nothing folds (R-MAT
// scale 22: 2.54e9 products, 2.53e9 entries): the table is pure insertion, and `atomicAdd(double)` on shared
// memory is a compare-and-swap loop (profiles/r01_rmat22_v3_balance.txt: 20-35 ps per product in the team
// kernels, 85-150 ps in the global-table kernel).  For rows with 2 * nnz >= products this file does what the
// reference's B2 = true branch does after its hash map (collect, sort by column: mul_hash.rs:164-175), without the
// map:
//   A  expand the row's products (column, a_ik * b_kj) into shared memory in the reference's product order;
//   B  count them into NB order-preserving buckets over the row's own column range (native u32 shared atomics),
//      scan; buckets that came out crowded (power-law columns: R-MAT's hub columns put 20-50 products of a row
//      into one of 8192 linear buckets) are split again, linearly over their own key range, into as many
//      sub-buckets as they hold products — a second count + scan over one counter per staged product;
//   C  every thread holds its 8 products in registers and scatters them to their (sub-)buckets (the staging
//      arrays are permuted in place);
//   D  one thread per bucket (1-2 entries on average) sorts it and folds equal columns (first product stored,
//      the others added: mul_hash.rs:154-161) — the only place where products meet;
//   E  scan of the per-bucket distinct counts = position of every bucket in the output row;
//   F  C is written once, consecutive threads writing consecutive entries.
// No hash table, no atomicCAS, no floating-point atomics.  Rows up to 8192 products take one block
// (k_num_esc, 128-1024 threads); longer rows (k_num_esc_heavy) are processed in column ranges of at most 8192
// products each, chosen from a coarse histogram of the row, by persistent 1024-thread blocks.  A row whose
// columns crowd into one bucket (more than ESC_BUCKET_MAX entries) is handed to the global-table hash kernel
// through a fallback list (spam_stats.fallbacks[3]).
#pragma once
#include "common.cuh"
#include "rowhash.cuh"

namespace {

constexpr int ESC_ITEMS = 8;        // staged products per thread; also counters per thread in the scans
constexpr u32 ESC_BUCKET_MAX = 48;  // longest bucket one thread sorts; beyond: fallback list

// In-place exclusive scan of a shared u32 array of 8 * TT counters (thread rt owns arr[8 rt .. 8 rt + 8): two
// 128-bit accesses).  Returns the total; *s_mx receives the largest counter.  All TT threads must call it.
template <int TT>
__device__ __forceinline__ u32 esc_scan8(u32* arr, int rt, u32* s_warp, u32* s_mx) {
  const int lane = rt & 31, w = rt >> 5;
  if (rt == 0) *s_mx = 0;
  uint4* p = reinterpret_cast<uint4*>(arr) + 2 * rt;
  const uint4 a = p[0], b = p[1];
  const u32 sum = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
  u32 mx = max(max(max(a.x, a.y), max(a.z, a.w)), max(max(b.x, b.y), max(b.z, b.w)));
  const u32 x = warp_incl_scan_u32(sum, lane);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, d));
  if (lane == 31) s_warp[w] = x;
  __syncthreads();
  if (lane == 0 && mx) atomicMax(s_mx, mx);
  u32 woff = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < TT / 32; ++i) {
    const u32 s = s_warp[i];
    if (i < w) woff += s;
    tot += s;
  }
  uint4 oa, ob;
  oa.x = woff + x - sum; oa.y = oa.x + a.x; oa.z = oa.y + a.y; oa.w = oa.z + a.z;
  ob.x = oa.w + a.w; ob.y = ob.x + b.x; ob.z = ob.y + b.y; ob.w = ob.z + b.z;
  p[0] = oa; p[1] = ob;
  __syncthreads();
  return tot;
}

constexpr u32 ESC_SPLIT_MIN = 4;  // level-1 buckets longer than this are split again

// level-2 bucket of key k: level-1 bucket b = (k - kmin) >> bshift holds n_b products at [start_b, start_b + n_b) of
// the sorted order; when n_b > ESC_SPLIT_MIN its key range is cut linearly into n_b sub-buckets.  The result,
// start_b + sub, is monotone in k and unique per (bucket, sub-bucket).
template <int NB>
__device__ __forceinline__ u32 esc_bucket2(const u32* cnt, u32 n, u32 k, u32 kmin, int bshift) {
  const u32 d = k - kmin;
  const u32 b = d >> bshift;
  const u32 start = cnt[b];
  const u32 nb = (b + 1 < (u32)NB ? cnt[b + 1] : n) - start;
  u32 sub = 0;
  if (nb > ESC_SPLIT_MIN) sub = (u32)(((u64)(d - (b << bshift)) * nb) >> bshift);
  return start + sub;
}

// Steps B-F on the products the block's threads hold in registers: thread rt has k[i], v[i] for the bits i set in
// vmask; n = number of products in the block.  `cnt` and `sub` must be zero on entry; skey / sval are scratch
// (free to overwrite once every thread has its products: the caller's barrier, or the first one in here).
// Writes the distinct entries, sorted by column, to c_col/c_val[cbase ...) — at most `room` of them — and
// returns their number through `uniq`.  Returns false (block-uniform, nothing useful written) when some
// sub-bucket is still longer than ESC_BUCKET_MAX.
template <class V, int TT>
__device__ __forceinline__ bool esc_core(u32 (&k)[ESC_ITEMS], V (&v)[ESC_ITEMS], u32 vmask, u32* skey, V* sval, u32* cnt,
                                         u32* sub, u32* s_warp, u32* s_mx, u32 n, u32 kmin, int bshift, int rt,
                                         u32* __restrict__ c_col, V* __restrict__ c_val, u64 cbase, u32 room,
                                         u32& uniq) {
  constexpr int NB = TT * ESC_ITEMS;
#pragma unroll
  for (int i = 0; i < ESC_ITEMS; ++i)
    if (vmask >> i & 1) atomicAdd(&cnt[(k[i] - kmin) >> bshift], 1u);
  __syncthreads();
  esc_scan8<TT>(cnt, rt, s_warp, s_mx);  // cnt[b] = first position of level-1 bucket b
  if (*s_mx > ESC_SPLIT_MIN) {  // some bucket is crowded: second level
#pragma unroll
    for (int i = 0; i < ESC_ITEMS; ++i)
      if (vmask >> i & 1) atomicAdd(&sub[esc_bucket2<NB>(cnt, n, k[i], kmin, bshift)], 1u);
  } else {
#pragma unroll
    for (int i = 0; i < ESC_ITEMS; ++i)
      if (vmask >> i & 1) atomicAdd(&sub[cnt[(k[i] - kmin) >> bshift]], 1u);
  }
  __syncthreads();
  esc_scan8<TT>(sub, rt, s_warp, s_mx);  // sub[j] = first position of level-2 bucket j
  if (*s_mx > ESC_BUCKET_MAX) { uniq = 0; return false; }
#pragma unroll
  for (int i = 0; i < ESC_ITEMS; ++i) {
    if (vmask >> i & 1) {
      const u32 pos = atomicAdd(&sub[esc_bucket2<NB>(cnt, n, k[i], kmin, bshift)], 1u);  // afterwards sub[j] = end of j
      skey[pos] = k[i]; sval[pos] = v[i];
    }
  }
  __syncthreads();
  u32 u[ESC_ITEMS];
#pragma unroll
  for (int j = 0; j < ESC_ITEMS; ++j) {
    const u32 b = rt + j * TT;
    const u32 lo = b ? sub[b - 1] : 0u, hi = sub[b];
    u32 un = hi - lo;
    if (un >= 2) {
      for (u32 i = lo + 1; i < hi; ++i) {  // insertion sort, stable
        const u32 kk = skey[i];
        const V vv = sval[i];
        u32 q = i;
        while (q > lo && skey[q - 1] > kk) { skey[q] = skey[q - 1]; sval[q] = sval[q - 1]; --q; }
        skey[q] = kk; sval[q] = vv;
      }
      u32 w = lo;
      for (u32 i = lo + 1; i < hi; ++i) {  // fold equal columns: first product stored, the others added
        if (skey[i] == skey[w]) sval[w] = Num<V>::add(sval[w], sval[i]);
        else { ++w; skey[w] = skey[i]; sval[w] = sval[i]; }
      }
      un = w - lo + 1;
    }
    u[j] = un;
    cnt[b] = un;  // the level-1 starts are no longer needed: cnt now counts the distinct columns per bucket
  }
  __syncthreads();
  uniq = esc_scan8<TT>(cnt, rt, s_warp, s_mx);
#pragma unroll
  for (int j = 0; j < ESC_ITEMS; ++j) {
    const u32 b = rt + j * TT;
    if (u[j]) {
      const u32 lo = b ? sub[b - 1] : 0u, off = cnt[b];
      for (u32 i = 0; i < u[j]; ++i) {
        if (off + i < room) { c_col[cbase + off + i] = skey[lo + i]; c_val[cbase + off + i] = sval[lo + i]; }
      }
    }
  }
  return true;
}

// the same on n products staged in skey/sval[0, n) (staging complete: block barrier before the call)
template <class V, int TT>
__device__ __forceinline__ bool esc_finish(u32* skey, V* sval, u32* cnt, u32* sub, u32* s_warp, u32* s_mx, u32 n,
                                           u32 kmin, int bshift, int rt, u32* __restrict__ c_col,
                                           V* __restrict__ c_val, u64 cbase, u32 room, u32& uniq) {
  u32 k[ESC_ITEMS];
  V v[ESC_ITEMS];
  u32 vmask = 0;
#pragma unroll
  for (int i = 0; i < ESC_ITEMS; ++i) {
    const u32 p = rt + i * TT;
    k[i] = kmin; v[i] = Num<V>::zero();
    if (p < n) { k[i] = skey[p]; v[i] = sval[p]; vmask |= 1u << i; }
  }
  __syncthreads();  // everybody holds its products: the staging arrays may be overwritten
  return esc_core<V, TT>(k, v, vmask, skey, sval, cnt, sub, s_warp, s_mx, n, kmin, bshift, rt, c_col, c_val, cbase, room, uniq);
}

__device__ __forceinline__ void esc_give_back(Counters* cnt_dev, u32* fb_list, u32 row) {
  const u32 i = atomicAdd(&cnt_dev->fb_list_n, 1u);
  fb_list[i] = row;
  atomicAdd(&cnt_dev->fb_esc, 1u);
}

template <class V, int NW>
constexpr size_t num_esc_smem() { return (size_t)32 * NW * ESC_ITEMS * (sizeof(V) + 4 + 4 + 4); }

// One block per row: products <= FCAP = 8 * 32 * NW and entries of the A row <= products (no empty B rows among
// them, or few: the binning guarantees len(A row) <= products).
// Expansion with every load of the row in flight at once: the B row lengths of ALL entries of the A row are
// fetched in one step (a_col -> b_ptr, two dependent round trips for the whole row), scanned, and every thread
// then owns 8 CONSECUTIVE products, finds the entry of its first product by binary search in the scanned lengths
// (shared memory) and walks on — one more round trip (b_col, b_val) for the whole row.  The first version walked
// the A row 32 entries at a time with all warps in lock step: ~3 dependent DRAM round trips per 32 entries,
// 12-25 us per row, nothing else resident on the SM to hide them.
template <class V, int NW>
__global__ void __launch_bounds__(32 * NW)
k_num_esc(u32 n, const u32* __restrict__ perm, const u64* __restrict__ a_ptr, const u32* __restrict__ a_col,
          const V* __restrict__ a_val, const u64* __restrict__ b_ptr, const u32* __restrict__ b_col,
          const V* __restrict__ b_val, const u64* __restrict__ c_ptr, u32* __restrict__ c_col, V* __restrict__ c_val,
          Counters* cnt_dev, u32* fb_list) {
  constexpr int TT = 32 * NW, FCAP = TT * ESC_ITEMS, NB = FCAP;
  extern __shared__ __align__(16) unsigned char sm_esc[];
  V* sval = reinterpret_cast<V*>(sm_esc);        // [FCAP]
  u32* skey = reinterpret_cast<u32*>(sval + FCAP);  // [FCAP]
  u32* cnt = skey + FCAP;                        // [NB]
  u32* sub = cnt + NB;                           // [FCAP]
  u32* s_base = skey;                            // until the scatter: scanned B row lengths of the A row's entries
  __shared__ u32 s_warp[32];
  __shared__ u32 s_kmin, s_kmax, s_mx;
  const int rt = threadIdx.x, lane = rt & 31;
  if (blockIdx.x >= n) return;
  const u32 row = perm ? perm[blockIdx.x] : blockIdx.x;
  const u64 c0 = c_ptr[row];
  const u32 z = (u32)(c_ptr[row + 1] - c0);
  if (z == 0) return;
  const u64 lo = a_ptr[row];
  const u32 alen = (u32)min((u64)FCAP, a_ptr[row + 1] - lo);
#pragma unroll
  for (int i = 0; i < ESC_ITEMS; ++i) {
    const u32 e = rt + i * TT;
    u32 len = 0;
    if (e < alen) { const u32 kk = a_col[lo + e]; len = (u32)(b_ptr[kk + 1] - b_ptr[kk]); }
    s_base[e] = len;
    cnt[e] = 0; sub[e] = 0;
  }
  if (rt == 0) { s_kmin = 0xFFFFFFFFu; s_kmax = 0; }
  __syncthreads();
  u32 nprod = esc_scan8<TT>(s_base, rt, s_warp, &s_mx);  // s_base[e] = first product of entry e
  if (nprod > (u32)FCAP) nprod = FCAP;
  // my products: [8 rt, 8 rt + 8)
  u32 k[ESC_ITEMS];
  V v[ESC_ITEMS];
  u32 vmask = 0, kmin = 0xFFFFFFFFu, kmax = 0;
  const u32 p0 = (u32)rt * ESC_ITEMS;
  if (p0 < nprod) {
    u32 l = 0, h = alen - 1;  // last entry e with s_base[e] <= p0
    while (l < h) {
      const u32 mid = (l + h + 1) >> 1;
      if (s_base[mid] <= p0) l = mid; else h = mid - 1;
    }
    u32 e = l;
    u32 kk = a_col[lo + e];
    u64 bl = b_ptr[kk];
    V av = a_val[lo + e];
    u32 ebase = s_base[e];
    u32 enext = e + 1 < alen ? s_base[e + 1] : nprod;
#pragma unroll
    for (int i = 0; i < ESC_ITEMS; ++i) {
      const u32 p = p0 + i;
      k[i] = 0; v[i] = Num<V>::zero();
      if (p < nprod) {
        while (p >= enext) {  // next entry with a non-empty B row
          ++e;
          ebase = enext;
          enext = e + 1 < alen ? s_base[e + 1] : nprod;
          if (p < enext) { kk = a_col[lo + e]; bl = b_ptr[kk]; av = a_val[lo + e]; }
        }
        const u64 addr = bl + (p - ebase);
        k[i] = b_col[addr];
        v[i] = Num<V>::mul(av, b_val[addr]);
        vmask |= 1u << i;
        kmin = min(kmin, k[i]);
        kmax = max(kmax, k[i]);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < ESC_ITEMS; ++i) { k[i] = 0; v[i] = Num<V>::zero(); }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(FULL, kmin, d));
    kmax = max(kmax, __shfl_xor_sync(FULL, kmax, d));
  }
  if (lane == 0 && kmin <= kmax) { atomicMin(&s_kmin, kmin); atomicMax(&s_kmax, kmax); }
  __syncthreads();  // also: everybody is done with s_base before the scatter reuses skey
  kmin = s_kmin;
  const u32 range = s_kmax - kmin;
  const int rbits = range ? 32 - __clz(range) : 0;
  constexpr int LGNB = 31 - __builtin_clz((unsigned)NB);
  const int bshift = rbits > LGNB ? rbits - LGNB : 0;
  u32 uniq;
  if (!esc_core<V, TT>(k, v, vmask, skey, sval, cnt, sub, s_warp, &s_mx, nprod, kmin, bshift, rt, c_col, c_val, c0, z, uniq)) {
    if (rt == 0) esc_give_back(cnt_dev, fb_list, row);
  }
}

constexpr int ESCH_T = 1024;
constexpr int ESCH_LGNBC = 13;  // coarse buckets per row: 8192 (= 8 * ESCH_T, so esc_scan8 applies)
template <class V>
constexpr size_t num_esc_heavy_smem() { return (size_t)ESCH_T * ESC_ITEMS * (sizeof(V) + 4 + 4 + 4 + 4); }

// Rows with more than 8192 products: persistent blocks, dynamic row queue.  A coarse histogram of the row over
// the column range of B picks column ranges of at most 8192 products; each range is staged (one sweep over the
// row's products, keeping those inside the range) and finished like a short row.  Every warp sweeps its own
// chunks of the A row (32 entries each), so 32 dependent load chains are in flight per block.
template <class V>
__global__ void __launch_bounds__(ESCH_T)
k_num_esc_heavy(u32 n, const u32* __restrict__ perm, const u64* __restrict__ a_ptr, const u32* __restrict__ a_col,
                const V* __restrict__ a_val, const u64* __restrict__ b_ptr, const u32* __restrict__ b_col,
                const V* __restrict__ b_val, const u64* __restrict__ c_ptr, u32* __restrict__ c_col,
                V* __restrict__ c_val, u32 b_cols, u32* work, Counters* cnt_dev, u32* fb_list) {
  constexpr int TT = ESCH_T, NW = TT / 32, FCAP = TT * ESC_ITEMS, NB = FCAP, NBC = 1 << ESCH_LGNBC;
  static_assert(NBC == TT * ESC_ITEMS, "esc_scan8 scans 8 counters per thread");
  extern __shared__ __align__(16) unsigned char sm_esc[];
  V* sval = reinterpret_cast<V*>(sm_esc);
  u32* skey = reinterpret_cast<u32*>(sval + FCAP);
  u32* cnt = skey + FCAP;
  u32* sub = cnt + NB;
  u32* ccnt = sub + FCAP;  // [NBC] coarse histogram, then its exclusive prefix
  __shared__ u32 s_warp[32];
  __shared__ u32 s_item, s_mx, s_n;
  const int rt = threadIdx.x, lane = rt & 31, wid = rt >> 5;
  const int cbits = b_cols > 1 ? 32 - __clz(b_cols - 1) : 0;
  const int cshift = cbits > ESCH_LGNBC ? cbits - ESCH_LGNBC : 0;
  for (;;) {
    if (rt == 0) s_item = atomicAdd(work, 1u);
    __syncthreads();
    const u32 item = s_item;
    __syncthreads();
    if (item >= n) break;
    const u32 row = perm ? perm[item] : item;
    const u64 c0 = c_ptr[row];
    const u32 z = (u32)(c_ptr[row + 1] - c0);
    if (z == 0) continue;
    const u64 lo = a_ptr[row], hi = a_ptr[row + 1];
#pragma unroll
    for (int i = 0; i < ESC_ITEMS; ++i) ccnt[rt + i * TT] = 0;
    __syncthreads();
    for (u64 ec = lo + 32ull * wid; ec < hi; ec += 32ull * NW) {
      const AChunk<u32> c = load_chunk<u32, false, true>(ec, hi, lane, a_col, nullptr, b_ptr);
      for (u32 p0 = 0; p0 < c.total; p0 += 32) {
        u64 addr;
        u32 dummy;
        locate<u32, false>(c, p0 + lane, addr, dummy);
        if (p0 + lane < c.total) atomicAdd(&ccnt[b_col[addr] >> cshift], 1u);
      }
    }
    __syncthreads();
    const u32 ftot = esc_scan8<TT>(ccnt, rt, s_warp, &s_mx);  // ccnt[i] = products in coarse buckets < i
    bool ok = true;
    u32 out = 0;
    u32 s = 0;
    while (s < (u32)NBC) {
      const u32 base_s = ccnt[s];
      // largest e in [s, NBC] with products(s .. e) <= FCAP (every thread does the same search)
      u32 l = s, h = NBC;
      while (l < h) {
        const u32 mid = (l + h + 1) >> 1;
        const u32 pm = (mid < (u32)NBC ? ccnt[mid] : ftot) - base_s;
        if (pm <= (u32)FCAP) l = mid; else h = mid - 1;
      }
      const u32 e = l;
      if (e == s) { ok = false; break; }  // one coarse bucket holds more than a range can stage
      const u32 nr = (e < (u32)NBC ? ccnt[e] : ftot) - base_s;
      if (nr) {
        const u32 clo = s << cshift;
        const u64 chi = (u64)e << cshift;
#pragma unroll
        for (int i = 0; i < ESC_ITEMS; ++i) { cnt[rt + i * TT] = 0; sub[rt + i * TT] = 0; }
        if (rt == 0) s_n = 0;
        __syncthreads();
        for (u64 ec = lo + 32ull * wid; ec < hi; ec += 32ull * NW) {
          const AChunk<V> c = load_chunk<V, true, true>(ec, hi, lane, a_col, a_val, b_ptr);
          for (u32 p0 = 0; p0 < c.total; p0 += 32) {
            u64 addr;
            V av;
            locate<V, true>(c, p0 + lane, addr, av);
            u32 key = 0;
            bool in = false;
            if (p0 + lane < c.total) { key = b_col[addr]; in = key >= clo && (u64)key < chi; }
            const unsigned m = __ballot_sync(FULL, in);
            if (m) {
              u32 wbase = 0;
              if (lane == 0) wbase = atomicAdd(&s_n, (u32)__popc(m));
              wbase = __shfl_sync(FULL, wbase, 0);
              if (in) {
                const u32 dst = wbase + __popc(m & ((1u << lane) - 1u));
                if (dst < (u32)FCAP) { skey[dst] = key; sval[dst] = Num<V>::mul(av, b_val[addr]); }
              }
            }
          }
        }
        __syncthreads();
        const u32 width = (u32)(chi - clo - 1);  // largest (key - clo) in the range
        const int rbits = width ? 32 - __clz(width) : 0;
        constexpr int LGNB = 31 - __builtin_clz((unsigned)NB);
        const int bshift = rbits > LGNB ? rbits - LGNB : 0;
        u32 uniq;
        if (!esc_finish<V, TT>(skey, sval, cnt, sub, s_warp, &s_mx, nr, clo, bshift, rt, c_col, c_val, c0 + out,
                               z - out, uniq)) { ok = false; break; }
        out += uniq;
        __syncthreads();
      }
      s = e;
    }
    if (!ok && rt == 0) esc_give_back(cnt_dev, fb_list, row);
    __syncthreads();
  }
}

}  // namespace
