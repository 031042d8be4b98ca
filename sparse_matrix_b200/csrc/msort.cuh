// msort.cuh — numeric bins for rows of C that do not compress, as a MERGE TREE (round 2).
//
// With sorted rows of B, the products of a row of C — in the reference's order: A-row storage order, then B-row order
// (mul_hash.rs:146-162) — are len(A row) sorted runs.  A row of C is their k-way merge with equal columns folded.
// The hash bins (rowhash.cuh) pay a probe, an atomicCAS and a CAS-loop atomicAdd(double) per product plus a two-pass
// drain: 17-27 warp instructions per product on R-MAT, where almost nothing folds.  The bucket sort (esc.cuh) leaves
// most lanes idle in buckets of 0-3 entries.  Here:
//   A  B-row lengths of all entries of the A row, scanned: run boundaries (u16, kept for all passes);
//   B  expansion: every thread owns 8 consecutive products, stores (column << 32 | product index) and the product's
//      value a_ik * b_kj in shared memory — the index makes every key unique and the order among equal columns the
//      reference's product order;
//   C  ceil(log2(len(A row))) merge passes between two shared buffers: in pass L the runs of 2^L consecutive A
//      entries are merged pairwise; every thread produces 8 consecutive output positions — a merge-path binary search
//      for its starting split, then a sequential two-way merge (and a new search whenever it crosses into the next
//      pair).  ~15 thread instructions per product and pass, every lane busy;
//   D  fold: a position whose column differs from its predecessor's is an entry of C; its thread adds the following
//      equal columns in product order (first product stored, the others added: mul_hash.rs:154-161 — bit-identical
//      floats), block scan of the entry counts, C written once, 8 consecutive entries per thread.
// No hash table, no atomics at all.  Bins: products <= 1024 / 2048 / 4096 / 8192 (the esc bins 11..14, same rule:
// nnz > 256, 2 nnz >= products, len(A row) <= products); rows beyond go to the global-table kernel as before.
#pragma once
#include "common.cuh"
#include "esc.cuh"

namespace {

template <class V, int NW>
constexpr size_t num_msort_smem() {
  return (size_t)32 * NW * ESC_ITEMS * (8 + 8 + sizeof(V) + 2) + 32;
}

template <class V, int NW>
__global__ void __launch_bounds__(32 * NW)
k_num_msort(u32 n, const u32* __restrict__ perm, const u64* __restrict__ a_ptr, const u32* __restrict__ a_col,
            const V* __restrict__ a_val, const u64* __restrict__ b_ptr, const u32* __restrict__ b_col,
            const V* __restrict__ b_val, const u64* __restrict__ c_ptr, u32* __restrict__ c_col, V* __restrict__ c_val) {
  constexpr int TT = 32 * NW, FCAP = TT * ESC_ITEMS;
  extern __shared__ __align__(16) unsigned char sm_ms[];
  u64* buf0 = reinterpret_cast<u64*>(sm_ms);                      // [FCAP]
  u64* buf1 = buf0 + FCAP;                                        // [FCAP]
  V* sval = reinterpret_cast<V*>(buf1 + FCAP);                    // [FCAP] value of product p
  unsigned short* sb = reinterpret_cast<unsigned short*>(sval + FCAP);  // [FCAP + 1 (+ pad)] first product of A entry e
  u32* tmp = reinterpret_cast<u32*>(buf1);                        // [FCAP] the same as u32 while it is being scanned
  __shared__ u32 s_warp[32];
  __shared__ u32 s_mx;
  const int rt = threadIdx.x, lane = rt & 31, wid = rt >> 5;
  if (blockIdx.x >= n) return;
  const u32 row = perm ? perm[blockIdx.x] : blockIdx.x;
  const u64 c0 = c_ptr[row];
  const u32 z = (u32)(c_ptr[row + 1] - c0);
  if (z == 0) return;
  const u64 lo = a_ptr[row];
  const u32 alen = (u32)min((u64)FCAP, a_ptr[row + 1] - lo);
  // A. run lengths
#pragma unroll
  for (int i = 0; i < ESC_ITEMS; ++i) {
    const u32 e = rt + i * TT;
    u32 len = 0;
    if (e < alen) { const u32 kk = a_col[lo + e]; len = (u32)(b_ptr[kk + 1] - b_ptr[kk]); }
    tmp[e] = len;
  }
  __syncthreads();
  u32 nprod = esc_scan8<TT>(tmp, rt, s_warp, &s_mx);  // tmp[e] = first product of entry e
  if (nprod > (u32)FCAP) nprod = FCAP;                // cannot happen: the bin holds rows of at most FCAP products
#pragma unroll
  for (int i = 0; i < ESC_ITEMS; ++i) {
    const u32 e = rt + i * TT;
    if (e < alen) sb[e] = (unsigned short)min(tmp[e], nprod);
  }
  if (rt == 0) sb[alen] = (unsigned short)nprod;
  // B. expansion: my products are [8 rt, 8 rt + 8)
  const u32 p0 = (u32)rt * ESC_ITEMS;
  const u32 pend = min(p0 + (u32)ESC_ITEMS, nprod);
  u32 e0 = 0;
  if (p0 < nprod) {
    u32 l = 0, h = alen - 1;  // last entry e with tmp[e] <= p0
    while (l < h) {
      const u32 mid = (l + h + 1) >> 1;
      if (tmp[mid] <= p0) l = mid; else h = mid - 1;
    }
    e0 = l;
    u32 e = l;
    u32 kk = a_col[lo + e];
    u64 bl = b_ptr[kk];
    V av = a_val[lo + e];
    u32 ebase = tmp[e];
    u32 enext = e + 1 < alen ? tmp[e + 1] : nprod;
#pragma unroll
    for (int i = 0; i < ESC_ITEMS; ++i) {
      const u32 p = p0 + i;
      if (p < nprod) {
        while (p >= enext) {  // next entry with a non-empty B row
          ++e;
          ebase = enext;
          enext = e + 1 < alen ? tmp[e + 1] : nprod;
          if (p < enext) { kk = a_col[lo + e]; bl = b_ptr[kk]; av = a_val[lo + e]; }
        }
        const u64 addr = bl + (p - ebase);
        buf0[p] = ((u64)b_col[addr] << 32) | p;
        sval[p] = Num<V>::mul(av, b_val[addr]);
      }
    }
  }
  __syncthreads();  // products staged, run boundaries in sb, tmp (= buf1) free
  // C. merge passes
  u64* src = buf0;
  u64* dst = buf1;
  for (u32 L = 0; (1u << L) < alen; ++L) {
    u32 p = p0, e = e0;
    while (p < pend) {
      while ((u32)sb[e + 1] <= p) ++e;  // entry of position p (p < nprod = sb[alen])
      const u32 eb = (e >> (L + 1)) << (L + 1);
      const u32 lo_p = sb[eb], mid = sb[min(alen, eb + (1u << L))], hi = sb[min(alen, eb + (2u << L))];
      const u32 stop = min(pend, hi);
      const u32 d = p - lo_p, la = mid - lo_p, lb = hi - mid;
      u32 ilo = d > lb ? d - lb : 0u, ihi = min(d, la);
      while (ilo < ihi) {  // merge path: how many of the first d outputs come from the left run
        const u32 i = (ilo + ihi) >> 1;
        if (src[lo_p + i] < src[mid + (d - 1 - i)]) ilo = i + 1; else ihi = i;
      }
      u32 ai = lo_p + ilo, bi = mid + (d - ilo);
      u64 ka = ai < mid ? src[ai] : ~0ull, kb = bi < hi ? src[bi] : ~0ull;
      for (; p < stop; ++p) {  // branch-free step: one store, one load (the first version's if / else ran both sides)
        const bool ta = ka < kb;
        dst[p] = ta ? ka : kb;
        ai += ta ? 1u : 0u;
        bi += ta ? 0u : 1u;
        const u32 nx = ta ? ai : bi;
        const u64 nv = nx < (ta ? mid : hi) ? src[nx] : ~0ull;
        ka = ta ? nv : ka;
        kb = ta ? kb : nv;
      }
      e = min(alen - 1, eb + (2u << L));  // first entry of the next pair (the walk above skips empty ones)
    }
    __syncthreads();
    u64* t = src; src = dst; dst = t;
  }
  // D. fold and write
  u32 hm = 0;
  if (p0 < nprod) {
    u32 prev = p0 ? (u32)(src[p0 - 1] >> 32) : 0xFFFFFFFFu;  // columns are < 2^32 - 1
#pragma unroll
    for (int i = 0; i < ESC_ITEMS; ++i) {
      const u32 p = p0 + i;
      if (p < nprod) {
        const u32 k = (u32)(src[p] >> 32);
        if (k != prev) hm |= 1u << i;
        prev = k;
      }
    }
  }
  const u32 heads = __popc(hm);
  const u32 incl = warp_incl_scan_u32(heads, lane);
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  u32 woff = 0;
#pragma unroll
  for (int i = 0; i < NW; ++i) woff += i < wid ? s_warp[i] : 0u;
  u32 o = woff + incl - heads;
#pragma unroll
  for (int i = 0; i < ESC_ITEMS; ++i) {
    if (hm >> i & 1) {
      const u32 p = p0 + i;
      const u64 me = src[p];
      const u32 k = (u32)(me >> 32);
      V acc = sval[(u32)me];
      for (u32 q = p + 1; q < nprod; ++q) {
        const u64 nx = src[q];
        if ((u32)(nx >> 32) != k) break;
        acc = Num<V>::add(acc, sval[(u32)nx]);
      }
      if (o < z) { c_col[c0 + o] = k; c_val[c0 + o] = acc; }
      ++o;
    }
  }
}

}  // namespace
