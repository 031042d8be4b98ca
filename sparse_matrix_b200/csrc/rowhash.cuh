// rowhash.cuh — the hash bins: NW warps own one row of C and its linear-probing table in shared memory.
//
// Design points, each forced by a B200 measurement (profiles/r01_*):
//  * NO SHARED-MEMORY ATOMICS FOR VALUES WHEN ONE WARP OWNS THE ROW (NW = 1).  atomicAdd(double) on shared
//    memory is a CAS loop (~64+ cycles per warp instruction); with one per product the 27-point stencil ran
//    12x slower than its instruction count.  A warp that owns its table updates values with a plain
//    read-modify-write: in DIRECT mode a batch is (part of) one B row, whose columns are distinct, so the
//    lanes hold distinct slots; in FLAT mode equal columns inside a batch are found with
//    __match_any_sync and folded by the lowest lane in lane (= reference) order.  Sums are therefore
//    bit-identical to the reference (mul_hash.rs:145-162).  Only NEW keys use atomicCAS.  With NW > 1
//    warps sharing a table, values use atomicAdd.
//  * TWO ENUMERATIONS OF THE PRODUCTS.  DIRECT: one A entry (one B row) per batch, (start, length, a_ik)
//    broadcast by shuffle from the lane that loaded them — cheapest when B rows are about a warp long
//    (stencils).  FLAT: the B row lengths of 32 A entries are prefix-summed across the warp and every
//    batch of 32 consecutive products is spread over the lanes (5-step shuffle search) — all lanes busy
//    whatever the row lengths (R-MAT: mean 16, heavy tail).  The host picks per matrix (mean / max row
//    length of B, cached with the matrix).  Both follow the reference's product order.
//  * FIBONACCI HASH, high bits.  linprobe's (key*107) & (len-1) (linprobe/src/lib.rs:13,29-31) only
//    sees the low bits of the column: stencil planes with n^2 = 0 mod 256 all collide (15 probe rounds
//    per batch).  Table size, probing, sentinel and load factor stay linprobe's.
//  * SORT.  The B2 = true branch (mul_hash.rs:164-175) as a shared-memory bitonic network was as expensive
//    as the accumulation (NW = 1) or 85% of the kernel (block per row on R-MAT).  NW = 1: (column, index)
//    packed in one u32 and sorted in registers by warp shuffles.  NW > 1: occupied slots are counted into
//    order-preserving buckets over the row's column range, scanned, their slot indices scattered bucket by
//    bucket inside shared memory, and every entry then ranks itself among the 1-2 entries of its bucket
//    and stores (key, value) at its final place in C (one nearly coalesced pass over C).
//  * SHARED MEMORY THROUGH EXPLICIT 32-BIT SHARED ADDRESSES (ld/st/atom.shared): with generic pointers
//    the compiler re-derived the shared window (S2R SR_CgaCtaId + LEA) in every loop iteration.
#pragma once
#include "common.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr u32 SYM_REDO = 0xFFFFFFFFu;  // row_nnz marker: the optimistic symbolic table overflowed, redo the row
constexpr int ROWS_PER_BLOCK_W1 = 4;  // NW = 1: four independent warps (rows) per 128-thread block
// Two changes to the team / global-table kernels were tried in r2 and dropped (R-MAT 22, numeric pass 79 ms):
//  * groups of four warps taking separate 32-entry chunks of the A row, to save the chunk prologue every warp pays:
//    109 ms — all warps on one chunk is what keeps enough loads in flight;
//  * loading chunk k + 1 (a_col -> b_ptr) before chunk k is processed: 87 ms (and the symbolic pass 15.4 -> 16.5 ms).

__device__ __forceinline__ u32 slot_fib(u32 key, u32 shift) { return (key * 2654435769u) >> shift; }

// ---- shared-memory accessors on 32-bit shared addresses ------------------------------------------
__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ u32 lds32(u32 a) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(u32 a, u32 v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ u32 atoms_cas32(u32 a, u32 cmp, u32 val) {
  u32 old;
  asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(a), "r"(cmp), "r"(val) : "memory");
  return old;
}
template <class V> struct SV;
template <> struct SV<float> {
  static __device__ __forceinline__ float ld(u32 a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
  static __device__ __forceinline__ void st(u32 a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v)); }
};
template <> struct SV<double> {
  static __device__ __forceinline__ double ld(u32 a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
  static __device__ __forceinline__ void st(u32 a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v)); }
};
template <> struct SV<int32_t> {
  static __device__ __forceinline__ int32_t ld(u32 a) { int32_t v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
  static __device__ __forceinline__ void st(u32 a, int32_t v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v)); }
};
template <> struct SV<int64_t> {
  static __device__ __forceinline__ int64_t ld(u32 a) { long long v; asm volatile("ld.shared.s64 %0, [%1];" : "=l"(v) : "r"(a)); return (int64_t)v; }
  static __device__ __forceinline__ void st(u32 a, int64_t v) { asm volatile("st.shared.s64 [%0], %1;" ::"r"(a), "l"((long long)v)); }
};

__device__ __forceinline__ u32 warp_incl_scan_u32(u32 x, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u32 y = __shfl_up_sync(FULL, x, d);
    if (lane >= d) x += y;
  }
  return x;
}

// Find-or-insert for lanes with pending == true, table at shared address kbase (u32 slots).  On return s
// is the key's slot; fresh = this lane created the entry (exactly one lane per new key).
__device__ __forceinline__ void probe_insert(u32 kbase, u32 mask, u32 key, bool pending, u32& s, bool& fresh) {
  fresh = false;
  while (pending) {
    const u32 a = kbase + 4u * s;
    const u32 cur = lds32(a);
    if (cur == key) break;
    if (cur == EMPTY_KEY) {
      const u32 old = atoms_cas32(a, EMPTY_KEY, key);
      if (old == EMPTY_KEY) { fresh = true; break; }
      if (old == key) break;
    }
    s = (s + 1) & mask;
  }
}

// 32 A entries resident in lanes: B row start, length, inclusive prefix of lengths, a_ik
template <class V>
struct AChunk {
  u64 bl;
  u32 len, ps, total;
  V av;
};

template <class V, bool WITH_VAL, bool WITH_SCAN>
__device__ __forceinline__ AChunk<V> load_chunk(u64 ec, u64 hi, int lane, const u32* __restrict__ a_col,
                                                const V* __restrict__ a_val, const u64* __restrict__ b_ptr) {
  AChunk<V> c;
  c.bl = 0; c.len = 0; c.av = V(); c.ps = 0; c.total = 0;
  if (ec + lane < hi) {
    const u32 k = a_col[ec + lane];
    if (WITH_VAL) c.av = a_val[ec + lane];
    c.bl = b_ptr[k];
    c.len = (u32)(b_ptr[k + 1] - c.bl);
  }
  if (WITH_SCAN) {
    c.ps = warp_incl_scan_u32(c.len, lane);
    c.total = __shfl_sync(FULL, c.ps, 31);
  }
  return c;
}

// FLAT: product p (0 <= p < total) of the chunk -> (address in B, a_ik): 5-step shuffle search for the
// first lane whose inclusive prefix exceeds p.
template <class V, bool WITH_VAL>
__device__ __forceinline__ void locate(const AChunk<V>& c, u32 p, u64& addr, V& av) {
  int e = 0;
#pragma unroll
  for (int step = 16; step > 0; step >>= 1) {
    const u32 t = __shfl_sync(FULL, c.ps, e + step - 1);
    if (t <= p) e += step;
  }
  const u32 pe = __shfl_sync(FULL, c.ps, e), le = __shfl_sync(FULL, c.len, e);
  const u64 bl = __shfl_sync(FULL, c.bl, e);
  if (WITH_VAL) av = __shfl_sync(FULL, c.av, e);
  addr = bl + (p - (pe - le));
}

// ------------------------------------------------------------------------------------------------
// SYMBOLIC: distinct columns of one row (mul_hash.rs:66-103).  NW = 1: ROWS_PER_BLOCK_W1 rows per block.
// ------------------------------------------------------------------------------------------------
template <int NW, int CAP, bool DIRECT>
__global__ void __launch_bounds__(NW == 1 ? 32 * ROWS_PER_BLOCK_W1 : 32 * NW)
k_sym_row(u32 n, const u32* __restrict__ perm, const u64* __restrict__ a_ptr, const u32* __restrict__ a_col,
          const u64* __restrict__ b_ptr, const u32* __restrict__ b_col, const u32* __restrict__ flop,
          u32* __restrict__ row_nnz, int mode) {
  // mode 0: the table is sized from the row's product count f (linprobe's rule, an upper bound on the distinct
  //         columns).
  // mode 1 (DIRECT only): OPTIMISTIC — all CAP keys whatever f is.  High-compression rows (27-point stencil:
  //         f = 729, 125 distinct) fit a table a quarter of the size f asks for, which doubles the resident
  //         warps; a row that passes CAP/2 distinct columns gives up and leaves SYM_REDO in row_nnz.
  // mode 2: only the rows that gave up (row_nnz == SYM_REDO), with the full-size table.
  static_assert(!DIRECT || NW == 1, "DIRECT enumeration is a single-warp mode");
  extern __shared__ u32 sm_sym_keys[];
  __shared__ u32 s_total;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const u32 item = NW == 1 ? blockIdx.x * ROWS_PER_BLOCK_W1 + wid : blockIdx.x;
  if (item >= n) return;  // NW == 1: the warp leaves alone (no block barrier below); NW > 1: whole block
  const u32 row = perm ? perm[item] : item;
  const u32 f = flop[row];
  if (mode == 2 && row_nnz[row] != SYM_REDO) return;
  if (f == 0) { if (threadIdx.x % (32 * NW) == 0) row_nnz[row] = 0; return; }
  const u32 kbase = smem_addr(sm_sym_keys + (NW == 1 ? wid * CAP : 0));
  const int rw = NW == 1 ? 0 : wid;       // warp index inside the row's team
  const int rt = rw * 32 + lane;          // thread index inside the team
  u32 cap = table_size_u32(f);
  if (cap > (u32)CAP) cap = CAP;
  const bool optimistic = DIRECT && mode == 1;
  const u32 mask = cap - 1, shift = 32 - (31 - __clz(cap));
  for (u32 s = rt; s < cap; s += 32 * NW) sts32(kbase + 4u * s, EMPTY_KEY);
  if (NW > 1 && threadIdx.x == 0) s_total = 0;
  if (NW == 1) __syncwarp(); else __syncthreads();
  const u64 lo = a_ptr[row], hi = a_ptr[row + 1];
  u32 cnt = 0;
  for (u64 ec = lo; ec < hi; ec += 32) {
    if (DIRECT) {
      const AChunk<u32> c = load_chunk<u32, false, false>(ec, hi, lane, a_col, nullptr, b_ptr);
      const int na = (int)((hi - ec) < 32 ? (hi - ec) : 32);
      u64 nbl = __shfl_sync(FULL, c.bl, 0);
      u32 nlen = __shfl_sync(FULL, c.len, 0);
      u32 nkey = ((u32)lane < nlen) ? b_col[nbl + lane] : 0u;
      for (int i = 0; i < na; ++i) {
        const u64 bl = nbl;
        const u32 len = nlen, key0 = nkey;
        if (i + 1 < na) {  // prefetch the first batch of the next B row
          nbl = __shfl_sync(FULL, c.bl, i + 1);
          nlen = __shfl_sync(FULL, c.len, i + 1);
          nkey = ((u32)lane < nlen) ? b_col[nbl + lane] : 0u;
        }
        {
          bool fresh;
          u32 s = slot_fib(key0, shift);
          probe_insert(kbase, mask, key0, (u32)lane < len, s, fresh);
          cnt += fresh ? 1u : 0u;
        }
        for (u32 j0 = 32; j0 < len; j0 += 32) {  // B rows longer than a warp
          const bool active = j0 + lane < len;
          const u32 key = active ? b_col[bl + j0 + lane] : 0u;
          bool fresh;
          u32 s = slot_fib(key, shift);
          probe_insert(kbase, mask, key, active, s, fresh);
          cnt += fresh ? 1u : 0u;
          if (optimistic && __any_sync(FULL, fresh)) {  // a long B row can add many keys: check per batch
            u32 tot = cnt;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) tot += __shfl_xor_sync(FULL, tot, d);
            if (tot > cap / 2) { if (lane == 0) row_nnz[row] = SYM_REDO; return; }
          }
        }
        if (optimistic) {  // at most 32 new keys since the last check: the table never fills up
          u32 tot = cnt;
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) tot += __shfl_xor_sync(FULL, tot, d);
          if (tot > cap / 2) { if (lane == 0) row_nnz[row] = SYM_REDO; return; }
        }
      }
    } else {
      const AChunk<u32> c = load_chunk<u32, false, true>(ec, hi, lane, a_col, nullptr, b_ptr);
      // software pipeline: the next batch's search + load is issued before the current batch is probed
      u32 p0 = 32 * rw;
      u64 addr;
      u32 dummy;
      u32 nkey = 0;
      bool nact = false;
      if (p0 < c.total) {
        locate<u32, false>(c, p0 + lane, addr, dummy);
        nact = p0 + lane < c.total;
        nkey = nact ? b_col[addr] : 0u;
      }
      while (p0 < c.total) {
        const u32 key = nkey;
        const bool active = nact;
        p0 += 32 * NW;
        if (p0 < c.total) {
          locate<u32, false>(c, p0 + lane, addr, dummy);
          nact = p0 + lane < c.total;
          nkey = nact ? b_col[addr] : 0u;
        }
        u32 s = slot_fib(key, shift);
        bool fresh;
        probe_insert(kbase, mask, key, active, s, fresh);
        cnt += fresh ? 1u : 0u;
      }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(FULL, cnt, d);
  if (NW == 1) {
    if (lane == 0) row_nnz[row] = cnt;  // mul_hash.rs:95
  } else {
    if (lane == 0 && cnt) atomicAdd(&s_total, cnt);
    __syncthreads();
    if (threadIdx.x == 0) row_nnz[row] = s_total;
  }
}

template <int NW, int CAP>
constexpr size_t sym_row_smem() { return (size_t)(NW == 1 ? ROWS_PER_BLOCK_W1 : 1) * CAP * sizeof(u32); }

// ------------------------------------------------------------------------------------------------
// register bitonic sort of packed u32 (column << IDXBITS | index): EPL elements per lane, element
// e = r*32 + lane.  Steps with j >= 32 pair registers of one lane; steps with j < 32 pair lanes.
// ------------------------------------------------------------------------------------------------
template <int EPL>
__device__ __forceinline__ void warp_sort_packed(u32 (&x)[EPL], int lane) {
#pragma unroll
  for (int k = 2; k <= 32 * EPL; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int jr = j >> 5;
#pragma unroll
        for (int r = 0; r < EPL; ++r) {
          if ((r & jr) == 0) {
            const bool up = (((r * 32) & k) == 0);
            const u32 a = x[r], b = x[r | jr];
            const u32 lo = min(a, b), hi = max(a, b);
            x[r] = up ? lo : hi;
            x[r | jr] = up ? hi : lo;
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < EPL; ++r) {
          const bool up = (((r * 32 + lane) & k) == 0);
          const bool lower = ((lane & j) == 0);
          const u32 other = __shfl_xor_sync(FULL, x[r], j);
          x[r] = (lower == up) ? min(x[r], other) : max(x[r], other);
        }
      }
    }
  }
}

template <class V, int EPL>
__device__ __forceinline__ void warp_sort_store(u32 kbase, u32 vbase, u32 z, int idxbits, int lane,
                                                u32* __restrict__ c_col, V* __restrict__ c_val, u64 c0) {
  u32 x[EPL];
#pragma unroll
  for (int r = 0; r < EPL; ++r) {
    const u32 e = r * 32 + lane;
    x[r] = (e < z) ? ((lds32(kbase + 4u * e) << idxbits) | e) : 0xFFFFFFFFu;
  }
  warp_sort_packed<EPL>(x, lane);
  const u32 imask = (1u << idxbits) - 1u;
#pragma unroll
  for (int r = 0; r < EPL; ++r) {
    const u32 e = r * 32 + lane;
    if (e < z) {
      c_col[c0 + e] = x[r] >> idxbits;
      c_val[c0 + e] = SV<V>::ld(vbase + (u32)sizeof(V) * (x[r] & imask));
    }
  }
}

// fallback for columns too wide to pack: bitonic network on the key/value pairs in shared memory
template <class V>
__device__ __forceinline__ void warp_bitonic_sort(u32 kbase, u32 vbase, u32 n2, int lane) {
  for (u32 k = 2; k <= n2; k <<= 1) {
    for (u32 j = k >> 1; j > 0; j >>= 1) {
      for (u32 p = lane; p < (n2 >> 1); p += 32) {
        const u32 i = 2 * p - (p & (j - 1));
        const u32 l = i + j;
        const bool up = (i & k) == 0;
        const u32 ki = lds32(kbase + 4u * i), kl = lds32(kbase + 4u * l);
        if ((ki > kl) == up && ki != kl) {
          sts32(kbase + 4u * i, kl); sts32(kbase + 4u * l, ki);
          const V vi = SV<V>::ld(vbase + (u32)sizeof(V) * i), vl = SV<V>::ld(vbase + (u32)sizeof(V) * l);
          SV<V>::st(vbase + (u32)sizeof(V) * i, vl); SV<V>::st(vbase + (u32)sizeof(V) * l, vi);
        }
      }
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// NUMERIC (mul_hash.rs:105-201)
// ------------------------------------------------------------------------------------------------
// NW = 1, DIRECT: the lanes of a batch hold distinct columns -> distinct slots: plain read-modify-write
template <class V>
__device__ __forceinline__ void accumulate_direct(u32 kbase, u32 vbase, u32 mask, u32 shift, u32 key, V prod, bool active) {
  bool fresh;
  u32 s = slot_fib(key, shift);
  probe_insert(kbase, mask, key, active, s, fresh);
  if (active) {
    const u32 va = vbase + (u32)sizeof(V) * s;
    SV<V>::st(va, fresh ? prod : Num<V>::add(SV<V>::ld(va), prod));  // first product stored, not added to 0
  }
  __syncwarp();
}

// NW = 1, FLAT: equal slots inside the batch are folded in lane (= reference) order; the lowest lane writes
template <class V>
__device__ __forceinline__ void accumulate_fold(u32 kbase, u32 vbase, u32 mask, u32 shift, u32 key, V prod, bool active,
                                                int lane) {
  bool fresh;
  u32 s = slot_fib(key, shift);
  probe_insert(kbase, mask, key, active, s, fresh);
  const unsigned peers = __match_any_sync(FULL, active ? s : (0x80000000u | (u32)lane));
  const bool leader = active && (__ffs(peers) - 1) == lane;
  const bool any_fresh = (__ballot_sync(FULL, fresh) & peers) != 0;
  const u32 va = vbase + (u32)sizeof(V) * s;
  V acc = prod;
  if (leader && !any_fresh) acc = Num<V>::add(SV<V>::ld(va), prod);
  unsigned rem = leader ? (peers & ~(1u << lane)) : 0u;
  while (__any_sync(FULL, rem != 0)) {  // rarely more than one round
    const int src = rem ? (__ffs(rem) - 1) : lane;
    const V pv = __shfl_sync(FULL, prod, src);
    if (rem) { acc = Num<V>::add(acc, pv); rem &= rem - 1; }
  }
  if (leader) SV<V>::st(va, acc);
  __syncwarp();
}

// team kernels: number of drain buckets, as many as fit without costing a resident block
// (per block: table CAP * (4 + sizeof V), counters (NBMAX + 1) * 4, ord CAP/2 * 2 bytes)
template <class V, int NW, int CAP>
struct NumRowCfg {
  static constexpr int NBMAX = CAP == 4096 ? CAP : (CAP <= 2048 ? CAP / 2 : CAP / 4);
};
constexpr u32 RANK_BUCKET_MAX = 512;  // longest bucket ranked by counting; beyond: whole-row bitonic fallback

template <class V, int NW, int CAP, bool DIRECT>
__global__ void __launch_bounds__(NW == 1 ? 32 * ROWS_PER_BLOCK_W1 : 32 * NW)
k_num_row(u32 n, const u32* __restrict__ perm, const u64* __restrict__ a_ptr, const u32* __restrict__ a_col,
          const V* __restrict__ a_val, const u64* __restrict__ b_ptr, const u32* __restrict__ b_col,
          const V* __restrict__ b_val, const u64* __restrict__ c_ptr, u32* __restrict__ c_col, V* __restrict__ c_val,
          int pack_ok, Counters* cnt_dev) {
  static_assert(!DIRECT || NW == 1, "DIRECT enumeration is a single-warp mode");
  constexpr int RPB = NW == 1 ? ROWS_PER_BLOCK_W1 : 1;
  constexpr int TT = 32 * NW;  // threads in the row's team
  constexpr u32 SZ = (u32)sizeof(V);
  extern __shared__ __align__(16) unsigned char sm_num_raw[];
  __shared__ u32 s_warp[32];
  __shared__ u32 s_kmin, s_kmax, s_maxcnt;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const u32 item = NW == 1 ? blockIdx.x * RPB + wid : blockIdx.x;
  if (item >= n) return;
  const u32 row = perm ? perm[item] : item;
  const u64 c0 = c_ptr[row];
  const u32 z = (u32)(c_ptr[row + 1] - c0);
  if (z == 0) return;  // mul_hash.rs:141-143
  V* vals = reinterpret_cast<V*>(sm_num_raw) + (NW == 1 ? wid * CAP : 0);                     // [RPB][CAP]
  u32* keys = reinterpret_cast<u32*>(reinterpret_cast<V*>(sm_num_raw) + RPB * CAP) + (NW == 1 ? wid * CAP : 0);
  u32* cnt = reinterpret_cast<u32*>(reinterpret_cast<V*>(sm_num_raw) + RPB * CAP) + RPB * CAP;  // NW > 1: [CAP/2+1]
  const u32 kbase = smem_addr(keys), vbase = smem_addr(vals);
  const int rw = NW == 1 ? 0 : wid, rt = rw * 32 + lane;
  u32 cap = table_size_u32(z);  // map.rs:49-58
  if (cap > (u32)CAP) cap = CAP;
  const u32 mask = cap - 1, shift = 32 - (31 - __clz(cap));
  for (u32 s = rt; s < cap; s += TT) { sts32(kbase + 4u * s, EMPTY_KEY); if (NW > 1) vals[s] = Num<V>::zero(); }
  if (NW > 1) {
    for (u32 b = rt; b <= (u32)NumRowCfg<V, NW, CAP>::NBMAX; b += TT) cnt[b] = 0;  // drain bucket counters
    if (threadIdx.x == 0) { s_kmin = 0xFFFFFFFFu; s_kmax = 0; s_maxcnt = 0; }
  }
  if (NW == 1) __syncwarp(); else __syncthreads();

  const u64 lo = a_ptr[row], hi = a_ptr[row + 1];
  u32 kmin = 0xFFFFFFFFu, kmax = 0;
  for (u64 ec = lo; ec < hi; ec += 32) {
    if (DIRECT) {
      const AChunk<V> c = load_chunk<V, true, false>(ec, hi, lane, a_col, a_val, b_ptr);
      const int na = (int)((hi - ec) < 32 ? (hi - ec) : 32);
      u64 nbl = __shfl_sync(FULL, c.bl, 0);
      u32 nlen = __shfl_sync(FULL, c.len, 0);
      V nav = __shfl_sync(FULL, c.av, 0);
      u32 nkey = 0;
      V nbv = Num<V>::zero();
      if ((u32)lane < nlen) { nkey = b_col[nbl + lane]; nbv = b_val[nbl + lane]; }
      for (int i = 0; i < na; ++i) {
        const u64 bl = nbl;
        const u32 len = nlen, key0 = nkey;
        const V av = nav, bv0 = nbv;
        if (i + 1 < na) {
          nbl = __shfl_sync(FULL, c.bl, i + 1);
          nlen = __shfl_sync(FULL, c.len, i + 1);
          nav = __shfl_sync(FULL, c.av, i + 1);
          nkey = 0;
          nbv = Num<V>::zero();
          if ((u32)lane < nlen) { nkey = b_col[nbl + lane]; nbv = b_val[nbl + lane]; }
        }
        accumulate_direct<V>(kbase, vbase, mask, shift, key0, Num<V>::mul(av, bv0), (u32)lane < len);
        for (u32 j0 = 32; j0 < len; j0 += 32) {
          const bool active = j0 + lane < len;
          u32 key = 0;
          V bv = Num<V>::zero();
          if (active) { key = b_col[bl + j0 + lane]; bv = b_val[bl + j0 + lane]; }
          accumulate_direct<V>(kbase, vbase, mask, shift, key, Num<V>::mul(av, bv), active);
        }
      }
    } else {
      const AChunk<V> c = load_chunk<V, true, true>(ec, hi, lane, a_col, a_val, b_ptr);
      u32 p0 = 32 * rw;
      u64 addr;
      V av;
      u32 nkey = 0;
      V nprod = Num<V>::zero();
      bool nact = false;
      if (p0 < c.total) {
        locate<V, true>(c, p0 + lane, addr, av);
        nact = p0 + lane < c.total;
        if (nact) { nkey = b_col[addr]; nprod = Num<V>::mul(av, b_val[addr]); }
      }
      while (p0 < c.total) {
        const u32 key = nkey;
        const V prod = nprod;
        const bool active = nact;
        p0 += TT;
        if (p0 < c.total) {  // next batch: search + loads in flight while this one is accumulated
          locate<V, true>(c, p0 + lane, addr, av);
          nact = p0 + lane < c.total;
          nkey = 0;
          nprod = Num<V>::zero();
          if (nact) { nkey = b_col[addr]; nprod = Num<V>::mul(av, b_val[addr]); }
        }
        if (NW == 1) {
          accumulate_fold<V>(kbase, vbase, mask, shift, key, prod, active, lane);
        } else {
          u32 s = slot_fib(key, shift);
          bool fresh;
          probe_insert(kbase, mask, key, active, s, fresh);
          if (active) {
            Num<V>::atomic_add(&vals[s], prod);
            kmin = min(kmin, key);  // column range of the row, for the bucket drain below
            kmax = max(kmax, key);
          }
        }
      }
    }
  }

  if (NW == 1) {
    __syncwarp();
    // drain (map.rs:59-63): compact occupied slots to the front, in place, 32 slots per step
    u32 run = 0;
    for (u32 base = 0; base < cap; base += 32) {
      u32 kk = 0xFFFFFFFFu;
      if (base + lane < cap) kk = lds32(kbase + 4u * (base + lane));
      V vv = Num<V>::zero();
      if (kk != EMPTY_KEY) vv = SV<V>::ld(vbase + SZ * (base + lane));
      const unsigned occ = __ballot_sync(FULL, kk != EMPTY_KEY);
      __syncwarp();
      if (kk != EMPTY_KEY) {
        const u32 pos = run + __popc(occ & ((1u << lane) - 1u));
        sts32(kbase + 4u * pos, kk);
        SV<V>::st(vbase + SZ * pos, vv);
      }
      run += __popc(occ);
      __syncwarp();
    }
    const u32 n2 = npow2_u32(z);
    if (pack_ok) {
      const int idxbits = 31 - __clz(n2 < 2 ? 2 : n2);
      if (n2 <= 32) warp_sort_store<V, 1>(kbase, vbase, z, idxbits, lane, c_col, c_val, c0);
      else if (n2 <= 64) warp_sort_store<V, 2>(kbase, vbase, z, idxbits, lane, c_col, c_val, c0);
      else if (n2 <= 128) warp_sort_store<V, 4>(kbase, vbase, z, idxbits, lane, c_col, c_val, c0);
      else if (CAP >= 512 && n2 <= 256) warp_sort_store<V, (CAP >= 512 ? 8 : 1)>(kbase, vbase, z, idxbits, lane, c_col, c_val, c0);
      else if (CAP >= 1024) warp_sort_store<V, (CAP >= 1024 ? 16 : 1)>(kbase, vbase, z, idxbits, lane, c_col, c_val, c0);
    } else {
      if (lane == 0) atomicAdd(&cnt_dev->fb_warp_bitonic, 1u);
      for (u32 s = z + lane; s < n2; s += 32) sts32(kbase + 4u * s, EMPTY_KEY);
      __syncwarp();
      warp_bitonic_sort<V>(kbase, vbase, n2, lane);
      for (u32 s = lane; s < z; s += 32) { c_col[c0 + s] = lds32(kbase + 4u * s); c_val[c0 + s] = SV<V>::ld(vbase + SZ * s); }
    }
    return;
  }

  // ---- NW > 1: order-preserving bucket drain -------------------------------------------------
  // The occupied slots are counted into NB buckets over the row's own column range, the counts scanned,
  // and the SLOT INDICES scattered bucket by bucket into `ord` (shared memory).  Then every entry finds its
  // rank inside its bucket by counting the smaller keys of that bucket (mostly 1-2 entries) and stores
  // (key, value) at its final position in C.  C is written once, nearly coalesced; nothing is sorted in
  // global memory (the previous version scattered into C and re-sorted the buckets there: 60% of the
  // kernel's instructions and two thirds of its stall samples, profiles/r01_rmat20_v3_team_drain.txt).
  __syncthreads();
  const u32 NB = min(cap, (u32)NumRowCfg<V, NW, CAP>::NBMAX);
  u16* ord = reinterpret_cast<u16*>(cnt + NumRowCfg<V, NW, CAP>::NBMAX + 1);  // [CAP/2] slot index of the i-th entry
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(FULL, kmin, d));
    kmax = max(kmax, __shfl_xor_sync(FULL, kmax, d));
  }
  if (lane == 0) { atomicMin(&s_kmin, kmin); atomicMax(&s_kmax, kmax); }
  __syncthreads();
  kmin = s_kmin;
  const u32 range = s_kmax - kmin;
  const int lgnb = 31 - __clz(NB);
  const int rbits = range ? 32 - __clz(range) : 0;
  const int bshift = rbits > lgnb ? rbits - lgnb : 0;
  for (u32 s = rt; s < cap; s += TT) {
    const u32 kk = lds32(kbase + 4u * s);
    if (kk != EMPTY_KEY) atomicAdd(&cnt[(kk - kmin) >> bshift], 1u);
  }
  __syncthreads();
  {
    const u32 chunk = (NB + TT - 1) / TT;
    const u32 b0 = rt * chunk, b1 = min(NB, b0 + chunk);
    u32 sum = 0, mx = 0;
    for (u32 b = b0; b < b1; ++b) { const u32 cc = cnt[b]; sum += cc; mx = max(mx, cc); }
    const u32 x = warp_incl_scan_u32(sum, lane);
    if (lane == 31) s_warp[wid] = x;
    mx = max(mx, __shfl_xor_sync(FULL, mx, 16));
    mx = max(mx, __shfl_xor_sync(FULL, mx, 8));
    mx = max(mx, __shfl_xor_sync(FULL, mx, 4));
    mx = max(mx, __shfl_xor_sync(FULL, mx, 2));
    mx = max(mx, __shfl_xor_sync(FULL, mx, 1));
    if (lane == 0 && mx) atomicMax(&s_maxcnt, mx);
    __syncthreads();
    u32 woff = 0;
    for (int w = 0; w < wid; ++w) woff += s_warp[w];
    u32 runb = woff + x - sum;
    for (u32 b = b0; b < b1; ++b) { const u32 cc = cnt[b]; cnt[b] = runb; runb += cc; }
  }
  __syncthreads();
  if (s_maxcnt <= RANK_BUCKET_MAX) {
    for (u32 s = rt; s < cap; s += TT) {
      const u32 kk = lds32(kbase + 4u * s);
      if (kk != EMPTY_KEY) {
        const u32 pos = atomicAdd(&cnt[(kk - kmin) >> bshift], 1u);  // afterwards cnt[b] = end of bucket b
        ord[pos] = (u16)s;
      }
    }
    __syncthreads();
    for (u32 p = rt; p < z; p += TT) {
      const u32 s = ord[p];
      const u32 kk = lds32(kbase + 4u * s);
      const u32 b = (kk - kmin) >> bshift;
      const u32 lo_b = b ? cnt[b - 1] : 0u, hi_b = cnt[b];
      u32 rank = 0;
      if (hi_b - lo_b > 1) {
        for (u32 j = lo_b; j < hi_b; ++j) rank += (lds32(kbase + 4u * (u32)ord[j]) < kk) ? 1u : 0u;
      }
      c_col[c0 + lo_b + rank] = kk;
      c_val[c0 + lo_b + rank] = SV<V>::ld(vbase + SZ * s);
    }
  } else {
    // pathological column distribution: compact in shared memory and run the bitonic network
    if (threadIdx.x == 0) { s_kmax = 0; atomicAdd(&cnt_dev->fb_team_bitonic, 1u); }  // s_kmax reused as the compaction cursor
    __syncthreads();
    for (u32 base = 0; base < cap; base += TT) {
      const u32 s = base + rt;
      u32 kk = EMPTY_KEY;
      V vv = Num<V>::zero();
      if (s < cap) { kk = keys[s]; vv = vals[s]; }
      __syncthreads();
      if (kk != EMPTY_KEY) { const u32 pos = atomicAdd(&s_kmax, 1u); keys[pos] = kk; vals[pos] = vv; }
      __syncthreads();
    }
    const u32 n2 = npow2_u32(z);
    for (u32 s = z + rt; s < n2; s += TT) keys[s] = EMPTY_KEY;
    __syncthreads();
    for (u32 k = 2; k <= n2; k <<= 1) {
      for (u32 j = k >> 1; j > 0; j >>= 1) {
        for (u32 p = rt; p < (n2 >> 1); p += TT) {
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
    for (u32 s = rt; s < z; s += TT) { c_col[c0 + s] = keys[s]; c_val[c0 + s] = vals[s]; }
  }
}

template <class V, int NW, int CAP>
constexpr size_t num_row_smem() {
  return NW == 1 ? (size_t)ROWS_PER_BLOCK_W1 * CAP * (sizeof(V) + 4)
                 : (size_t)CAP * (sizeof(V) + 4) + (size_t)(NumRowCfg<V, NW, CAP>::NBMAX + 1) * 4 + (size_t)CAP;
}

}  // namespace
