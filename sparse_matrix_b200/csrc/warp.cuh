// warp.cuh — warp-per-row bins: a linear-probing table private to ONE warp.
//
// What the B200 measurements (profiles/r01_*) said, in order:
//  1. atomicAdd(double) on shared memory is a CAS loop costing ~64+ cycles per warp instruction; with
//     one per product the 27-point-stencil product ran 12x slower than its instruction count.
//     => values are accumulated with a PLAIN read-modify-write.  That is safe because the warp owns
//     the table and no two lanes of one instruction hold the same slot: one B row has distinct
//     columns (W = 32), and with several B rows per instruction (W < 32) equal slots are found with
//     __match_any_sync and folded by the lowest lane.  Only the insertion of a NEW key uses
//     atomicCAS (one per distinct column, not one per product).
//  2. linprobe's slot = (key*107) & (len-1) (linprobe/src/lib.rs:13,29-31) only looks at the low
//     log2(len) bits of the column: columns that differ by a multiple of the table size collide.  In
//     a 27-point stencil on an n^3 grid with n^2 = 0 mod 256 (n = 96, 160) the five z-planes of a row
//     land on the same slots and probing ran ~15 rounds per batch.
//     => slot = HIGH bits of a Fibonacci multiplicative hash.  Everything else is linprobe's design
//     (open addressing, linear probing, size max(16, 2*npow2(n)), u32::MAX = empty).  Slot order is
//     not observable: rows are emitted sorted by column.
//  3. the per-row bitonic sort in shared memory cost as much as the accumulation.
//     => the (column, slot) pairs are packed into one u32 and sorted in REGISTERS with warp shuffles
//     (min/max network, no index arithmetic, no barriers); values are gathered by slot afterwards.
//
// Batches follow A-row storage order, so with W = 32 every sum is accumulated exactly in the
// reference's order (mul_hash.rs:145-162): bit-identical floats.
#pragma once
#include "common.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int WARPS_PER_BLOCK = 4;

__device__ __forceinline__ u32 slot_fib(u32 key, u32 shift) { return (key * 2654435769u) >> shift; }

// Find-or-insert for the lanes with pending == true.  On return s is the key's slot; fresh = this
// lane created the entry (exactly one lane per new key).
__device__ __forceinline__ void warp_probe(u32* keys, u32 mask, u32 key, bool pending, u32& s, bool& fresh) {
  volatile u32* vkeys = keys;
  fresh = false;
  while (pending) {
    const u32 cur = vkeys[s];
    if (cur == key) break;
    if (cur == EMPTY_KEY) {
      const u32 old = atomicCAS(&keys[s], EMPTY_KEY, key);
      if (old == EMPTY_KEY) { fresh = true; break; }
      if (old == key) break;
    }
    s = (s + 1) & mask;
  }
}

// ------------------------------------------------------------------------------------------------
// register bitonic sort of packed u32 (column << IDXBITS | slot): EPL elements per lane, element
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
            const int e = r * 32;  // lane bits do not matter for (e & k) when k >= 64
            const bool up = ((e & k) == 0);
            const u32 a = x[r], b = x[r | jr];
            const u32 lo = min(a, b), hi = max(a, b);
            x[r] = up ? lo : hi;
            x[r | jr] = up ? hi : lo;
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < EPL; ++r) {
          const int e = r * 32 + lane;
          const bool up = ((e & k) == 0);
          const bool lower = ((lane & j) == 0);
          const u32 other = __shfl_xor_sync(FULL, x[r], j);
          x[r] = (lower == up) ? min(x[r], other) : max(x[r], other);
        }
      }
    }
  }
}

// sort the z compacted entries (keys[0..z), vals[0..z)) and write them to C.  n2 = npow2(z) <= 32*EPL.
template <class V, int EPL>
__device__ __forceinline__ void warp_sort_store(const volatile u32* keys, const volatile V* vals, u32 z, int idxbits,
                                                int lane, u32* __restrict__ c_col, V* __restrict__ c_val, u64 c0) {
  u32 x[EPL];
#pragma unroll
  for (int r = 0; r < EPL; ++r) {
    const u32 e = r * 32 + lane;
    x[r] = (e < z) ? ((keys[e] << idxbits) | e) : 0xFFFFFFFFu;
  }
  warp_sort_packed<EPL>(x, lane);
  const u32 imask = (1u << idxbits) - 1u;
#pragma unroll
  for (int r = 0; r < EPL; ++r) {
    const u32 e = r * 32 + lane;
    if (e < z) {
      c_col[c0 + e] = x[r] >> idxbits;
      c_val[c0 + e] = vals[x[r] & imask];
    }
  }
}

// fallback: bitonic sort of n2 key/value pairs in the warp's shared memory (columns too wide to pack)
template <class V>
__device__ __forceinline__ void warp_bitonic_sort(volatile u32* keys, volatile V* vals, u32 n2, int lane) {
  for (u32 k = 2; k <= n2; k <<= 1) {
    for (u32 j = k >> 1; j > 0; j >>= 1) {
      for (u32 p = lane; p < (n2 >> 1); p += 32) {
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
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// SYMBOLIC, one warp per row.  The A row is read 32 entries at a time into lanes (coalesced), each
// lane resolves its B row extent, and the batch loop broadcasts (start, length) by shuffle: the only
// dependent global load left in the loop is the B entry itself, prefetched one group ahead.
// Requires nnz(B) < 2^32 (u32 offsets); the host falls back to the block-per-row kernels otherwise.
// ------------------------------------------------------------------------------------------------
template <int CAP>
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK) k_sym_warp(u32 n, const u32* __restrict__ perm,
                                                                    const u64* __restrict__ a_ptr,
                                                                    const u32* __restrict__ a_col,
                                                                    const u64* __restrict__ b_ptr,
                                                                    const u32* __restrict__ b_col,
                                                                    const u32* __restrict__ flop,
                                                                    u32* __restrict__ row_nnz, int wshift) {
  extern __shared__ u32 sm_warp_keys[];  // [WARPS_PER_BLOCK][CAP]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const u32 item = blockIdx.x * WARPS_PER_BLOCK + wid;
  if (item >= n) return;  // whole warp leaves together; no block-level barrier in this kernel
  const u32 row = perm ? perm[item] : item;
  const u32 f = flop[row];
  if (f == 0) { if (lane == 0) row_nnz[row] = 0; return; }
  u32* keys = sm_warp_keys + wid * CAP;
  u32 cap = table_size_u32(f);
  if (cap > (u32)CAP) cap = CAP;
  const u32 mask = cap - 1, shift = 32 - (31 - __clz(cap));
  for (u32 s = lane; s < cap; s += 32) keys[s] = EMPTY_KEY;
  __syncwarp();
  const int W = 1 << wshift, sub = lane >> wshift, sl = lane & (W - 1), nsub = 32 >> wshift;
  const u64 lo = a_ptr[row], hi = a_ptr[row + 1];
  u32 cnt = 0;
  for (u64 ec = lo; ec < hi; ec += 32) {
    const int na = (int)((hi - ec) < 32 ? (hi - ec) : 32);
    u32 my_bl = 0, my_len = 0;
    if (lane < na) { const u32 k = a_col[ec + lane]; my_bl = (u32)b_ptr[k]; my_len = (u32)b_ptr[k + 1] - my_bl; }
    int g = sub;
    u32 nbl = __shfl_sync(FULL, my_bl, g & 31);
    u32 nlen = __shfl_sync(FULL, my_len, g & 31);
    if (g >= na) nlen = 0;
    u32 nkey = ((u32)sl < nlen) ? b_col[nbl + sl] : 0u;
    for (int i0 = 0; i0 < na; i0 += nsub) {
      const u32 bl = nbl, len = nlen, key0 = nkey;
      if (i0 + nsub < na) {
        g = i0 + nsub + sub;
        nbl = __shfl_sync(FULL, my_bl, g & 31);
        nlen = __shfl_sync(FULL, my_len, g & 31);
        if (g >= na) nlen = 0;
        nkey = ((u32)sl < nlen) ? b_col[nbl + sl] : 0u;
      }
      {
        bool fresh;
        u32 s = slot_fib(key0, shift);
        warp_probe(keys, mask, key0, (u32)sl < len, s, fresh);
        cnt += fresh ? 1u : 0u;
      }
      if (__any_sync(FULL, len > (u32)W)) {  // B rows longer than the sub-group
        for (u32 j0 = W; __any_sync(FULL, j0 < len); j0 += W) {
          const bool active = j0 + sl < len;
          const u32 key = active ? b_col[bl + j0 + sl] : 0u;
          bool fresh;
          u32 s = slot_fib(key, shift);
          warp_probe(keys, mask, key, active, s, fresh);
          cnt += fresh ? 1u : 0u;
        }
      }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(FULL, cnt, d);
  if (lane == 0) row_nnz[row] = cnt;
}

// ------------------------------------------------------------------------------------------------
// NUMERIC, one warp per row
// ------------------------------------------------------------------------------------------------
template <class V>
__device__ __forceinline__ void warp_accumulate(u32* keys, volatile V* vals, u32 mask, u32 shift, u32 key, V prod,
                                                bool active, int W, int nsub, unsigned submask, int lane) {
  bool fresh;
  u32 s = slot_fib(key, shift);
  warp_probe(keys, mask, key, active, s, fresh);
  if (W == 32) {
    // one B row: distinct columns, hence distinct slots: plain read-modify-write
    if (active) vals[s] = fresh ? prod : Num<V>::add(vals[s], prod);
  } else {
    // fold equal slots across sub-groups in sub-group (= A-row) order; the lowest lane writes
    const unsigned peers = __match_any_sync(FULL, active ? (ull)s : ((1ull << 32) | (ull)lane));
    const bool leader = active && (__ffs(peers) - 1) == lane;
    const bool any_fresh = (__ballot_sync(FULL, fresh) & peers) != 0;
    V acc = Num<V>::zero();
    bool have = false;
    if (leader && !any_fresh) { acc = vals[s]; have = true; }
    for (int g = 0; g < nsub; ++g) {
      const unsigned mg = peers & (submask << (g * W));
      const int psrc = mg ? (__ffs(mg) - 1) : lane;
      const V p = __shfl_sync(FULL, prod, psrc);
      if (leader && mg) { acc = have ? Num<V>::add(acc, p) : p; have = true; }
    }
    if (leader) vals[s] = acc;
  }
  __syncwarp();
}

template <class V, int CAP>
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK) k_num_warp(u32 n, const u32* __restrict__ perm,
                                                                    const u64* __restrict__ a_ptr,
                                                                    const u32* __restrict__ a_col,
                                                                    const V* __restrict__ a_val,
                                                                    const u64* __restrict__ b_ptr,
                                                                    const u32* __restrict__ b_col,
                                                                    const V* __restrict__ b_val,
                                                                    const u64* __restrict__ c_ptr,
                                                                    u32* __restrict__ c_col, V* __restrict__ c_val,
                                                                    int wshift, int pack_ok) {
  extern __shared__ __align__(16) unsigned char sm_warp_raw[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const u32 item = blockIdx.x * WARPS_PER_BLOCK + wid;
  if (item >= n) return;
  const u32 row = perm ? perm[item] : item;
  const u64 c0 = c_ptr[row];
  const u32 z = (u32)(c_ptr[row + 1] - c0);
  if (z == 0) return;
  volatile V* vals = reinterpret_cast<V*>(sm_warp_raw) + wid * CAP;                                  // [W][CAP]
  u32* keys = reinterpret_cast<u32*>(reinterpret_cast<V*>(sm_warp_raw) + WARPS_PER_BLOCK * CAP) + wid * CAP;
  volatile u32* vkeys = keys;
  u32 cap = table_size_u32(z);  // map.rs:49-58
  if (cap > (u32)CAP) cap = CAP;
  const u32 mask = cap - 1, shift = 32 - (31 - __clz(cap));
  for (u32 s = lane; s < cap; s += 32) keys[s] = EMPTY_KEY;
  __syncwarp();
  const int W = 1 << wshift, sub = lane >> wshift, sl = lane & (W - 1), nsub = 32 >> wshift;
  const unsigned submask = (W == 32) ? FULL : ((1u << W) - 1u);
  const u64 lo = a_ptr[row], hi = a_ptr[row + 1];
  for (u64 ec = lo; ec < hi; ec += 32) {
    const int na = (int)((hi - ec) < 32 ? (hi - ec) : 32);
    u32 my_bl = 0, my_len = 0;
    V my_av = Num<V>::zero();
    if (lane < na) {
      const u32 k = a_col[ec + lane];
      my_av = a_val[ec + lane];
      my_bl = (u32)b_ptr[k];
      my_len = (u32)b_ptr[k + 1] - my_bl;
    }
    int g = sub;
    u32 nbl = __shfl_sync(FULL, my_bl, g & 31);
    u32 nlen = __shfl_sync(FULL, my_len, g & 31);
    V nav = __shfl_sync(FULL, my_av, g & 31);
    if (g >= na) nlen = 0;
    u32 nkey = 0;
    V nbv = Num<V>::zero();
    if ((u32)sl < nlen) { nkey = b_col[nbl + sl]; nbv = b_val[nbl + sl]; }
    for (int i0 = 0; i0 < na; i0 += nsub) {
      const u32 bl = nbl, len = nlen, key0 = nkey;
      const V av = nav, bv0 = nbv;
      if (i0 + nsub < na) {
        g = i0 + nsub + sub;
        nbl = __shfl_sync(FULL, my_bl, g & 31);
        nlen = __shfl_sync(FULL, my_len, g & 31);
        nav = __shfl_sync(FULL, my_av, g & 31);
        if (g >= na) nlen = 0;
        nkey = 0;
        nbv = Num<V>::zero();
        if ((u32)sl < nlen) { nkey = b_col[nbl + sl]; nbv = b_val[nbl + sl]; }
      }
      warp_accumulate<V>(keys, vals, mask, shift, key0, Num<V>::mul(av, bv0), (u32)sl < len, W, nsub, submask, lane);
      if (__any_sync(FULL, len > (u32)W)) {
        for (u32 j0 = W; __any_sync(FULL, j0 < len); j0 += W) {
          const bool active = j0 + sl < len;
          u32 key = 0;
          V bv = Num<V>::zero();
          if (active) { key = b_col[bl + j0 + sl]; bv = b_val[bl + j0 + sl]; }
          warp_accumulate<V>(keys, vals, mask, shift, key, Num<V>::mul(av, bv), active, W, nsub, submask, lane);
        }
      }
    }
  }
  __syncwarp();
  // drain (map.rs:59-63): compact the occupied slots to the front, in place, 32 slots per step
  u32 run = 0;
  for (u32 base = 0; base < cap; base += 32) {
    u32 kk = 0xFFFFFFFFu;
    if (base + lane < cap) kk = vkeys[base + lane];
    V vv = Num<V>::zero();
    if (kk != EMPTY_KEY) vv = vals[base + lane];
    const unsigned occ = __ballot_sync(FULL, kk != EMPTY_KEY);
    __syncwarp();
    if (kk != EMPTY_KEY) {
      const u32 pos = run + __popc(occ & ((1u << lane) - 1u));
      vkeys[pos] = kk;
      vals[pos] = vv;
    }
    run += __popc(occ);
    __syncwarp();
  }
  // B2 = true branch: sort by column (mul_hash.rs:164-175)
  const u32 n2 = npow2_u32(z);
  if (pack_ok) {
    const int idxbits = 31 - __clz(n2 < 2 ? 2 : n2);  // log2(n2), >= 1
    if (n2 <= 32) warp_sort_store<V, 1>(vkeys, vals, z, idxbits, lane, c_col, c_val, c0);
    else if (n2 <= 64) warp_sort_store<V, 2>(vkeys, vals, z, idxbits, lane, c_col, c_val, c0);
    else if (n2 <= 128) warp_sort_store<V, 4>(vkeys, vals, z, idxbits, lane, c_col, c_val, c0);
    else if (CAP >= 512 && n2 <= 256) warp_sort_store<V, (CAP >= 512 ? 8 : 1)>(vkeys, vals, z, idxbits, lane, c_col, c_val, c0);
    else if (CAP >= 1024) warp_sort_store<V, (CAP >= 1024 ? 16 : 1)>(vkeys, vals, z, idxbits, lane, c_col, c_val, c0);
  } else {
    for (u32 s = z + lane; s < n2; s += 32) vkeys[s] = EMPTY_KEY;
    __syncwarp();
    warp_bitonic_sort<V>(vkeys, vals, n2, lane);
    for (u32 s = lane; s < z; s += 32) { c_col[c0 + s] = vkeys[s]; c_val[c0 + s] = vals[s]; }
  }
}

}  // namespace
