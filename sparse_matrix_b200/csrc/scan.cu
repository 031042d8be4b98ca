// scan.cu — single-pass exclusive prefix sum with decoupled look-back (u32 counts -> u64 offsets).
// Replaces the reference's sequential checked_inclusive_scan (spam_csr/src/lib.rs:267-274), which
// produces [0, v0, v0+v1, ...] of length n+1; so does this.
//
// HBM-bound: reads 4 B and writes 8 B per row exactly once.  Tiles are handed out by an atomic
// counter so a tile's predecessors are always resident or finished (forward progress for the
// look-back spin), and each tile publishes {flag,value} in ONE 64-bit word, so no fence is needed.
#include "common.cuh"

namespace {

constexpr int SCAN_T = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_T * SCAN_ITEMS;
constexpr u64 FLAG_AGG = 1ull << 62;   // tile aggregate available
constexpr u64 FLAG_PFX = 2ull << 62;   // inclusive prefix available
constexpr u64 VAL_MASK = (1ull << 62) - 1;

__device__ __forceinline__ u64 warp_incl_scan_u64(u64 x, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    u64 y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  return x;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_lookback(const u32* __restrict__ in, u64* __restrict__ out, u64 n,
                                                          volatile u64* state, u32* tile_counter, ull* total_out,
                                                          u32* max_out) {
  __shared__ u32 s_tile;
  __shared__ u64 s_warp[SCAN_T / 32];
  __shared__ u64 s_prefix;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  __syncthreads();
  const u32 tile = s_tile;
  const u64 ibase = (u64)tile * SCAN_TILE + (u64)tid * SCAN_ITEMS;

  u32 v[SCAN_ITEMS];
  if (ibase + SCAN_ITEMS <= n && (((uintptr_t)in & 15) == 0)) {
    const uint4* p = reinterpret_cast<const uint4*>(in + ibase);
    uint4 x = p[0], y = p[1];
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) v[i] = (ibase + i < n) ? in[ibase + i] : 0u;
  }
  u64 tsum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) tsum += v[i];
  if (max_out) {  // largest input value, for callers that pick a code path by it (dok.cu)
    u32 mx = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) mx = max(mx, v[i]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    if (lane == 0 && mx) atomicMax(max_out, mx);
  }

  const u64 incl = warp_incl_scan_u64(tsum, lane);
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();

  if (wid == 0) {
    u64 w = (lane < SCAN_T / 32) ? s_warp[lane] : 0;
    u64 wi = warp_incl_scan_u64(w, lane);
    if (lane < SCAN_T / 32) s_warp[lane] = wi - w;  // exclusive warp offsets
    const u64 block_total = __shfl_sync(0xffffffffu, wi, 31);
    if (lane == 0) state[tile] = (tile == 0 ? FLAG_PFX : FLAG_AGG) | block_total;
    u64 excl = 0;
    if (tile > 0) {
      long long p = (long long)tile - 1;  // nearest predecessor examined by lane 0
      for (;;) {
        const long long idx = p - lane;
        u64 s;
        if (idx >= 0) {
          do { s = state[idx]; } while ((s >> 62) == 0);
        } else {
          s = FLAG_PFX;  // virtual tile -1 with prefix 0
        }
        const unsigned pm = __ballot_sync(0xffffffffu, (s >> 62) == 2);
        const int first = pm ? (__ffs(pm) - 1) : 32;
        u64 c = (lane <= first) ? (s & VAL_MASK) : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        excl += c;
        if (pm) break;
        p -= 32;
      }
      if (lane == 0) state[tile] = FLAG_PFX | (excl + block_total);
    }
    if (lane == 0) s_prefix = excl;
  }
  __syncthreads();

  u64 run = s_prefix + s_warp[wid] + (incl - tsum);
  if (ibase + SCAN_ITEMS <= n && (((uintptr_t)out & 15) == 0)) {
    ulonglong2* q = reinterpret_cast<ulonglong2*>(out + ibase);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i += 2) {
      ulonglong2 o;
      o.x = run; run += v[i];
      o.y = run; run += v[i + 1];
      q[i / 2] = o;
    }
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
      if (ibase + i < n) out[ibase + i] = run;
      run += v[i];
    }
  }
  // the last tile's last thread holds the grand total (items past n are zeros)
  const u64 tile_end = (u64)(tile + 1) * SCAN_TILE;
  if (tid == SCAN_T - 1 && tile_end >= n) {
    out[n] = run;
    if (total_out) *total_out = run;
  }
}

}  // namespace

// Zeroed look-back state (one 64-bit word per tile) and tile counter from the handle's grow-only buffer.
int lookback_workspace(spam_handle* h, u64 tiles, u64** state, u32** tile_counter) {
  if (h->scan_ws_cap < tiles + 1) {
    if (h->scan_ws) CKS(dev_free(h, h->scan_ws));
    h->scan_ws = nullptr;
    h->scan_ws_cap = 0;
    u64* ws = nullptr;
    CKS(dev_alloc_t(h, &ws, 2 * tiles + 1));
    h->scan_ws = ws;
    h->scan_ws_cap = 2 * tiles + 1;
  }
  CK(cudaMemsetAsync(h->scan_ws, 0, (tiles + 1) * sizeof(u64), h->stream));
  *state = h->scan_ws;
  *tile_counter = reinterpret_cast<u32*>(h->scan_ws + tiles);
  return SPAM_OK;
}

int scan_u32_to_u64(spam_handle* h, const u32* in, u64* out, u64 n, ull* d_total, u32* d_max) {
  if (n == 0) {
    CK(cudaMemsetAsync(out, 0, sizeof(u64), h->stream));
    if (d_total) CK(cudaMemsetAsync(d_total, 0, sizeof(ull), h->stream));
    return SPAM_OK;
  }
  const u64 tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  // tile states + the tile counter live in one grow-only buffer owned by the handle: one memset, no
  // allocation per scan (the DOK->CSR build runs eight scans per call)
  if (h->scan_ws_cap < tiles + 1) {
    if (h->scan_ws) CKS(dev_free(h, h->scan_ws));
    h->scan_ws = nullptr;
    h->scan_ws_cap = 0;
    u64* ws = nullptr;
    CKS(dev_alloc_t(h, &ws, 2 * tiles + 1));
    h->scan_ws = ws;
    h->scan_ws_cap = 2 * tiles + 1;
  }
  u64* state = h->scan_ws;
  CK(cudaMemsetAsync(state, 0, (tiles + 1) * sizeof(u64), h->stream));
  u32* tile_counter = reinterpret_cast<u32*>(state + tiles);
  k_scan_lookback<<<(unsigned)tiles, SCAN_T, 0, h->stream>>>(in, out, n, state, tile_counter, d_total, d_max);
  count_launch(h);
  CK(cudaGetLastError());
  return SPAM_OK;
}
