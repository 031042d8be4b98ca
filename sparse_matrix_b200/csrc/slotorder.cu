// slotorder.cu — the reference's UNSORTED output order (B2 = false), SURVEY §8f rank 2.
//
// `&a * &b` and the reference's only benchmark return CsrMatrix<T, false>: mul_hash_numeric drains its
// linear-probing map in SLOT order (spam_csr/src/mul_hash.rs:176-186, linprobe/src/map.rs:59-63).  For row i with
// z distinct columns the map has cap = max(16, 2 * npow2(z)) slots (shrink_to, map.rs:49-58); column j goes to the
// first free slot at or after (j * 107) & (cap - 1) (lib.rs:13,29-31, map.rs:66-93) at the moment of its FIRST
// product, products being visited in A-row storage order and, inside, B-row storage order (mul_hash.rs:145-162).
// The order is therefore a function of (columns of the row, first-appearance order) only.
//
// The device computes the row sorted (any bin), then this pass permutes it:
//   * z <= 32: one thread replays the row's products into a private copy of the table (shared memory,
//     [slot][thread]) exactly like the reference, then walks the slots and fetches each column's value from the
//     sorted row by binary search;
//   * longer rows: (1) every product finds its column in the sorted row (binary search) and lowers that column's
//     time stamp to its own position in the product order (atomicMin) — afterwards ts[j] = position of the first
//     product of column j; (2) PRIORITY insertion: all columns insert concurrently, but a slot held by a column
//     with a LATER time stamp is taken over (compare-and-swap) and the evicted column moves on — the layout this
//     converges to is the unique one in which every column sits behind earlier columns only, i.e. the sequential
//     layout (the determinism argument of phase-concurrent linear probing: Shun & Blelloch 2014); (3) the slots
//     are compacted in index order.  One block per row with ts and table in shared memory up to 4096 columns,
//     persistent blocks with global scratch beyond.
// Values are not recomputed: they are the sorted pass's sums, moved.
#include "common.cuh"
#include "rowhash.cuh"

namespace {

constexpr u32 SO_SMALL = 32;    // thread-per-row replay up to this many columns
constexpr u32 SO_MID = 4096;    // block-per-row with shared-memory table up to this many columns
constexpr int SO_ST = 128;      // threads of the replay kernel
constexpr int SO_MT = 256;      // threads of the block-per-row kernel
constexpr int SO_LT = 1024;     // threads of the global-scratch kernel

__device__ __forceinline__ u32 lower_bound_u32(const u32* __restrict__ a, u32 n, u32 key) {
  u32 lo = 0, hi = n;
  while (lo < hi) {
    const u32 mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// rows with 32 < z <= SO_MID go to list_m, longer ones to list_l; max z into cnt->max_nnz
__global__ void __launch_bounds__(256) k_so_classify(u64 m, const u64* __restrict__ c_ptr, u32* __restrict__ list_m,
                                                     u32* __restrict__ list_l, Counters* cnt) {
  const u64 row = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= m) return;
  const u64 z = c_ptr[row + 1] - c_ptr[row];
  if (z > SO_MID) { list_l[atomicAdd(&cnt->work_b, 1u)] = (u32)row; atomicMax(&cnt->max_nnz, (u32)min(z, (u64)0xFFFFFFFFull)); }
  else if (z > SO_SMALL) list_m[atomicAdd(&cnt->work_a, 1u)] = (u32)row;
}

template <class V>
__global__ void __launch_bounds__(SO_ST) k_so_replay(u64 m, const u64* __restrict__ a_ptr, const u32* __restrict__ a_col,
                                                     const u64* __restrict__ b_ptr, const u32* __restrict__ b_col,
                                                     const u64* __restrict__ c_ptr, const u32* __restrict__ s_col,
                                                     const V* __restrict__ s_val, u32* __restrict__ o_col,
                                                     V* __restrict__ o_val) {
  __shared__ u32 tab[2 * SO_SMALL * SO_ST];  // [slot][thread]
  const u64 row = (u64)blockIdx.x * SO_ST + threadIdx.x;
  if (row >= m) return;
  const u64 c0 = c_ptr[row];
  const u32 z = (u32)min(c_ptr[row + 1] - c0, (u64)(SO_SMALL + 1));
  if (z == 0 || z > SO_SMALL) return;
  const u32 cap = table_size_u32(z), mask = cap - 1;  // map.rs:49-58
  u32* t = tab + threadIdx.x;
  for (u32 s = 0; s < cap; ++s) t[s * SO_ST] = EMPTY_KEY;
  const u64 lo = a_ptr[row], hi = a_ptr[row + 1];
  for (u64 e = lo; e < hi; ++e) {
    const u32 k = a_col[e];
    const u64 bl = b_ptr[k], bh = b_ptr[k + 1];
    for (u64 j = bl; j < bh; ++j) {
      const u32 key = b_col[j];
      u32 s = slot_of(key, mask);
      for (;;) {
        const u32 cur = t[s * SO_ST];
        if (cur == key) break;
        if (cur == EMPTY_KEY) { t[s * SO_ST] = key; break; }
        s = (s + 1) & mask;
      }
    }
  }
  u32 w = 0;
  for (u32 s = 0; s < cap; ++s) {  // drain in slot order (map.rs:59-63)
    const u32 key = t[s * SO_ST];
    if (key != EMPTY_KEY && w < z) {
      const u32 j = lower_bound_u32(s_col + c0, z, key);
      o_col[c0 + w] = key;
      o_val[c0 + w] = s_val[c0 + j];
      ++w;
    }
  }
}

// Steps (1)-(3) of the header on one row; ts / tab may be shared or global memory (generic pointers).  TT threads.
template <class V, int TT>
__device__ __forceinline__ void so_row(u32 row, u32* ts, u32* tab, u32* s_warp, const u64* __restrict__ a_ptr,
                                       const u32* __restrict__ a_col, const u64* __restrict__ b_ptr,
                                       const u32* __restrict__ b_col, const u64* __restrict__ c_ptr,
                                       const u32* __restrict__ s_col, const V* __restrict__ s_val,
                                       u32* __restrict__ o_col, V* __restrict__ o_val) {
  const int rt = threadIdx.x, lane = rt & 31, wid = rt >> 5;
  const u64 c0 = c_ptr[row];
  const u32 z = (u32)(c_ptr[row + 1] - c0);
  const u32 cap = 2u * npow2_u32(z), mask = cap - 1;  // z > 32: max(16, .) is moot
  for (u32 j = rt; j < z; j += TT) ts[j] = 0xFFFFFFFFu;
  for (u32 s = rt; s < cap; s += TT) tab[s] = EMPTY_KEY;
  __syncthreads();
  // (1) first-appearance time stamps: position of every product in the reference's enumeration order
  const u64 lo = a_ptr[row], hi = a_ptr[row + 1];
  u32 base = 0;
  for (u64 ec = lo; ec < hi; ec += 32) {
    const AChunk<u32> c = load_chunk<u32, false, true>(ec, hi, lane, a_col, nullptr, b_ptr);
    for (u32 p0 = 32u * wid; p0 < c.total; p0 += TT) {
      u64 addr;
      u32 dummy;
      locate<u32, false>(c, p0 + lane, addr, dummy);
      if (p0 + lane < c.total) {
        const u32 j = lower_bound_u32(s_col + c0, z, b_col[addr]);
        atomicMin(&ts[j], base + p0 + lane);
      }
    }
    base += c.total;
  }
  __syncthreads();
  // (2) priority insertion: tab[s] = index (into the sorted row) of the column that owns slot s
  for (u32 j = rt; j < z; j += TT) {
    u32 me = j;
    u32 tme = ts[me];
    u32 s = slot_of(s_col[c0 + me], mask);
    for (;;) {
      const u32 cur = *(volatile u32*)&tab[s];
      if (cur == EMPTY_KEY) {
        if (atomicCAS(&tab[s], EMPTY_KEY, me) == EMPTY_KEY) break;
        continue;  // somebody else took it meanwhile: look at the slot again
      }
      const u32 tcur = ts[cur];
      if (tcur < tme) { s = (s + 1) & mask; continue; }  // an earlier column: it stays, move on
      if (atomicCAS(&tab[s], cur, me) == cur) { me = cur; tme = tcur; s = (s + 1) & mask; }  // evicted: it moves on
    }
  }
  __syncthreads();
  // (3) compaction in slot order: thread rt owns slots [rt * per, rt * per + per)
  const u32 per = (cap + TT - 1) / TT;
  const u32 s0 = min(cap, (u32)rt * per), s1 = min(cap, s0 + per);
  u32 mine = 0;
  for (u32 s = s0; s < s1; ++s) mine += tab[s] != EMPTY_KEY ? 1u : 0u;
  const u32 x = warp_incl_scan_u32(mine, lane);
  if (lane == 31) s_warp[wid] = x;
  __syncthreads();
  u32 w = x - mine;
  for (int i = 0; i < wid; ++i) w += s_warp[i];
  for (u32 s = s0; s < s1; ++s) {
    const u32 j = tab[s];
    if (j != EMPTY_KEY) {
      o_col[c0 + w] = s_col[c0 + j];
      o_val[c0 + w] = s_val[c0 + j];
      ++w;
    }
  }
  __syncthreads();
}

template <class V>
__global__ void __launch_bounds__(SO_MT) k_so_mid(const u32* __restrict__ list, const u32* n_dev, u32* work,
                                                  const u64* __restrict__ a_ptr, const u32* __restrict__ a_col,
                                                  const u64* __restrict__ b_ptr, const u32* __restrict__ b_col,
                                                  const u64* __restrict__ c_ptr, const u32* __restrict__ s_col,
                                                  const V* __restrict__ s_val, u32* __restrict__ o_col,
                                                  V* __restrict__ o_val) {
  extern __shared__ u32 sm_so[];
  u32* ts = sm_so;             // [SO_MID]
  u32* tab = sm_so + SO_MID;   // [4 * SO_MID]
  __shared__ u32 s_warp[32];
  __shared__ u32 s_item;
  const u32 n = *n_dev;
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(work, 1u);
    __syncthreads();
    const u32 item = s_item;
    __syncthreads();
    if (item >= n) break;
    so_row<V, SO_MT>(list[item], ts, tab, s_warp, a_ptr, a_col, b_ptr, b_col, c_ptr, s_col, s_val, o_col, o_val);
  }
}

template <class V>
__global__ void __launch_bounds__(SO_LT) k_so_long(const u32* __restrict__ list, const u32* n_dev, u32* work,
                                                   u32* scratch, u64 stride, const u64* __restrict__ a_ptr,
                                                   const u32* __restrict__ a_col, const u64* __restrict__ b_ptr,
                                                   const u32* __restrict__ b_col, const u64* __restrict__ c_ptr,
                                                   const u32* __restrict__ s_col, const V* __restrict__ s_val,
                                                   u32* __restrict__ o_col, V* __restrict__ o_val) {
  __shared__ u32 s_warp[32];
  __shared__ u32 s_item;
  u32* ts = scratch + (u64)blockIdx.x * stride;  // [zmax]
  u32* tab = ts + stride / 5;                    // [4 * zmax]
  const u32 n = *n_dev;
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(work, 1u);
    __syncthreads();
    const u32 item = s_item;
    __syncthreads();
    if (item >= n) break;
    so_row<V, SO_LT>(list[item], ts, tab, s_warp, a_ptr, a_col, b_ptr, b_col, c_ptr, s_col, s_val, o_col, o_val);
  }
}

template <class V>
int slot_order_typed(spam_handle* h, const spam_dcsr* a, const spam_dcsr* b, spam_dcsr* c, u32* o_col, void* o_val_) {
  V* o_val = (V*)o_val_;
  const u64 m = a->rows;
  DevGuard g(h);
  u32 *list_m = nullptr, *list_l = nullptr;
  CKS(g.alloc(&list_m, m));
  CKS(g.alloc(&list_l, m));
  CK(cudaMemsetAsync(h->d_cnt, 0, sizeof(Counters), h->stream));
  k_so_classify<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(m, c->ptr, list_m, list_l, h->d_cnt);
  k_so_replay<V><<<(unsigned)((m + SO_ST - 1) / SO_ST), SO_ST, 0, h->stream>>>(m, a->ptr, a->idx, b->ptr, b->idx, c->ptr, c->idx,
                                                                                (const V*)c->val, o_col, o_val);
  count_launch(h, 2);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h->h_cnt, h->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const u32 n_mid = h->h_cnt->work_a, n_long = h->h_cnt->work_b, zmax = h->h_cnt->max_nnz;
  // the lists' lengths stay on the device; the row queues of the two kernels are work_c and total_nnz's low word
  if (n_mid) {
    unsigned grid = (unsigned)h->num_sms * 2;
    if (grid > n_mid) grid = n_mid;
    constexpr size_t smem = (size_t)5 * SO_MID * sizeof(u32);
    CK(cudaFuncSetAttribute(k_so_mid<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_so_mid<V><<<grid, SO_MT, smem, h->stream>>>(list_m, &h->d_cnt->work_a, &h->d_cnt->work_c, a->ptr, a->idx, b->ptr, b->idx, c->ptr,
                                               c->idx, (const V*)c->val, o_col, o_val);
    count_launch(h);
    CK(cudaGetLastError());
  }
  if (n_long) {
    const u64 stride = 5ull * (2ull * npow2_u64(zmax));  // ts: zmax (rounded up), table: at most 4 * zmax
    unsigned grid = (unsigned)h->num_sms;
    if (grid > n_long) grid = n_long;
    u32* scratch = nullptr;
    CKS(g.alloc(&scratch, (size_t)grid * stride));
    k_so_long<V><<<grid, SO_LT, 0, h->stream>>>(list_l, &h->d_cnt->work_b, &h->d_cnt->max_flop, scratch, stride, a->ptr, a->idx, b->ptr,
                                                b->idx, c->ptr, c->idx, (const V*)c->val, o_col, o_val);
    count_launch(h);
    CK(cudaGetLastError());
  }
  return SPAM_OK;
}

}  // namespace

// c: the product A * B with rows sorted by column (any bin).  Replaces c's col_idx / val by the same rows in the
// reference's slot order.  `b` must be the ORIGINAL right-hand side (its row order decides which product of a
// column comes first), not its sorted copy.
int slot_order_dev(spam_handle* h, const spam_dcsr* a, const spam_dcsr* b, spam_dcsr* c) {
  if (c->nnz == 0 || a->rows == 0) return SPAM_OK;
  DevGuard g(h);
  u32* o_col = nullptr;
  void* o_val = nullptr;
  CKS(g.alloc(&o_col, c->nnz));
  CKS(g.alloc_bytes(&o_val, c->nnz * dtype_size(c->dtype)));
  int st;
  switch (c->dtype) {
    case SPAM_F32: st = slot_order_typed<float>(h, a, b, c, o_col, o_val); break;
    case SPAM_F64: st = slot_order_typed<double>(h, a, b, c, o_col, o_val); break;
    case SPAM_I32: st = slot_order_typed<int32_t>(h, a, b, c, o_col, o_val); break;
    case SPAM_I64: st = slot_order_typed<int64_t>(h, a, b, c, o_col, o_val); break;
    default: st = spam_fail(h, SPAM_EINVAL, "bad dtype");
  }
  if (st != SPAM_OK) return st;
  if (c->owning) {
    dev_free(h, c->idx); dev_free(h, c->val);
    c->idx = o_col; c->val = o_val;
    g.release(o_col); g.release(o_val);
  } else {  // a view over caller-owned arrays: copy back
    CK(cudaMemcpyAsync(c->idx, o_col, c->nnz * sizeof(u32), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(c->val, o_val, c->nnz * dtype_size(c->dtype), cudaMemcpyDeviceToDevice, h->stream));
  }
  c->rows_sorted = -1;
  return SPAM_OK;
}
