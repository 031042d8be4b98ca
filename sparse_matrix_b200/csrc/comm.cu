// comm.cu — the multi-GPU layer behind the C ABI: one process per GPU, NCCL for the rendezvous and
// the small exchanges, and the assembled product written by the ranks themselves into each other's memory.
//
// The reference splits the rows of C over its threads in flop-balanced contiguous blocks and lets every thread
// write its own slice of ONE output (spam_csr/src/mul_hash.rs:38-64, :121-128).  Across GPUs: rank r multiplies
// rows [row_start_r, row_start_r+1) of A by a replicated B.  What the ranks exchange:
//   * B (and A) once, from rank 0                                  spam_comm_broadcast (ncclBroadcast)
//   * per-rank nnz and row counts, 2 x u64                          ncclAllGather
//   * the shards of C.  NCCL has no all-gather-v, and a gather after the product costs more than the
//     product (r1: 20 ms for 0.6 ms of work).  Here every rank owns a buffer for the WHOLE C (cudaMalloc, exported
//     with cudaIpcGetMemHandle, mapped by every peer).  A rank's numeric kernels write its rows straight into
//     its own copy at their final, offset-fixed position (spgemm_numeric_into); a push kernel then stores that
//     slice into the same position of every peer's copy over NVLink (plain st.global on peer-mapped pointers,
//     16 bytes per thread per peer).  The rank's block is cut into sub-blocks of rows so that the push of
//     sub-block k runs (on a high-priority side stream) while the numeric kernels of sub-block k+1 run.
//     SPAM gather mode 1 does the same gather with one ncclBroadcast per source rank in one NCCL group instead
//     (the fallback when peer mapping is unavailable).
// NCCL is bound at run time (dlopen of the libnccl.so.2 the process already has, else the system one): the
// library has no link-time dependency on it and single-GPU users never load it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace {

constexpr int MAXR = 16;  // ranks per node

struct NcclApi {
  void* lib;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char* (*GetErrorString)(ncclResult_t);
};

NcclApi g_nccl = {};

int load_nccl(std::string* err) {
  if (g_nccl.lib) return SPAM_OK;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the process already uses (torch's)
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { if (err) *err = std::string("dlopen libnccl.so.2: ") + dlerror(); return SPAM_ECUDA; }
  NcclApi a = {};
  a.lib = lib;
#define BIND(field, name) *(void**)(&a.field) = dlsym(lib, name); if (!a.field) { if (err) *err = std::string("dlsym ") + name; return SPAM_ECUDA; }
  BIND(GetUniqueId, "ncclGetUniqueId")
  BIND(CommInitRank, "ncclCommInitRank")
  BIND(CommDestroy, "ncclCommDestroy")
  BIND(AllGather, "ncclAllGather")
  BIND(Broadcast, "ncclBroadcast")
  BIND(AllReduce, "ncclAllReduce")
  BIND(Send, "ncclSend")
  BIND(Recv, "ncclRecv")
  BIND(GroupStart, "ncclGroupStart")
  BIND(GroupEnd, "ncclGroupEnd")
  BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
  g_nccl = a;
  return SPAM_OK;
}

struct GatherBuf {
  void* local;       // cudaMalloc'ed (IPC-exportable), whole-C sized
  size_t cap;        // bytes
  void* peer[MAXR];  // peer[r] = rank r's buffer mapped here (peer[rank] = local); null in NCCL-only mode
  bool mapped;       // the peers' copies are mapped (a buffer first used in NCCL mode is not)
};

}  // namespace

struct CommState {
  ncclComm_t comm;
  int rank, world;
  bool peer_ok;            // cudaIpcOpenMemHandle worked for every peer
  GatherBuf buf[3];        // 0 row_ptr (u64), 1 col_idx (u32), 2 val
  unsigned char* d_x;      // device staging for the small exchanges: MAXR * 256 bytes send + recv
  unsigned char* h_x;      // pinned mirror
  cudaStream_t push;       // high-priority stream of the push kernels
  cudaEvent_t ev_ready, ev_pushed;
  int push_blocks;         // SPAM_PUSH_BLOCKS at init: grid of the push kernel (0: two blocks of 512 threads per SM)
  int push_split;          // SPAM_PUSH_SPLIT=1 at init: every block of the push kernel serves one peer (k_push_split)
};

namespace {

#define NCK(call)                                                                              \
  do {                                                                                         \
    ncclResult_t _r = (call);                                                                  \
    if (_r != ncclSuccess) return spam_fail(h, SPAM_ECUDA, (std::string(#call ": ") + g_nccl.GetErrorString(_r)).c_str()); \
  } while (0)

struct PeerDst {
  void* p[MAXR];
  int n;
};

// src[0, bytes) -> the same bytes of every dst.p[j]; src and all dst share their alignment (same offset into
// 256-byte aligned buffers).  Head and tail in 4-byte words, the body in 16-byte words.
__global__ void __launch_bounds__(512) k_push(const unsigned char* __restrict__ src, PeerDst dst, u64 bytes) {
  const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x, nth = (u64)gridDim.x * blockDim.x;
  const u64 mis = (16 - ((uintptr_t)src & 15)) & 15;
  const u64 head = mis < bytes ? mis : bytes;       // bytes before the first 16-byte boundary (multiple of 4)
  const u64 body = (bytes - head) / 16;
  const u64 tail0 = head + body * 16;
  const uint4* s16 = reinterpret_cast<const uint4*>(src + head);
  for (u64 i = tid; i < body; i += nth) {
    const uint4 v = __ldcs(s16 + i);
#pragma unroll 1
    for (int j = 0; j < dst.n; ++j) reinterpret_cast<uint4*>((unsigned char*)dst.p[j] + head)[i] = v;
  }
  const u64 nh = head / 4, nt = (bytes - tail0) / 4;
  if (tid < nh + nt) {
    const u64 off = tid < nh ? tid * 4 : tail0 + (tid - nh) * 4;
    const u32 v = *reinterpret_cast<const u32*>(src + off);
    for (int j = 0; j < dst.n; ++j) *reinterpret_cast<u32*>((unsigned char*)dst.p[j] + off) = v;
  }
}

// the same copy with the destinations split over the blocks: block b streams the whole slice to peer b % n (the slice
// is read n times out of L2 / HBM, every store stream has one destination)
__global__ void __launch_bounds__(512) k_push_split(const unsigned char* __restrict__ src, PeerDst dst, u64 bytes) {
  const int peer = blockIdx.x % dst.n;
  const u64 bpp = gridDim.x / dst.n;                       // blocks per peer (grid is a multiple of n)
  const u64 tid = (u64)(blockIdx.x / dst.n) * blockDim.x + threadIdx.x, nth = bpp * blockDim.x;
  unsigned char* d = (unsigned char*)dst.p[peer];
  const u64 mis = (16 - ((uintptr_t)src & 15)) & 15;
  const u64 head = mis < bytes ? mis : bytes;
  const u64 body = (bytes - head) / 16;
  const u64 tail0 = head + body * 16;
  const uint4* s16 = reinterpret_cast<const uint4*>(src + head);
  uint4* d16 = reinterpret_cast<uint4*>(d + head);
  for (u64 i = tid; i < body; i += nth) d16[i] = s16[i];
  const u64 nh = head / 4, nt = (bytes - tail0) / 4;
  if (tid < nh + nt) {
    const u64 off = tid < nh ? tid * 4 : tail0 + (tid - nh) * 4;
    *reinterpret_cast<u32*>(d + off) = *reinterpret_cast<const u32*>(src + off);
  }
}

int push_range(spam_handle* h, CommState* c, int which, u64 off_bytes, u64 bytes, bool dma) {
  if (!bytes) return SPAM_OK;
  if (dma) {  // copy engines instead of SMs: one peer-to-peer copy per destination
    for (int r = 1; r < c->world; ++r) {
      const int dst = (c->rank + r) % c->world;  // every rank starts with a different peer
      CK(cudaMemcpyAsync((unsigned char*)c->buf[which].peer[dst] + off_bytes, (const unsigned char*)c->buf[which].local + off_bytes,
                         bytes, cudaMemcpyDeviceToDevice, c->push));
    }
    return SPAM_OK;
  }
  PeerDst d;
  d.n = 0;
  for (int r = 0; r < c->world; ++r)
    if (r != c->rank) d.p[d.n++] = (unsigned char*)c->buf[which].peer[r] + off_bytes;
  if (!d.n) return SPAM_OK;
  u64 blocks = (bytes / 16 + 511) / 512;
  // measured (R-MAT 22): 8 GPUs: 32 blocks 472 GB/s received per GPU, two blocks per SM 552 GB/s; 2 GPUs: two blocks per SM
  // 73 ms per step, four blocks per SM with two loads in flight 81 ms (the push then takes SMs from the product)
  const u64 cap = c->push_blocks ? (u64)c->push_blocks : (u64)h->num_sms * 2;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  if (c->push_split && d.n > 1) {
    u64 bpp = (blocks + d.n - 1) / d.n;
    if (bpp == 0) bpp = 1;
    k_push_split<<<(unsigned)(bpp * d.n), 512, 0, c->push>>>((const unsigned char*)c->buf[which].local + off_bytes, d, bytes);
  } else {
    k_push<<<(unsigned)blocks, 512, 0, c->push>>>((const unsigned char*)c->buf[which].local + off_bytes, d, bytes);
  }
  count_launch(h);
  CK(cudaGetLastError());
  return SPAM_OK;
}

// all-gather of `bytes` (<= 256) per rank through the staging buffers; result on the host, in c->h_x + MAXR*256
int small_allgather(spam_handle* h, CommState* c, const void* mine, size_t bytes) {
  unsigned char* send_h = c->h_x;
  unsigned char* recv_h = c->h_x + MAXR * 256;
  unsigned char* send_d = c->d_x;
  unsigned char* recv_d = c->d_x + MAXR * 256;
  memcpy(send_h, mine, bytes);
  CK(cudaMemcpyAsync(send_d, send_h, bytes, cudaMemcpyHostToDevice, h->stream));
  NCK(g_nccl.AllGather(send_d, recv_d, bytes, ncclUint8, c->comm, h->stream));
  CK(cudaMemcpyAsync(recv_h, recv_d, bytes * c->world, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return SPAM_OK;
}

void close_peers(CommState* c, GatherBuf* b) {
  for (int r = 0; r < c->world; ++r) {
    if (r != c->rank && b->peer[r]) cudaIpcCloseMemHandle(b->peer[r]);
    b->peer[r] = nullptr;
  }
}

// Grow buffer `which` to at least `bytes` on EVERY rank (collective: the ranks agree on the new capacity, so
// either all reallocate or none) and map the peers' copies.
int ensure_buf(spam_handle* h, CommState* c, int which, size_t bytes, bool want_peers) {
  GatherBuf* b = &c->buf[which];
  u64 need = (bytes > b->cap || (want_peers && c->world > 1 && c->peer_ok && !b->mapped)) ? 1 : 0;
  // one small all-gather decides: any rank short of space -> everybody reallocates to the largest request
  u64 mine[2] = {need, (u64)bytes};
  CKS(small_allgather(h, c, mine, sizeof(mine)));
  const u64* all = reinterpret_cast<const u64*>(c->h_x + MAXR * 256);
  u64 any = 0, mx = 0;
  for (int r = 0; r < c->world; ++r) { any |= all[2 * r]; if (all[2 * r + 1] > mx) mx = all[2 * r + 1]; }
  if (!any) return SPAM_OK;
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaStreamSynchronize(c->push));
  close_peers(c, b);
  b->mapped = false;
  // nobody may free a buffer that a peer still has mapped: wait until every rank has closed its mappings
  { u64 z = 0; CKS(small_allgather(h, c, &z, sizeof(z))); }
  if (b->local) { CK(cudaFree(b->local)); b->local = nullptr; b->cap = 0; }
  size_t cap = (size_t)(mx + mx / 8 + 4096);  // headroom: the next product of a similar size fits
  cap = (cap + 255) & ~(size_t)255;
  cudaError_t e = cudaMalloc(&b->local, cap);
  if (e != cudaSuccess) { cudaGetLastError(); return spam_fail(h, SPAM_ENOMEM, "gather buffer allocation failed", e); }
  b->cap = cap;
  b->peer[c->rank] = b->local;
  if (!want_peers || c->world == 1) return SPAM_OK;
  cudaIpcMemHandle_t hd;
  CK(cudaIpcGetMemHandle(&hd, b->local));
  static_assert(sizeof(cudaIpcMemHandle_t) <= 256, "IPC handle does not fit the staging slot");
  CKS(small_allgather(h, c, &hd, sizeof(hd)));
  const unsigned char* hs = c->h_x + MAXR * 256;
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t ph;
    memcpy(&ph, hs + (size_t)r * sizeof(hd), sizeof(hd));
    e = cudaIpcOpenMemHandle(&b->peer[r], ph, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      b->peer[r] = nullptr;
      c->peer_ok = false;
    }
  }
  // every rank must know whether every mapping worked everywhere
  u64 ok = c->peer_ok ? 1 : 0;
  CKS(small_allgather(h, c, &ok, sizeof(ok)));
  const u64* oks = reinterpret_cast<const u64*>(c->h_x + MAXR * 256);
  for (int r = 0; r < c->world; ++r) if (!oks[r]) c->peer_ok = false;
  b->mapped = c->peer_ok;
  return SPAM_OK;
}

int barrier(spam_handle* h, CommState* c) {
  u32* d = reinterpret_cast<u32*>(c->d_x);
  NCK(g_nccl.AllReduce(d, d + 64, 1, ncclUint32, ncclSum, c->comm, h->stream));
  return SPAM_OK;
}

}  // namespace

extern "C" {

int spam_comm_unique_id(void* out128) {
  if (!out128) return SPAM_EINVAL;
  if (load_nccl(nullptr) != SPAM_OK) return SPAM_ECUDA;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return SPAM_ECUDA;
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, 128);
  return SPAM_OK;
}

int spam_comm_init(spam_handle* h, const void* id128, int rank, int world) {
  if (!h || !id128 || world < 1 || world > MAXR || rank < 0 || rank >= world) return spam_fail(h, SPAM_EINVAL, "bad argument");
  if (h->comm) return spam_fail(h, SPAM_ESTATE, "communicator already initialised");
  CK(cudaSetDevice(h->device));
  std::string err;
  if (load_nccl(&err) != SPAM_OK) return spam_fail(h, SPAM_ECUDA, err.c_str());
  CommState* c = new CommState();
  memset(c, 0, sizeof(*c));
  c->rank = rank; c->world = world; c->peer_ok = true;
  { const char* e = getenv("SPAM_PUSH_BLOCKS"); c->push_blocks = e ? atoi(e) : 0; }
  { const char* e = getenv("SPAM_PUSH_SPLIT"); c->push_split = e ? atoi(e) : 0; }
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) { delete c; return spam_fail(h, SPAM_ECUDA, g_nccl.GetErrorString(r)); }
  cudaError_t e = cudaMalloc((void**)&c->d_x, 2 * MAXR * 256);
  if (e == cudaSuccess) e = cudaMemset(c->d_x, 0, 2 * MAXR * 256);
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&c->h_x, 2 * MAXR * 256, cudaHostAllocDefault);
  int lo = 0, hi = 0;
  if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->push, cudaStreamNonBlocking, hi);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_pushed, cudaEventDisableTiming);
  h->comm = c;
  if (e != cudaSuccess) { spam_comm_destroy(h); return spam_fail(h, SPAM_ECUDA, "communicator resources", e); }
  return SPAM_OK;
}

int spam_comm_destroy(spam_handle* h) {
  if (!h) return SPAM_EINVAL;
  CommState* c = h->comm;
  if (!c) return SPAM_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (c->push) cudaStreamSynchronize(c->push);
  for (auto& b : c->buf) close_peers(c, &b);
  // collective when world > 1: nobody frees a buffer a peer still has mapped
  if (c->world > 1 && c->comm && c->d_x && c->h_x) { u64 z = 0; small_allgather(h, c, &z, sizeof(z)); }
  for (auto& b : c->buf) if (b.local) cudaFree(b.local);
  if (c->ev_ready) cudaEventDestroy(c->ev_ready);
  if (c->ev_pushed) cudaEventDestroy(c->ev_pushed);
  if (c->push) cudaStreamDestroy(c->push);
  if (c->d_x) cudaFree(c->d_x);
  if (c->h_x) cudaFreeHost(c->h_x);
  if (c->comm) g_nccl.CommDestroy(c->comm);
  delete c;
  h->comm = nullptr;
  return SPAM_OK;
}

int spam_comm_info(const spam_handle* h, int* rank, int* world, int* peer_mapped) {
  if (!h || !h->comm) return SPAM_ESTATE;
  if (rank) *rank = h->comm->rank;
  if (world) *world = h->comm->world;
  if (peer_mapped) *peer_mapped = h->comm->peer_ok ? 1 : 0;
  return SPAM_OK;
}

int spam_comm_broadcast(spam_handle* h, void* d_buf, uint64_t bytes, int root) {
  if (!h || !h->comm) return spam_fail(h, SPAM_ESTATE, "spam_comm_init first");
  if (!bytes) return SPAM_OK;
  if (!d_buf || root < 0 || root >= h->comm->world) return spam_fail(h, SPAM_EINVAL, "bad argument");
  CK(cudaSetDevice(h->device));
  NCK(g_nccl.Broadcast(d_buf, d_buf, bytes, ncclUint8, root, h->comm->comm, h->stream));
  return SPAM_OK;
}

int spam_comm_allgather_u64(spam_handle* h, const uint64_t* mine, uint32_t n, uint64_t* all) {
  if (!h || !h->comm) return spam_fail(h, SPAM_ESTATE, "spam_comm_init first");
  if (!mine || !all || n == 0 || n > 32) return spam_fail(h, SPAM_EINVAL, "1..32 values per rank");
  CK(cudaSetDevice(h->device));
  CKS(small_allgather(h, h->comm, mine, (size_t)n * 8));
  memcpy(all, h->comm->h_x + MAXR * 256, (size_t)n * 8 * h->comm->world);
  return SPAM_OK;
}

// all-gather-v as one NCCL group of in-place broadcasts: rank r's bytes sit at d_out + byte_offsets[r]
int spam_comm_allgatherv(spam_handle* h, void* d_out, const uint64_t* byte_offsets) {
  if (!h || !h->comm) return spam_fail(h, SPAM_ESTATE, "spam_comm_init first");
  if (!d_out || !byte_offsets) return spam_fail(h, SPAM_EINVAL, "bad argument");
  CommState* c = h->comm;
  CK(cudaSetDevice(h->device));
  NCK(g_nccl.GroupStart());
  for (int r = 0; r < c->world; ++r) {
    const u64 n = byte_offsets[r + 1] - byte_offsets[r];
    if (!n) continue;
    unsigned char* p = (unsigned char*)d_out + byte_offsets[r];
    ncclResult_t rr = g_nccl.Broadcast(p, p, n, ncclUint8, r, c->comm, h->stream);
    if (rr != ncclSuccess) { g_nccl.GroupEnd(); return spam_fail(h, SPAM_ECUDA, g_nccl.GetErrorString(rr)); }
  }
  NCK(g_nccl.GroupEnd());
  return SPAM_OK;
}

// C = A * B with A row-sharded and the result assembled on every rank.  a_block: this rank's rows
// [row_start, row_start + a_block.rows) of A (row_ptr rebased to 0); b: all of B, replicated.  *c is a NON-owning
// view of the handle's gather buffers (valid until the next gathered product or spam_comm_destroy; release the
// view itself with spam_dcsr_free).  nsub >= 1 sub-blocks pipeline the product with the exchange; mode 0 = peer
// stores by a kernel (falls back to 1 when the peers' buffers could not be mapped), 1 = grouped ncclBroadcast,
// 2 = peer-to-peer copies by the copy engines (cudaMemcpyAsync on the mapped buffers) instead of the push kernel.
// ms4 (optional, host): {symbolic + counts exchange, numeric (last sub-block done), whole call, 0}.
int spam_spgemm_gathered(spam_handle* h, const spam_dcsr* a_block, const spam_dcsr* b, uint64_t row_start,
                         uint64_t total_rows, int nsub, int mode, spam_dcsr** cout) {
  if (!h || !a_block || !b || !cout) return spam_fail(h, SPAM_EINVAL, "null argument");
  *cout = nullptr;
  CommState* c = h->comm;
  if (!c) return spam_fail(h, SPAM_ESTATE, "spam_comm_init first");
  if (row_start + a_block->rows > total_rows) return spam_fail(h, SPAM_EINVAL, "row block outside the matrix");
  if (nsub < 1) nsub = 1;
  if (nsub > 16) nsub = 16;
  if ((u64)nsub > a_block->rows) nsub = a_block->rows ? (int)a_block->rows : 1;
  CK(cudaSetDevice(h->device));
  const size_t es = dtype_size(a_block->dtype);
  const u64 m = a_block->rows;

  // ---- sub-blocks of my rows, balanced on the device-time estimate, and their symbolic passes ----
  u64 sub_start[17];
  sub_start[0] = 0; sub_start[nsub] = m;
  if (nsub > 1) {
    u64 tf = 0;
    CKS(spam_rows_to_parts_cost(h, a_block, b, (uint32_t)nsub, sub_start, &tf));
  }
  spam_dcsr view[16];
  SpgemmPending* pend[16] = {};
  u64 sub_nnz[16] = {};
  auto drop = [&]() { for (int s = 0; s < nsub; ++s) if (pend[s]) { spgemm_pending_free(h, pend[s]); pend[s] = nullptr; } };
  CKS(ensure_matrix_stats(h, a_block));
  // per-phase event timing is per product; a gathered product is several (time it from outside)
  struct TimingOff { spam_handle* h; bool was; ~TimingOff() { h->timing = was; } } timing_off{h, h->timing};
  h->timing = false;
  u64 my_nnz = 0, my_flops = 0;
  spam_stats acc = {};
  for (int s = 0; s < nsub; ++s) {
    view[s] = *a_block;               // rows [sub_start[s], sub_start[s+1]) without copying: row_ptr entries stay
    view[s].ptr = a_block->ptr + sub_start[s];  // absolute positions into the block's col_idx / val
    view[s].rows = sub_start[s + 1] - sub_start[s];
    view[s].owning = false;
    const int st = spgemm_symbolic_dev(h, &view[s], b, &pend[s]);
    if (st != SPAM_OK) { drop(); return st; }
    sub_nnz[s] = spgemm_pending_nnz(pend[s]);
    my_nnz += sub_nnz[s];
    my_flops += h->stats.flops;
    for (int i = 0; i < 16; ++i) { acc.sym_bin_rows[i] += h->stats.sym_bin_rows[i]; acc.num_bin_rows[i] += h->stats.num_bin_rows[i]; }
    acc.kernel_launches += h->stats.kernel_launches;
  }
  const u64 launches_mark = h->stats.kernel_launches;  // counted so far in the last sub-block's stats
  // ---- everyone's nnz and row counts ----
  u64 mine[2] = {my_nnz, m};
  { const int st = small_allgather(h, c, mine, sizeof(mine)); if (st != SPAM_OK) { drop(); return st; } }
  u64 nnz_of[MAXR], rows_of[MAXR], nnz_before = 0, rows_before = 0, total_nnz = 0, rows_sum = 0;
  {
    const u64* all = reinterpret_cast<const u64*>(c->h_x + MAXR * 256);
    for (int r = 0; r < c->world; ++r) {
      nnz_of[r] = all[2 * r]; rows_of[r] = all[2 * r + 1];
      if (r < c->rank) { nnz_before += nnz_of[r]; rows_before += rows_of[r]; }
      total_nnz += nnz_of[r]; rows_sum += rows_of[r];
    }
  }
  if (rows_sum != total_rows || rows_before != row_start) { drop(); return spam_fail(h, SPAM_EINVAL, "the ranks' row blocks do not tile the matrix in rank order"); }
  // mode < 0: pick by measurement — with one peer the copy engines leave every SM to the product (R-MAT 22 on 2 GPUs:
  // 61 ms against 73 ms with the push kernel); with more peers their per-destination copies run one after another and
  // the push kernel, which stores every 16 bytes to all peers at once, wins (8 GPUs: 61 ms against 82 ms)
  if (mode < 0) mode = c->world <= 2 ? 2 : 0;
  const bool want_peers = mode == 0 || mode == 2;
  const bool dma = mode == 2;
  int st = ensure_buf(h, c, 0, (total_rows + 1) * 8, want_peers);
  if (st == SPAM_OK) st = ensure_buf(h, c, 1, (total_nnz ? total_nnz : 1) * 4, want_peers);
  if (st == SPAM_OK) st = ensure_buf(h, c, 2, (total_nnz ? total_nnz : 1) * es, want_peers);
  if (st != SPAM_OK) { drop(); return st; }
  const bool peers = want_peers && c->peer_ok && c->world > 1;
  u64* g_ptr = (u64*)c->buf[0].local;
  u32* g_idx = (u32*)c->buf[1].local;
  void* g_val = c->buf[2].local;
  // Every rank has passed the exchanges above, so nobody still reads the previous product out of these buffers
  // (the contract of the returned view); the push stream may start writing into the peers.
  // ---- numeric per sub-block into the final position, push behind it ----
  u64 off = nnz_before;
  for (int s = 0; s < nsub; ++s) {
    const u64 r0 = row_start + sub_start[s], nr = view[s].rows;
    // offset-fixed row_ptr of the sub-block, written where the assembled row_ptr wants it (the last rank's last
    // sub-block also writes the final entry)
    const bool last = (s == nsub - 1) && (c->rank == c->world - 1);
    st = add_offset_u64(h, g_ptr + r0, nr + 1, off, spgemm_pending_cptr(pend[s]));
    SpgemmPending* p = pend[s];
    pend[s] = nullptr;
    if (st == SPAM_OK) st = spgemm_numeric_into(h, p, g_ptr + r0, g_idx, g_val); else spgemm_pending_free(h, p);
    if (st != SPAM_OK) { drop(); return st; }
    if (peers) {
      CK(cudaEventRecord(c->ev_ready, h->stream));
      CK(cudaStreamWaitEvent(c->push, c->ev_ready, 0));
      // entries r0 .. r0+nr-1 of row_ptr (+ the closing entry on the very last sub-block): entry r0+nr belongs
      // to the next sub-block / rank, which writes the same value
      st = push_range(h, c, 0, r0 * 8, (nr + (last ? 1 : 0)) * 8, dma);
      if (st == SPAM_OK) st = push_range(h, c, 1, off * 4, sub_nnz[s] * 4, dma);
      if (st == SPAM_OK) st = push_range(h, c, 2, off * es, sub_nnz[s] * es, dma);
      if (st != SPAM_OK) { drop(); return st; }
    }
    off += sub_nnz[s];
  }
  if (peers) {
    CK(cudaEventRecord(c->ev_pushed, c->push));
    CK(cudaStreamWaitEvent(h->stream, c->ev_pushed, 0));
    CKS(barrier(h, c));  // completes once every rank's pushes are done: the whole C is here
  } else if (c->world > 1) {
    u64 bo[MAXR + 1];
    // row_ptr: rank r owns entries [rows_before_r, rows_before_r + rows_r) and the last rank the closing entry
    u64 rb = 0;
    for (int r = 0; r < c->world; ++r) { bo[r] = rb * 8; rb += rows_of[r]; }
    bo[c->world] = (total_rows + 1) * 8;
    CKS(spam_comm_allgatherv(h, g_ptr, bo));
    u64 nb2 = 0;
    for (int r = 0; r < c->world; ++r) { bo[r] = nb2 * 4; nb2 += nnz_of[r]; }
    bo[c->world] = nb2 * 4;
    CKS(spam_comm_allgatherv(h, g_idx, bo));
    for (int r = 0; r <= c->world; ++r) bo[r] = bo[r] / 4 * es;
    CKS(spam_comm_allgatherv(h, g_val, bo));
  }
  spam_dcsr* out = new spam_dcsr();
  out->dtype = a_block->dtype; out->rows = total_rows; out->cols = b->cols; out->nnz = total_nnz;
  out->ptr = g_ptr; out->idx = g_idx; out->val = g_val; out->owning = false; out->rows_sorted = -1;
  h->stats.flops = my_flops; h->stats.nnz_c = my_nnz;
  h->stats.kernel_launches = acc.kernel_launches + (h->stats.kernel_launches - launches_mark);
  for (int i = 0; i < 16; ++i) { h->stats.sym_bin_rows[i] = acc.sym_bin_rows[i]; h->stats.num_bin_rows[i] = acc.num_bin_rows[i]; }
  *cout = out;
  return SPAM_OK;
}

// y = A x with A row-sharded: every rank computes its rows into d_y[row_start ...) of a full-length y and the
// pieces are exchanged (all-gather-v as grouped broadcasts; y is small next to A).  rows_of: every rank's row
// count (host, world entries).
int spam_spmv_gathered(spam_handle* h, const spam_dcsr* a_block, const void* d_x, void* d_y_full,
                       const uint64_t* rows_of) {
  if (!h || !a_block || !d_x || !d_y_full || !rows_of) return spam_fail(h, SPAM_EINVAL, "null argument");
  CommState* c = h->comm;
  if (!c) return spam_fail(h, SPAM_ESTATE, "spam_comm_init first");
  CK(cudaSetDevice(h->device));
  const size_t es = dtype_size(a_block->dtype);
  u64 bo[MAXR + 1];
  u64 rb = 0;
  for (int r = 0; r < c->world; ++r) { bo[r] = rb * es; rb += rows_of[r]; }
  bo[c->world] = rb * es;
  if (rows_of[c->rank] != a_block->rows) return spam_fail(h, SPAM_EINVAL, "rows_of[rank] != rows of the block");
  CKS(spmv_dev(h, a_block, d_x, (unsigned char*)d_y_full + bo[c->rank]));
  if (c->world > 1) CKS(spam_comm_allgatherv(h, d_y_full, bo));
  return SPAM_OK;
}

// DOK -> CSR with the triplet stream spread over the ranks: rank r holds the r-th contiguous piece of the stream
// (n_local triplets on the device).  The rows are range-partitioned — rank r owns rows [r * per, (r + 1) * per),
// per = ceil(rows / world) — every rank groups its piece by owner (stable), one grouped ncclSend / ncclRecv
// all-to-all moves the groups, and each rank builds its block with the single-GPU routine.  Received groups are
// laid out in source-rank order, so the concatenation is the global stream order restricted to the rank's rows:
// "last write wins" holds across ranks.  *out_block: rows of the rank's range (row indices rebased to 0);
// *row_start: first global row of the block.
int spam_dok_to_csr_sharded(spam_handle* h, int dtype, uint64_t rows, uint64_t cols, uint64_t n_local,
                            const void* d_tri_rows, const void* d_tri_cols, const void* d_tri_vals,
                            uint64_t* row_start, spam_dcsr** out_block) {
  if (!h || !out_block || !row_start || (n_local && (!d_tri_rows || !d_tri_cols || !d_tri_vals))) return spam_fail(h, SPAM_EINVAL, "bad argument");
  *out_block = nullptr;
  CommState* c = h->comm;
  if (!c) return spam_fail(h, SPAM_ESTATE, "spam_comm_init first");
  if (rows == 0 || rows >= 0xFFFFFFFFull || cols >= 0xFFFFFFFFull) return spam_fail(h, SPAM_ECOLS, "dimension >= 2^32-1 (or zero rows)");
  if (c->world > 32) return spam_fail(h, SPAM_EINVAL, "at most 32 ranks");
  CK(cudaSetDevice(h->device));
  const size_t es = dtype_size(dtype);
  const u64 per = (rows + c->world - 1) / c->world;
  const u64 r0 = std::min<u64>(rows, per * c->rank), r1 = std::min<u64>(rows, r0 + per);
  DevGuard g(h);
  u64 *s_r = nullptr, *s_c = nullptr, *q_r = nullptr, *q_c = nullptr;
  void *s_v = nullptr, *q_v = nullptr;
  CKS(g.alloc(&s_r, n_local ? n_local : 1));
  CKS(g.alloc(&s_c, n_local ? n_local : 1));
  CKS(g.alloc_bytes(&s_v, (n_local ? n_local : 1) * es));
  u64 cnt_to[MAXR * 2] = {};
  int st = dok_partition_dev(h, dtype, rows, cols, per, c->world, n_local, (const u64*)d_tri_rows, (const u64*)d_tri_cols,
                             d_tri_vals, s_r, s_c, s_v, cnt_to);
  // every rank must reach the collective below, error or not: a failed rank sends nothing and reports afterwards
  u64 mine[MAXR + 1] = {};
  for (int d = 0; d < c->world; ++d) mine[d] = st == SPAM_OK ? cnt_to[d] : 0;
  mine[c->world] = st == SPAM_OK ? 0 : 1;
  static_assert((MAXR + 1) * 8 <= 256, "counts row does not fit the staging slot");
  CKS(small_allgather(h, c, mine, (size_t)(c->world + 1) * 8));
  const u64* all = reinterpret_cast<const u64*>(c->h_x + MAXR * 256);
  u64 recv_from[MAXR] = {}, total = 0, failed = 0;
  for (int s = 0; s < c->world; ++s) {
    recv_from[s] = all[(size_t)s * (c->world + 1) + c->rank];
    failed |= all[(size_t)s * (c->world + 1) + c->world];
    total += recv_from[s];
  }
  if (st != SPAM_OK) return st;
  if (failed) return spam_fail(h, SPAM_EINDEX, "another rank's slice holds a triplet index out of range");
  if (total >= 0xFFFFFFFFull) return spam_fail(h, SPAM_EOVERFLOW, "more than 2^32-1 triplets for one rank");
  CKS(g.alloc(&q_r, total ? total : 1));
  CKS(g.alloc(&q_c, total ? total : 1));
  CKS(g.alloc_bytes(&q_v, (total ? total : 1) * es));
  u64 soff[MAXR + 1] = {}, roff[MAXR + 1] = {};
  for (int d = 0; d < c->world; ++d) { soff[d + 1] = soff[d] + cnt_to[d]; roff[d + 1] = roff[d] + recv_from[d]; }
  NCK(g_nccl.GroupStart());
  ncclResult_t rr = ncclSuccess;
  for (int p = 0; p < c->world && rr == ncclSuccess; ++p) {
    if (p == c->rank) continue;
    if (cnt_to[p]) {
      rr = g_nccl.Send(s_r + soff[p], cnt_to[p] * 8, ncclUint8, p, c->comm, h->stream);
      if (rr == ncclSuccess) rr = g_nccl.Send(s_c + soff[p], cnt_to[p] * 8, ncclUint8, p, c->comm, h->stream);
      if (rr == ncclSuccess) rr = g_nccl.Send((const char*)s_v + soff[p] * es, cnt_to[p] * es, ncclUint8, p, c->comm, h->stream);
    }
    if (recv_from[p] && rr == ncclSuccess) {
      rr = g_nccl.Recv(q_r + roff[p], recv_from[p] * 8, ncclUint8, p, c->comm, h->stream);
      if (rr == ncclSuccess) rr = g_nccl.Recv(q_c + roff[p], recv_from[p] * 8, ncclUint8, p, c->comm, h->stream);
      if (rr == ncclSuccess) rr = g_nccl.Recv((char*)q_v + roff[p] * es, recv_from[p] * es, ncclUint8, p, c->comm, h->stream);
    }
  }
  {
    const ncclResult_t ge = g_nccl.GroupEnd();
    if (rr != ncclSuccess || ge != ncclSuccess) return spam_fail(h, SPAM_ECUDA, g_nccl.GetErrorString(rr != ncclSuccess ? rr : ge));
  }
  const u64 self = cnt_to[c->rank];
  if (self) {
    CK(cudaMemcpyAsync(q_r + roff[c->rank], s_r + soff[c->rank], self * 8, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(q_c + roff[c->rank], s_c + soff[c->rank], self * 8, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync((char*)q_v + roff[c->rank] * es, (const char*)s_v + soff[c->rank] * es, self * es, cudaMemcpyDeviceToDevice, h->stream));
  }
  *row_start = r0;
  const u64 my_rows = r1 - r0;
  if (my_rows == 0) {  // more ranks than rows: an empty 0-row block is not representable (rows are NonZero): report 1 empty row
    return dok_to_csr_dev(h, dtype, 1, cols, 0, q_r, q_c, q_v, out_block);
  }
  return dok_to_csr_dev(h, dtype, my_rows, cols, total, q_r, q_c, q_v, out_block);
}

}  // extern "C"
