/* spam_cuda.h — C ABI of libspam_cuda.so: the B200 (sm_100a) drop-in for the hot path of
 * sledgehammervampire/sparse_matrix.
 *
 * What each entry point replaces (paths relative to the reference repo):
 *   spam_spgemm_symbolic / spam_spgemm_numeric   body of CsrMatrix::mul_hash
 *       spam_csr/src/mul_hash.rs:13-36  (rows_to_threads :38-64, mul_hash_symbolic :66-103,
 *       mul_hash_numeric :105-201), reached from `impl Mul for &CsrMatrix`
 *       spam_csr/src/lib.rs:292-297.  Two-phase because the reference allocates the result
 *       Vecs with exact capacity after the symbolic scan (mul_hash.rs:117-119): the caller
 *       (Rust) allocates, the library fills.
 *   spam_dok_to_csr / spam_dok_to_csr_fetch      `impl From<DokMatrix<T>> for CsrMatrix<T,true>`
 *       spam_csr/src/lib.rs:315-334 fed by DokMatrix::set_element spam_dok/src/lib.rs:167-176.
 *   spam_spmv                                    NEW API (the reference has no SpMV);
 *       semantics = a.mul_hash::<_,true>(&x) with x an n x 1 CsrMatrix.
 *   spam_rows_to_parts                           rows_to_threads partition formula
 *       spam_csr/src/mul_hash.rs:51-62 with tnum = number of GPUs.
 *
 * Conventions
 *   - Host-side indices are uint64_t (Rust usize on 64-bit; CsrMatrix fields
 *     spam_csr/src/lib.rs:25-32: vals, indices, offsets).  On the device col_idx is u32
 *     (the reference truncates keys to u32, mul_hash.rs:92,157), row_ptr u64.
 *   - Output rows are sorted by column unless the caller asks for the reference's unsorted order
 *     (`sorted = 0`); sorted rows are valid for both IS_SORTED variants (invariant6, lib.rs:69-77).
 *     Cancellation zeros are kept (mul_hash.rs:88-96).
 *   - Integer dtypes wrap (two's complement), floats: products mul-then-add, no FMA.
 *   - Every function returns a spam_status; nothing unwinds across the boundary.  The
 *     reference panics where we return an error (mul_hash.rs:47-48, lib.rs:270).
 *   - A handle owns one CUDA stream + workspace and is NOT thread-safe; use one handle
 *     per thread.  There is no CPU fallback: without a CUDA device spam_cuda_create fails.
 *   - Deviation from the reference: rows, cols and the number of entries of an operand must be
 *     < 2^32-1 (the reference only bounds the right-hand side's columns, mul_hash.rs:12): row ids
 *     are u32 on the device (row permutations, per-row counters).  nnz(C) and the product count
 *     are 64-bit.  Larger shapes return SPAM_ECOLS.
 *   - Operands are validated once per matrix on the device (row_ptr monotone from 0 to nnz,
 *     columns < cols: invariants 3, 4, 5, 7 of spam_csr/src/lib.rs:47-81): SPAM_EINVAL /
 *     SPAM_EINDEX where the reference would panic, never a device fault.
 */
#ifndef SPAM_CUDA_H
#define SPAM_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct spam_handle spam_handle;
typedef struct spam_dcsr spam_dcsr; /* device-resident CSR: u64 row_ptr[rows+1], u32 col_idx[nnz], T val[nnz] */

typedef enum spam_dtype { SPAM_F32 = 0, SPAM_F64 = 1, SPAM_I32 = 2, SPAM_I64 = 3 } spam_dtype;

typedef enum spam_status {
  SPAM_OK = 0,
  SPAM_EINVAL = 1,    /* null pointer / bad dtype / bad argument */
  SPAM_EDIM = 2,      /* A.cols != B.rows (the reference would panic out-of-bounds, mul_hash.rs:46) */
  SPAM_ECOLS = 3,     /* a dimension >= 2^32-1: u32::MAX is the empty-slot sentinel (mul_hash.rs:12, set.rs:110) */
  SPAM_ENOMEM = 4,    /* device or host allocation failed */
  SPAM_ECUDA = 5,     /* CUDA runtime error; see spam_last_error */
  SPAM_ESTATE = 6,    /* numeric/fetch called without a matching symbolic/build */
  SPAM_EOVERFLOW = 7, /* flop or nnz count overflowed (reference: checked_add().unwrap()) */
  SPAM_EINDEX = 8,    /* IndexError: a triplet or column index out of range (spam_matrix/src/lib.rs:13) */
  SPAM_EDTYPE = 9     /* operand dtypes differ */
} spam_status;

/* Per-call statistics of the last SpGEMM on this handle (device-side event timing). */
typedef struct spam_stats {
  uint64_t flops;        /* intermediate products P = sum over A entries of nnz(B row) (mul_hash.rs:44-48) */
  uint64_t nnz_c;
  uint64_t kernel_launches; /* number of kernels this library launched for the last call */
  uint64_t bytes_h2d, bytes_d2h;
  float ms_flop, ms_symbolic, ms_scan, ms_numeric, ms_total; /* CUDA-event times; 0 when timing disabled */
  /* rows per bin: 0 tiny, 1..8 hash bins (one per power of two), 9 heavy (global table), 10 merge */
  uint32_t sym_bin_rows[16]; /* symbolic pass, binned by intermediate products */
  uint32_t num_bin_rows[16]; /* numeric pass, binned by row nnz of C */
  /* How often the rarely taken code paths ran in the last call (tests assert that they are exercised):
   * [0] one-warp rows sorted by the shared-memory bitonic network (columns too wide to pack with an index),
   * [1] team rows whose bucket drain overflowed (compaction + block-wide bitonic network),
   * [2] global-table rows whose bucket drain overflowed (global-memory bitonic network),
   * [3] rows the bucket-sort (ESC) bins handed back to the global-table kernel,
   * [4] DOK->CSR / transpose path of the last build: 1 = counting sort by row (short segments),
   *     2 = LSD radix sort (some row or column holds more than 32 entries),
   * [5] 1 = the product ran as the one-pass merge kernel (every row a merge row by the cached statistics),
   * [6..7] reserved. */
  uint32_t fallbacks[8];
} spam_stats;

/* ---- lifecycle --------------------------------------------------------------------- */
int spam_cuda_create(spam_handle** h, int device);
int spam_cuda_destroy(spam_handle* h);
/* run on an existing CUDA stream (cudaStream_t as void*); NULL restores the handle's own stream */
int spam_cuda_set_stream(spam_handle* h, void* cuda_stream);
/* enable per-phase event timing into spam_stats (adds event records, no syncs beyond the API's own) */
int spam_cuda_set_timing(spam_handle* h, int enabled);
int spam_cuda_get_stats(spam_handle* h, spam_stats* out); /* syncs on the last product when timing is on */
/* Running totals of the CUDA-event phase times (flop, symbolic, scan, numeric, total; ms) over the products
 * completed since the last reset, and their count.  With timing enabled every product records its phase
 * events into one of two event sets; a set is read back after the next host synchronisation that covers it,
 * so a loop of products can be timed per phase without any extra synchronisation inside the loop. */
int spam_cuda_get_phase_totals(spam_handle* h, double* ms5 /* 5 entries */, uint64_t* products, int reset);
int spam_cuda_synchronize(spam_handle* h);
const char* spam_strerror(int status);
const char* spam_last_error(const spam_handle* h);
/* number of symbols below that the library exports, and the ABI version (tests check this) */
int spam_cuda_abi_version(void);

/* pinned host memory for callers that want full-rate PCIe copies */
int spam_host_alloc(void** p, uint64_t bytes);
int spam_host_free(void* p);

/* ---- host-buffer path (what the vendored spam_csr::mul_hash body calls) ---------------- */
/* Phase 1: uploads A and B (B may alias A: same pointers => uploaded once), runs flop count +
 * symbolic + scan.  Writes c_ptr[0..a_rows] (caller-owned, a_rows+1 entries) and *c_nnz. */
int spam_spgemm_symbolic(spam_handle* h, int dtype, uint64_t a_rows, uint64_t a_cols, const uint64_t* a_ptr,
                         const uint64_t* a_idx, const void* a_val, uint64_t b_rows, uint64_t b_cols,
                         const uint64_t* b_ptr, const uint64_t* b_idx, const void* b_val, uint64_t* c_ptr,
                         uint64_t* c_nnz);
/* Phase 2: numeric pass; fills caller-owned c_idx[c_nnz], c_val[c_nnz].  Releases the pending product.
 * sorted = 1: rows sorted by column — mul_hash::<_, true> (mul_hash.rs:164-175), and a valid CsrMatrix<T, false>
 *             as well (invariant6 only asks for distinct columns).
 * sorted = 0: the rows in the reference's B2 = false order, column for column: the slot order of linprobe's map
 *             (mul_hash.rs:176-186, map.rs:59-63) under the reference's insertion order.  Same entries and values
 *             as sorted = 1, permuted inside each row by one more pass (slotorder.cu). */
int spam_spgemm_numeric(spam_handle* h, uint64_t* c_idx, void* c_val, int sorted);

/* y = A x with dense x (a_cols) and y (a_rows); empty rows give 0. */
int spam_spmv(spam_handle* h, int dtype, uint64_t a_rows, uint64_t a_cols, const uint64_t* a_ptr,
              const uint64_t* a_idx, const void* a_val, const void* x, void* y);

/* Triplet stream (DokMatrix::set_element order: last write wins, writing zero deletes) -> sorted CSR.
 * Phase 1 writes c_ptr[rows+1] and *c_nnz; phase 2 fills c_idx/c_val. */
int spam_dok_to_csr(spam_handle* h, int dtype, uint64_t rows, uint64_t cols, uint64_t n_triplets,
                    const uint64_t* tri_rows, const uint64_t* tri_cols, const void* tri_vals, uint64_t* c_ptr,
                    uint64_t* c_nnz);
int spam_dok_to_csr_fetch(spam_handle* h, uint64_t* c_idx, void* c_val);

/* MatrixMarket coordinate text -> triplet stream (host code, no device needed).  Restates the grammar of
 * `parse_matrix_market`, spam_dok/src/lib.rs:282-478: integer -> kind SPAM_I64, real -> SPAM_F64; general or
 * symmetric (both (r, c) and (c, r) are emitted); 1-based indices; zeros skipped; entry lines are read until the
 * first line that does not match, the rest is ignored; the declared entry count is not used.  Feed the triplets
 * to spam_dok_to_csr (a later duplicate replaces an earlier one, as in the reference's BTreeMap::insert) to
 * get what `CsrMatrix::from(dok)` gives.  SPAM_EINVAL: grammar error; SPAM_EDTYPE: complex / pattern /
 * skew-symmetric / hermitian; SPAM_EDIM: a zero dimension (HasZeroDimension); SPAM_EINDEX: index 0.
 * tri_rows / tri_cols / tri_vals are malloc'ed: release with spam_mm_free. */
typedef struct spam_mm {
  int kind;                  /* SPAM_I64 or SPAM_F64 */
  uint64_t rows, cols;
  uint64_t declared_entries; /* third number of the size line */
  uint64_t n;                /* triplets produced */
  uint64_t* tri_rows;
  uint64_t* tri_cols;
  void* tri_vals;            /* n values of 8 bytes (int64_t or double) */
  char err[96];
} spam_mm;
int spam_mm_parse(const char* text, uint64_t len, spam_mm* out);
void spam_mm_free(spam_mm* m);

/* ---- device-resident path (benchmarks, multi-GPU, chained products) ---------------------- */
int spam_csr_upload(spam_handle* h, int dtype, uint64_t rows, uint64_t cols, uint64_t nnz, const uint64_t* ptr,
                    const uint64_t* idx, const void* val, spam_dcsr** out);
/* non-owning view over caller-managed device memory (e.g. torch tensors): u64 ptr, u32 idx, T val */
int spam_dcsr_wrap(spam_handle* h, int dtype, uint64_t rows, uint64_t cols, uint64_t nnz, const void* d_ptr_u64,
                   const void* d_idx_u32, const void* d_val, spam_dcsr** out);
int spam_dcsr_info(const spam_dcsr* m, int* dtype, uint64_t* rows, uint64_t* cols, uint64_t* nnz, void** d_ptr,
                   void** d_idx, void** d_val);
int spam_dcsr_download(spam_handle* h, const spam_dcsr* m, uint64_t* ptr, uint64_t* idx, void* val);
int spam_dcsr_free(spam_handle* h, spam_dcsr* m);
/* rows [r0, r1) of m as a new owning matrix with row_ptr rebased to 0 (the per-rank A block) */
int spam_dcsr_slice_rows(spam_handle* h, const spam_dcsr* m, uint64_t r0, uint64_t r1, spam_dcsr** out);

/* rows rows[0..n) of m (host array of row indices, any order, repeats allowed) as a new owning matrix:
 * how a caller pulls a sample of rows of a large device-resident product back to the host */
int spam_dcsr_select_rows(spam_handle* h, const spam_dcsr* m, const uint64_t* rows, uint64_t n, spam_dcsr** out);

/* Transpose.  Replaces `Matrix::transpose` of CsrMatrix, spam_csr/src/lib.rs:256-264 (an O(rows*cols)
 * loop of set_element calls in the reference; for matrices without explicit zeros the same result as
 * DokMatrix::transpose, spam_dok/src/lib.rs:178-188, followed by From<DokMatrix>): every stored entry
 * (i, j, v) — explicit zeros included, CsrMatrix::set_element stores them — becomes (j, i, v); rows of
 * the result are sorted by column whatever the order inside the input's rows.  Device: *out is a new
 * owning matrix.  Host: t_ptr has cols+1 entries, t_idx / t_val have nnz = ptr[rows] entries (the caller
 * knows all sizes up front, so one phase). */
int spam_dcsr_transpose(spam_handle* h, const spam_dcsr* m, spam_dcsr** out);
int spam_csr_transpose(spam_handle* h, int dtype, uint64_t rows, uint64_t cols, const uint64_t* ptr,
                       const uint64_t* idx, const void* val, uint64_t* t_ptr, uint64_t* t_idx, void* t_val);

/* Elementwise C = A + B (op 0) or A - B (op 1).  Replaces `impl Add / Sub for CsrMatrix` ->
 * `apply_elementwise`, spam_csr/src/lib.rs:83-149, 276-290: per row the union of the two column sets; a column
 * in both rows gives f(t1, t2), in one row f(t, 0) or f(0, t); nothing is filtered (cancellation zeros stay);
 * rows of the result are sorted by column (the reference's unsorted branch iterates a std HashMap, whose
 * order is unspecified).  `op | 2` applies the IS_SORTED = false branch's rule for entries only in the left
 * operand (kept as they are instead of f(t, 0), lib.rs:119-137; differs only for t = -0.0 under add).
 * SPAM_EDIM when the shapes differ (the reference asserts, lib.rs:87-91).
 * Host path in two phases like DOK -> CSR: phase 1 writes c_ptr[rows+1] and *c_nnz, phase 2 fills. */
int spam_dcsr_ewise(spam_handle* h, int op, const spam_dcsr* a, const spam_dcsr* b, spam_dcsr** out);
int spam_csr_ewise(spam_handle* h, int op, int dtype, uint64_t rows, uint64_t cols, const uint64_t* a_ptr,
                   const uint64_t* a_idx, const void* a_val, const uint64_t* b_ptr, const uint64_t* b_idx,
                   const void* b_val, uint64_t* c_ptr, uint64_t* c_nnz);
int spam_csr_ewise_fetch(spam_handle* h, uint64_t* c_idx, void* c_val);

/* C = A * B, all on the device; *c is a new owning matrix (stream-ordered allocation). */
int spam_spgemm_dev(spam_handle* h, const spam_dcsr* a, const spam_dcsr* b, spam_dcsr** c);
/* the same with the B2 const generic as an argument: sorted = 0 gives the reference's unsorted (slot) order */
int spam_spgemm_dev_b2(spam_handle* h, const spam_dcsr* a, const spam_dcsr* b, int sorted, spam_dcsr** c);
int spam_spmv_dev(spam_handle* h, const spam_dcsr* a, const void* d_x, void* d_y);
/* device triplets (u64 rows, u64 cols, T vals) -> device CSR */
int spam_dok_to_csr_dev(spam_handle* h, int dtype, uint64_t rows, uint64_t cols, uint64_t n_triplets,
                        const void* d_tri_rows, const void* d_tri_cols, const void* d_tri_vals, spam_dcsr** out);

/* Flop-balanced contiguous row blocks: row_starts[0]=0, row_starts[parts]=a.rows,
 * row_starts[t] = partition_point(ps <= ceil(total/parts)*t) - 1   (mul_hash.rs:51-62).
 * Also returns the total intermediate-product count. */
int spam_rows_to_parts(spam_handle* h, const spam_dcsr* a, const spam_dcsr* b, uint32_t parts,
                       uint64_t* row_starts /* parts+1, host */, uint64_t* total_flops);

/* Same formula on a per-row device-time estimate instead of the raw product count:
 * cost_i = flop_i * w(flop_i), w = measured time per product of the kernel that rows of that size take
 * on B200 (team and global-table rows cost 2-4x more per product than one-warp rows).  The reference
 * balances CPU threads whose cost per product is flat; across GPUs a power-law matrix puts all the
 * heavy rows into the first block (R-MAT 22 over 8 GPUs: balance 0.78 by flops).
 * Equivalent to spam_rows_to_parts (up to rounding) when every row is in the same size class (stencils). */
int spam_rows_to_parts_cost(spam_handle* h, const spam_dcsr* a, const spam_dcsr* b, uint32_t parts,
                            uint64_t* row_starts /* parts+1, host */, uint64_t* total_flops);

/* add `offset` to every entry of a device u64 array (offset-fixing a rank's row_ptr shard before the
 * all-gather-v); n entries */
int spam_offset_u64(spam_handle* h, void* d_ptr_u64, uint64_t n, uint64_t offset);

/* ---- several GPUs of one node, one process (and one handle) per GPU ----------------------------
 * The reference's threads write disjoint slices of ONE output (mul_hash.rs:121-128); so do the ranks here.
 * NCCL (bound with dlopen: no link-time dependency) carries the rendezvous and the small exchanges; the shards
 * of C are stored by the ranks into each other's peer-mapped copy of the whole C (cudaIpc*), overlapped with the
 * product.  All spam_comm_* / *_gathered calls are collective: every rank calls them in the same order. */
/* rank 0 makes a 128-byte id and hands it to the other ranks by any host channel (ncclGetUniqueId) */
int spam_comm_unique_id(void* out128);
int spam_comm_init(spam_handle* h, const void* id128, int rank, int world);
int spam_comm_destroy(spam_handle* h);
int spam_comm_info(const spam_handle* h, int* rank, int* world, int* peer_mapped);
/* replicate a device buffer from `root` (B, and A before it is sliced): ncclBroadcast on the handle's stream */
int spam_comm_broadcast(spam_handle* h, void* d_buf, uint64_t bytes, int root);
/* n <= 32 values per rank -> all[world * n] on the host (per-rank nnz / row counts) */
int spam_comm_allgather_u64(spam_handle* h, const uint64_t* mine, uint32_t n, uint64_t* all);
/* all-gather-v, in place: rank r's bytes sit at d_out + byte_offsets[r] (byte_offsets: world + 1 entries, host);
 * NCCL has no ncclAllGatherv: one ncclBroadcast per source rank inside one NCCL group */
int spam_comm_allgatherv(spam_handle* h, void* d_out, const uint64_t* byte_offsets);
/* C = A * B, A row-sharded (a_block = this rank's rows [row_start, row_start + rows) with row_ptr rebased to 0;
 * the blocks tile A in rank order), B replicated; on return every rank holds the WHOLE C with its offset-fixed
 * row_ptr.  *c is a non-owning view of the handle's gather buffers: valid until the next gathered product on
 * this handle; release the view with spam_dcsr_free.  nsub (1..16) sub-blocks of rows pipeline the numeric
 * kernels with the exchange.  mode 0: peer stores over NVLink by a push kernel (falls back to 1 if the peers'
 * buffers cannot be mapped), mode 1: grouped ncclBroadcast, mode 2: like 0 with the copy engines doing the
 * peer-to-peer copies, mode -1: the library picks 0 or 2 by the number of ranks.  Phase timing
 * (spam_cuda_set_timing) is not collected here. */
int spam_spgemm_gathered(spam_handle* h, const spam_dcsr* a_block, const spam_dcsr* b, uint64_t row_start,
                         uint64_t total_rows, int nsub, int mode, spam_dcsr** c);
/* y = A x, A row-sharded, x replicated: each rank fills its rows of d_y_full (total rows entries) and the pieces
 * are exchanged; rows_of = every rank's row count (host, world entries) */
int spam_spmv_gathered(spam_handle* h, const spam_dcsr* a_block, const void* d_x, void* d_y_full,
                       const uint64_t* rows_of);

/* DOK -> CSR with the triplet stream spread over the ranks (rank r holds the r-th contiguous piece, on the device):
 * rows are range-partitioned (ceil(rows / world) per rank), the pieces are grouped by owner and exchanged with one
 * grouped ncclSend / ncclRecv all-to-all, every rank builds its row block with the single-GPU routine.  Last write
 * wins across ranks (the pieces arrive in stream order).  *row_start = first global row of *out_block. */
int spam_dok_to_csr_sharded(spam_handle* h, int dtype, uint64_t rows, uint64_t cols, uint64_t n_local,
                            const void* d_tri_rows, const void* d_tri_cols, const void* d_tri_vals,
                            uint64_t* row_start, spam_dcsr** out_block);

#ifdef __cplusplus
}
#endif
#endif /* SPAM_CUDA_H */
