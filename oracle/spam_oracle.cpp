// spam_oracle.cpp — CPU restatement of the reference's SpGEMM hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under sparse_matrix_b200/ (the product) may
// link, load or call this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, and there only as the checker or as the
// timed CPU baseline.
//
// PARITY STATUS: "parity unpinned" in the strict sense.  The reference
// (sledgehammervampire/sparse_matrix) is nightly-Rust; no Rust toolchain exists in the
// build container, so the reference itself can neither be compiled nor run here, and the
// reference ships no golden vectors / fixtures for this path.  What pins
// this restatement instead:
//   * the reference's own property tests, re-created against this file in
//     tests/test_oracle.py (spam_csr/src/tests.rs:356-371 dense-DOK equivalence on
//     Wrapping<i8>; spam_csr/src/mul_hash.rs:204-224 partition shape;
//     fuzz/fuzz_targets/mul_hash.rs:29,40-45 invariants + Higham bound);
//   * hand-derived known-answer vectors for the linprobe hash / slot order
//     (tests/golden/*.json, derived by hand from linprobe/src/lib.rs:13,
//     set.rs:36-160, map.rs:31-121);
//   * scipy.sparse as an independent structural cross-check.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference).  Compile with -ffp-contract=off: the reference multiplies, then
// adds (no FMA), sequentially in A-row storage order (mul_hash.rs:145-162).
//
// Build: see oracle/Makefile  (g++ -O3 -march=native -ffp-contract=off -shared -fPIC)

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

namespace {

using u32 = uint32_t;
using u64 = uint64_t;

// ---------------------------------------------------------------------------------
// linprobe  (linprobe/src/lib.rs:13-32)
// ---------------------------------------------------------------------------------
constexpr u32 HASH_SCAL = 107;        // lib.rs:13
constexpr u64 MIN_TABLE_SIZE = 16;    // lib.rs:14
constexpr u32 EMPTY = 0xFFFFFFFFu;    // set.rs:45 / set.rs:110 sentinel u32::MAX

// MulHasher::write_u32 + finish (lib.rs:20-31): hash = key.wrapping_mul(107) as u64
inline u64 mul_hash_of(u32 key) { return (u64)(u32)(key * HASH_SCAL); }

inline u64 next_pow2(u64 x) {  // usize::checked_next_power_of_two (0 -> 1)
  if (x <= 1) return 1;
  u64 p = 1;
  while (p < x) p <<= 1;
  return p;
}

// capacity rule shared by set.rs:38-43,56-63 and map.rs:33-38,50-55
inline u64 table_size_for(u64 capacity) { return std::max<u64>(2 * next_pow2(capacity), MIN_TABLE_SIZE); }

// set.rs:128-160 insert_raw
inline bool set_insert_raw(u32* slots, u64 len, u32 key, u64 hash) {
  u64 index = hash & (len - 1);
  for (;;) {
    u32& curr = slots[index];
    if (curr == key) return false;
    if (curr == EMPTY) { curr = key; return true; }
    index = (index + 1) & (len - 1);
  }
}

// linprobe::HashSet (set.rs:11-166)
struct HashSet {
  std::vector<u32> slots;
  u64 upper_bound;
  u64 items;
  // set.rs:24-26 new() == with_capacity(MIN_TABLE_SIZE / 4)
  HashSet() : HashSet(MIN_TABLE_SIZE / 4) {}
  // set.rs:37-54
  explicit HashSet(u64 capacity) : upper_bound(table_size_for(capacity)), items(0) {
    slots.assign(upper_bound, EMPTY);
  }
  // set.rs:55-64 — only ever lowers the bound
  void shrink_to(u64 capacity) { upper_bound = std::min(table_size_for(capacity), upper_bound); }
  u64 len() const { return items; }
  // set.rs:71-74
  void clear() { std::fill(slots.begin(), slots.begin() + upper_bound, EMPTY); items = 0; }
  // set.rs:77-107
  void grow() {
    if (upper_bound == slots.size()) slots.resize(slots.size() * 2, EMPTY);
    std::vector<u32> keys;
    keys.reserve(items);
    for (u64 i = 0; i < upper_bound; ++i)
      if (slots[i] != EMPTY) { keys.push_back(slots[i]); slots[i] = EMPTY; }
    upper_bound *= 2;
    for (u32 key : keys) set_insert_raw(slots.data(), upper_bound, key, mul_hash_of(key));
  }
  // set.rs:109-124
  void insert(u32 key) {
    if (set_insert_raw(slots.data(), upper_bound, key, mul_hash_of(key))) items += 1;
    if (items > upper_bound / 2) grow();
  }
};

// linprobe::HashMap<u32, V> (map.rs:9-121).  Slot = Option<(NonZeroU8, K, V)>; the
// NonZeroU8 is only a niche for the Option, restated as an `occupied` byte.
template <class V>
struct HashMap {
  struct Slot { unsigned char occupied; u32 key; V val; };
  std::vector<Slot> slots;
  u64 capacity;
  // map.rs:31-48
  explicit HashMap(u64 cap) : capacity(table_size_for(cap)) { slots.assign(capacity, Slot{0, 0, V()}); }
  // map.rs:49-58
  void shrink_to(u64 cap) { capacity = table_size_for(cap); }
  // map.rs:66-97 entry() + map.rs:105-120 and_modify / or_insert, fused as the one call
  // site uses them (mul_hash.rs:157-161): modify-add if present, else insert.
  template <class AddAssign>
  void upsert(u32 key, V v, AddAssign add_assign) {
    u64 index = mul_hash_of(key) & (capacity - 1);
    for (;;) {
      Slot& s = slots[index];
      if (s.occupied) {
        if (s.key == key) { add_assign(s.val, v); return; }
        index = (index + 1) & (capacity - 1);
      } else {
        s.occupied = 1; s.key = key; s.val = v;  // first product stored, not added to 0
        return;
      }
    }
  }
  // map.rs:59-63 drain(): slots[..capacity] in index order, taking each
  template <class F>
  void drain(F f) {
    for (u64 i = 0; i < capacity; ++i)
      if (slots[i].occupied) { f(slots[i].key, slots[i].val); slots[i].occupied = 0; }
  }
};

// ---------------------------------------------------------------------------------
// element arithmetic: Rust `t * t1` and `*t += t1` for T in {f32,f64} (IEEE, unfused)
// and integers in release mode / Wrapping<iN> (two's complement wrap).
// ---------------------------------------------------------------------------------
template <class T> struct Arith {
  static T mul(T a, T b) { return a * b; }
  static void add_assign(T& a, T b) { a += b; }
};
template <class I, class U> struct WrapArith {
  static I mul(I a, I b) { return (I)(U)((U)a * (U)b); }
  static void add_assign(I& a, I b) { a = (I)(U)((U)a + (U)b); }
};
template <> struct Arith<int8_t> : WrapArith<int8_t, uint8_t> {};
template <> struct Arith<int32_t> : WrapArith<int32_t, uint32_t> {};
template <> struct Arith<int64_t> : WrapArith<int64_t, uint64_t> {};

// CSR view with the reference's field names (spam_csr/src/lib.rs:25-32)
template <class T>
struct CsrView {
  u64 rows, cols;
  const T* vals;
  const u64* indices;
  const u64* offsets;
};

// spam_csr/src/lib.rs:267-274 checked_inclusive_scan: [0, v0, v0+v1, ...] (len+1).
// Returns false on overflow (the reference panics via checked_add().unwrap()).
bool checked_inclusive_scan(const std::vector<u64>& v, std::vector<u64>& out) {
  out.resize(v.size() + 1);
  out[0] = 0;
  u64 sum = 0;
  for (size_t i = 0; i < v.size(); ++i) {
    if (__builtin_add_overflow(sum, v[i], &sum)) return false;
    out[i + 1] = sum;
  }
  return true;
}

template <class F>
void parallel_blocks(const std::vector<u64>& rows_offset, F body) {
  // rayon::scope with one spawned task per (tlo, thi) window (mul_hash.rs:72-77,120-131)
  size_t nblk = rows_offset.size() - 1;
  if (nblk == 1) { body(0, rows_offset[0], rows_offset[1]); return; }
  std::vector<std::thread> th;
  th.reserve(nblk);
  for (size_t t = 0; t < nblk; ++t) th.emplace_back([&, t] { body(t, rows_offset[t], rows_offset[t + 1]); });
  for (auto& x : th) x.join();
}

// mul_hash.rs:38-64 rows_to_threads.  Returns false where the reference would panic
// (checked_add overflow, :47-48).
template <class T>
bool rows_to_threads(const CsrView<T>& a, const CsrView<T>& b, u64 tnum, std::vector<u64>& row_nz,
                     std::vector<u64>& rows_offset) {
  row_nz.assign(a.rows, 0);
  bool ok = true;
  // :39-50 (par_iter over rows; order-independent, so a plain partitioned loop)
  {
    u64 nt = std::max<u64>(1, std::min<u64>(tnum, a.rows));
    std::vector<u64> split(nt + 1);
    for (u64 t = 0; t <= nt; ++t) split[t] = a.rows * t / nt;
    std::vector<char> okv(nt, 1);
    parallel_blocks(split, [&](size_t t, u64 lo, u64 hi) {
      for (u64 i = lo; i < hi; ++i) {
        u64 sum = 0;
        for (u64 e = a.offsets[i]; e < a.offsets[i + 1]; ++e) {
          u64 k = a.indices[e];
          if (__builtin_add_overflow(sum, b.offsets[k + 1] - b.offsets[k], &sum)) okv[t] = 0;
        }
        row_nz[i] = sum;
      }
    });
    for (char c : okv) ok = ok && c;
  }
  std::vector<u64> ps;  // :51
  if (!checked_inclusive_scan(row_nz, ps)) return false;
  u64 total = ps.back();                     // :52
  u64 avg = (total + tnum - 1) / tnum;       // :55 unstable_div_ceil
  rows_offset.assign(1, 0);                  // :56
  for (u64 tid = 1; tid < tnum; ++tid) {     // :57-61
    // partition_point(|x| x <= avg*tid) - 1 ; ps is sorted, ps[0]=0 so the point is >= 1
    u64 bound = avg * tid;
    u64 pp = std::upper_bound(ps.begin(), ps.end(), bound) - ps.begin();
    rows_offset.push_back(pp - 1);
  }
  rows_offset.push_back(a.rows);             // :62
  return ok;
}

// mul_hash.rs:66-103 mul_hash_symbolic — overwrites row_nz (flop -> nnz) in place
template <class T>
void mul_hash_symbolic(const CsrView<T>& a, const CsrView<T>& b, std::vector<u64>& row_nz,
                       const std::vector<u64>& rows_offset) {
  parallel_blocks(rows_offset, [&](size_t, u64 tlo, u64 thi) {
    HashSet hs;                                              // :77
    for (u64 i = tlo; i < thi; ++i) {
      if (row_nz[i] == 0) continue;                          // :84-86
      hs.shrink_to(row_nz[i]);                               // :87
      for (u64 e = a.offsets[i]; e < a.offsets[i + 1]; ++e) {  // :88
        u64 k = a.indices[e];
        for (u64 j = b.offsets[k]; j < b.offsets[k + 1]; ++j)  // :89-90
          hs.insert((u32)b.indices[j]);                      // :92  `j as u32`
      }
      row_nz[i] = hs.len();                                  // :95
      hs.clear();                                            // :96
    }
  });
}

// mul_hash.rs:105-201 mul_hash_numeric::<B1, B2>.  `sorted` is B2.
// Optionally also accumulates sum|a*b| per output entry (tolerance denominator for the
// GPU parity tests; not part of the reference) when abs_out != nullptr.
template <class T>
bool mul_hash_numeric(const CsrView<T>& a, const CsrView<T>& b, const std::vector<u64>& row_nz,
                      const std::vector<u64>& rows_offset, bool sorted, std::vector<u64>& offsets,
                      u64*& indices, T*& vals) {
  if (!checked_inclusive_scan(row_nz, offsets)) return false;  // :117
  u64 nnz = offsets.back();                                    // :118
  indices = (u64*)std::malloc(std::max<u64>(1, nnz) * sizeof(u64));  // :119 with_capacity(nnz)
  vals = (T*)std::malloc(std::max<u64>(1, nnz) * sizeof(T));
  if (!indices || !vals) return false;
  std::vector<char> okv(rows_offset.size() - 1, 1);
  parallel_blocks(rows_offset, [&](size_t t, u64 tlo, u64 thi) {
    u64* tindices = indices + offsets[tlo];                    // :125-128 disjoint slices
    T* tvals = vals + offsets[tlo];
    u64 capacity = 0;                                          // :132
    for (u64 i = tlo; i < thi; ++i) capacity = std::max(capacity, row_nz[i]);
    HashMap<T> hm(capacity);                                   // :133
    u64 curr = 0;                                              // :134
    std::vector<std::pair<u32, T>> row;
    for (u64 i = tlo; i < thi; ++i) {
      if (row_nz[i] == 0) continue;                            // :141-143
      hm.shrink_to(row_nz[i]);                                 // :144
      for (u64 e = a.offsets[i]; e < a.offsets[i + 1]; ++e) {  // :145-149
        u64 k = a.indices[e];
        T t0 = a.vals[e];
        for (u64 j = b.offsets[k]; j < b.offsets[k + 1]; ++j) {  // :150-155
          T t1 = Arith<T>::mul(t0, b.vals[j]);                 // :154 `t * t1`
          hm.upsert((u32)b.indices[j], t1, Arith<T>::add_assign);  // :157-161
        }
      }
      if (sorted) {                                            // :164-175
        row.clear();
        hm.drain([&](u32 c, T v) { row.emplace_back(c, v); });
        // sort_unstable_by_key(col): keys are distinct so any sort agrees
        std::sort(row.begin(), row.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
        for (auto& cv : row) { tindices[curr] = cv.first; tvals[curr] = cv.second; curr += 1; }
      } else {                                                 // :176-186 slot order
        hm.drain([&](u32 c, T v) { tindices[curr] = c; tvals[curr] = v; curr += 1; });
      }
    }
    if (curr != offsets[thi] - offsets[tlo]) okv[t] = 0;       // :190 assert_eq!
  });
  for (char c : okv) if (!c) return false;
  return true;
}

// mul_hash.rs:13-36 mul_hash::<B1, B2>
template <class T>
int mul_hash(const CsrView<T>& a, const CsrView<T>& b, bool sorted, u64 tnum, u64** c_offsets, u64** c_indices,
             T** c_vals, u64* c_nnz, u64* flops_out) {
  if (tnum == 0) tnum = std::max(1u, std::thread::hardware_concurrency());  // :54 num_cpus::get()
  std::vector<u64> row_nz, rows_offset, offsets;
  if (!rows_to_threads(a, b, tnum, row_nz, rows_offset)) return 2;          // :17
  if (flops_out) { u64 s = 0; for (u64 x : row_nz) s += x; *flops_out = s; }
  mul_hash_symbolic(a, b, row_nz, rows_offset);                             // :22
  u64* indices = nullptr; T* vals = nullptr;
  if (!mul_hash_numeric(a, b, row_nz, rows_offset, sorted, offsets, indices, vals)) {  // :27-28
    std::free(indices); std::free(vals);
    return 3;
  }
  u64* offs = (u64*)std::malloc(offsets.size() * sizeof(u64));
  std::memcpy(offs, offsets.data(), offsets.size() * sizeof(u64));
  *c_offsets = offs; *c_indices = indices; *c_vals = vals; *c_nnz = offsets.back();
  return 0;
}

// ---------------------------------------------------------------------------------
// DokMatrix semantics (spam_dok/src/lib.rs)
// ---------------------------------------------------------------------------------
template <class T> inline bool is_zero(T t) { return t == T(0); }  // num_traits::Zero::is_zero

// Sequential DokMatrix::set_element stream (spam_dok/src/lib.rs:167-176: zero => remove,
// else insert/replace) followed by From<DokMatrix> for CsrMatrix<T,true>
// (spam_csr/src/lib.rs:315-334: BTreeMap order = (row, col) lexicographic
// (spam_dok/src/lib.rs:234-242), empty rows back-filled with repeated offsets :321,:325).
template <class T>
int dok_to_csr(u64 rows, u64 cols, u64 n, const u64* ri, const u64* ci, const T* v, u64** c_offsets,
               u64** c_indices, T** c_vals, u64* c_nnz) {
  std::map<std::pair<u64, u64>, T> entries;
  for (u64 t = 0; t < n; ++t) {
    if (!(ri[t] < rows && ci[t] < cols)) return 4;  // IndexError (spam_dok/src/lib.rs:168-170)
    auto key = std::make_pair(ri[t], ci[t]);
    if (is_zero(v[t])) entries.erase(key); else entries[key] = v[t];
  }
  u64 nnz = entries.size();
  u64* offs = (u64*)std::malloc((rows + 1) * sizeof(u64));
  u64* idx = (u64*)std::malloc(std::max<u64>(1, nnz) * sizeof(u64));
  T* vals = (T*)std::malloc(std::max<u64>(1, nnz) * sizeof(T));
  u64 filled = 0, pos = 0;  // offsets.len() so far
  for (auto& kv : entries) {
    u64 i = kv.first.first;
    while (filled < i + 1) offs[filled++] = pos;  // lib.rs:321
    vals[pos] = kv.second; idx[pos] = kv.first.second; pos++;
  }
  while (filled < rows + 1) offs[filled++] = pos;  // lib.rs:325
  *c_offsets = offs; *c_indices = idx; *c_vals = vals; *c_nnz = nnz;
  return 0;
}

// Dense triple-loop product, the reference's test oracle for tiny shapes
// (spam_dok/src/lib.rs:206-233): t = 0; for k: t = t + a_ik * b_kj (missing => zero);
// set_element drops zeros.  Output: dense row-major l x n array (zeros = absent).
template <class T>
void dok_dense_mul(u64 l, u64 m, u64 n, const T* a /*l x m*/, const T* b /*m x n*/, T* c /*l x n*/) {
  for (u64 i = 0; i < l; ++i)
    for (u64 j = 0; j < n; ++j) {
      T t = T(0);
      for (u64 k = 0; k < m; ++k) { T p = Arith<T>::mul(a[i * m + k], b[k * n + j]); Arith<T>::add_assign(t, p); }
      c[i * n + j] = t;
    }
}

// SpMV — ABSENT in the reference.  Only reference-expressible form:
// a.mul_hash::<_, true>(&x) with x an n x 1 CsrMatrix holding one explicit entry per k.
// Then every product lands on column 0: y_i = ((a_i,k1 * x_k1) + a_i,k2 * x_k2) + ... in
// A-row storage order, first product stored (mul_hash.rs:145-162); rows with no A
// entry produce no C entry, which the dense-output API defines as T::zero().
// Parity unpinned by any reference test.
template <class T>
void spmv_as_mul_hash(const CsrView<T>& a, const T* x, T* y) {
  for (u64 i = 0; i < a.rows; ++i) {
    bool first = true;
    T acc = T(0);
    for (u64 e = a.offsets[i]; e < a.offsets[i + 1]; ++e) {
      T p = Arith<T>::mul(a.vals[e], x[a.indices[e]]);
      if (first) { acc = p; first = false; } else Arith<T>::add_assign(acc, p);
    }
    y[i] = acc;
  }
}

enum { DT_F32 = 0, DT_F64 = 1, DT_I32 = 2, DT_I64 = 3, DT_I8 = 4 };

template <class F>
int dispatch(int dtype, F f) {
  switch (dtype) {
    case DT_F32: return f(float());
    case DT_F64: return f(double());
    case DT_I32: return f(int32_t());
    case DT_I64: return f(int64_t());
    case DT_I8: return f(int8_t());
    default: return 1;
  }
}

}  // namespace

extern "C" {

// Returns 0 ok; 1 bad dtype; 2 flop overflow (reference panics mul_hash.rs:47-48 /
// lib.rs:270); 3 internal assert (mul_hash.rs:190); 4 IndexError.
// Output buffers are malloc'd; release with oracle_free.  tnum = 0 => hardware threads.
int oracle_mul_hash(int dtype, uint64_t a_rows, uint64_t a_cols, const uint64_t* a_off, const uint64_t* a_idx,
                    const void* a_val, uint64_t b_rows, uint64_t b_cols, const uint64_t* b_off,
                    const uint64_t* b_idx, const void* b_val, int sorted, uint64_t tnum, uint64_t** c_off,
                    uint64_t** c_idx, void** c_val, uint64_t* c_nnz, uint64_t* flops) {
  (void)b_rows;  // the reference performs no dimension check
  return dispatch(dtype, [&](auto tag) {
    using T = decltype(tag);
    CsrView<T> a{a_rows, a_cols, (const T*)a_val, a_idx, a_off};
    CsrView<T> b{b_rows, b_cols, (const T*)b_val, b_idx, b_off};
    T* cv = nullptr;
    int rc = mul_hash(a, b, sorted != 0, tnum, c_off, c_idx, &cv, c_nnz, flops);
    *c_val = cv;
    return rc;
  });
}

// rows_to_threads alone (mul_hash.rs:38-64): flop[rows], rows_offset[tnum+1]
int oracle_rows_to_threads(uint64_t a_rows, const uint64_t* a_off, const uint64_t* a_idx, const uint64_t* b_off,
                           uint64_t tnum, uint64_t* flop_out, uint64_t* rows_offset_out) {
  CsrView<double> a{a_rows, 0, nullptr, a_idx, a_off};
  CsrView<double> b{0, 0, nullptr, nullptr, b_off};
  std::vector<u64> row_nz, ro;
  if (!rows_to_threads(a, b, tnum, row_nz, ro)) return 2;
  std::memcpy(flop_out, row_nz.data(), row_nz.size() * sizeof(u64));
  std::memcpy(rows_offset_out, ro.data(), ro.size() * sizeof(u64));
  return 0;
}

// symbolic only: row nnz of C (mul_hash.rs:66-103)
int oracle_symbolic(uint64_t a_rows, const uint64_t* a_off, const uint64_t* a_idx, const uint64_t* b_off,
                    const uint64_t* b_idx, uint64_t tnum, uint64_t* row_nnz_out) {
  CsrView<double> a{a_rows, 0, nullptr, a_idx, a_off};
  CsrView<double> b{0, 0, nullptr, b_idx, b_off};
  if (tnum == 0) tnum = std::max(1u, std::thread::hardware_concurrency());
  std::vector<u64> row_nz, ro;
  if (!rows_to_threads(a, b, tnum, row_nz, ro)) return 2;
  mul_hash_symbolic(a, b, row_nz, ro);
  std::memcpy(row_nnz_out, row_nz.data(), row_nz.size() * sizeof(u64));
  return 0;
}

int oracle_dok_to_csr(int dtype, uint64_t rows, uint64_t cols, uint64_t n, const uint64_t* ri, const uint64_t* ci,
                      const void* v, uint64_t** c_off, uint64_t** c_idx, void** c_val, uint64_t* c_nnz) {
  return dispatch(dtype, [&](auto tag) {
    using T = decltype(tag);
    T* cv = nullptr;
    int rc = dok_to_csr<T>(rows, cols, n, ri, ci, (const T*)v, c_off, c_idx, &cv, c_nnz);
    *c_val = cv;
    return rc;
  });
}

// Matrix::transpose of CsrMatrix (spam_csr/src/lib.rs:256-264):
//   for (j, i) in iproduct!(0..cols, 0..rows) {
//     if let Some(t) = self.set_element((i, j), T::zero()).unwrap() { new.set_element((j, i), t).unwrap(); } }
// set_element on a stored entry swaps the value in and returns the old one, explicit zeros included
// (lib.rs:213-226, 238-242); on a missing entry it inserts the zero into `self` (which is consumed, so
// that is unobservable) and returns None.  new.set_element((j, i), t) finds no entry and inserts at the end
// of row j in both the sorted and the unsorted branch because i only grows (lib.rs:227-236, 243-252).
// `literal` = exactly those loops (O(cols * nnz), tests only); otherwise the equivalent stable counting
// sort by column.  Outputs are caller-allocated: t_off[cols+1], t_idx[nnz], t_val[nnz].
int oracle_transpose(int dtype, uint64_t rows, uint64_t cols, const uint64_t* off, const uint64_t* idx, const void* val,
                     int literal, uint64_t* t_off, uint64_t* t_idx, void* t_val) {
  return dispatch(dtype, [&](auto tag) {
    using T = decltype(tag);
    const T* v = (const T*)val;
    T* tv = (T*)t_val;
    if (literal) {
      u64 out = 0;
      t_off[0] = 0;
      for (u64 j = 0; j < cols; ++j) {
        for (u64 i = 0; i < rows; ++i) {
          for (u64 e = off[i]; e < off[i + 1]; ++e) {
            if (idx[e] == j) {  // first match: binary_search / position on distinct columns
              t_idx[out] = i;
              tv[out] = v[e];
              ++out;
              break;
            }
          }
        }
        t_off[j + 1] = out;
      }
      return 0;
    }
    std::vector<u64> cnt(cols + 1, 0);
    for (u64 e = 0; e < off[rows]; ++e) {
      if (idx[e] >= cols) return 4;
      ++cnt[idx[e] + 1];
    }
    for (u64 j = 0; j < cols; ++j) cnt[j + 1] += cnt[j];
    std::memcpy(t_off, cnt.data(), (cols + 1) * sizeof(u64));
    for (u64 i = 0; i < rows; ++i) {
      for (u64 e = off[i]; e < off[i + 1]; ++e) {
        const u64 p = cnt[idx[e]]++;
        t_idx[p] = i;
        tv[p] = v[e];
      }
    }
    return 0;
  });
}

// impl Add / Sub for CsrMatrix -> apply_elementwise (spam_csr/src/lib.rs:83-149, 276-290).
//   IS_SORTED = true  (lib.rs:102-118): merge_join_by on the column of the two rows;
//       Both -> f(t1, t2), Left -> f(t, 0), Right -> f(0, t); output in column order.
//   IS_SORTED = false (lib.rs:119-137): the left row goes into a std HashMap; every right entry does
//       entry = f(entry-or-zero, t); entries only in the left row keep their value untouched.  The map's
//       iteration order is unspecified (RandomState), so this restatement emits the row sorted by column.
// op: 0 add, 1 sub (wrapping for integers, like the device scalar types).  No entry is ever dropped.
// Outputs are malloc'ed; free with oracle_free.  rc 2: shapes differ (the reference asserts, lib.rs:87-91).
int oracle_ewise(int dtype, int op, int is_sorted, uint64_t rows, uint64_t a_cols, uint64_t b_rows, uint64_t b_cols,
                 const uint64_t* a_off, const uint64_t* a_idx, const void* a_val, const uint64_t* b_off,
                 const uint64_t* b_idx, const void* b_val, uint64_t** c_off, uint64_t** c_idx, void** c_val,
                 uint64_t* c_nnz) {
  if (rows != b_rows || a_cols != b_cols) return 2;
  return dispatch(dtype, [&](auto tag) {
    using T = decltype(tag);
    const T* av = (const T*)a_val;
    const T* bv = (const T*)b_val;
    auto f = [op](T x, T y) -> T {
      if constexpr (std::is_integral<T>::value) {
        using U = typename std::make_unsigned<T>::type;
        return (T)(op == 0 ? (U)((U)x + (U)y) : (U)((U)x - (U)y));
      } else {
        return op == 0 ? x + y : x - y;
      }
    };
    std::vector<u64> off(rows + 1, 0), idx;
    std::vector<T> val;
    for (u64 r = 0; r < rows; ++r) {
      std::vector<std::pair<u64, T>> left, right, out;
      for (u64 e = a_off[r]; e < a_off[r + 1]; ++e) left.emplace_back(a_idx[e], av[e]);
      for (u64 e = b_off[r]; e < b_off[r + 1]; ++e) right.emplace_back(b_idx[e], bv[e]);
      if (is_sorted) {
        size_t i = 0, j = 0;
        while (i < left.size() || j < right.size()) {
          if (j == right.size() || (i < left.size() && left[i].first < right[j].first)) {
            out.emplace_back(left[i].first, f(left[i].second, T(0))); ++i;
          } else if (i == left.size() || right[j].first < left[i].first) {
            out.emplace_back(right[j].first, f(T(0), right[j].second)); ++j;
          } else {
            out.emplace_back(left[i].first, f(left[i].second, right[j].second)); ++i; ++j;
          }
        }
      } else {
        std::map<u64, T> row;  // ordered stand-in for the HashMap: only the iteration order differs
        for (auto& p : left) row[p.first] = p.second;  // collect(): a later duplicate key would replace
        for (auto& p : right) {
          auto it = row.find(p.first);
          if (it == row.end()) it = row.emplace(p.first, T(0)).first;
          it->second = f(it->second, p.second);
        }
        for (auto& p : row) out.emplace_back(p.first, p.second);
      }
      for (auto& p : out) { idx.push_back(p.first); val.push_back(p.second); }
      off[r + 1] = idx.size();
    }
    *c_nnz = idx.size();
    *c_off = (u64*)std::malloc((rows + 1) * sizeof(u64));
    *c_idx = (u64*)std::malloc(std::max<size_t>(1, idx.size()) * sizeof(u64));
    T* cv = (T*)std::malloc(std::max<size_t>(1, val.size()) * sizeof(T));
    std::memcpy(*c_off, off.data(), (rows + 1) * sizeof(u64));
    if (!idx.empty()) std::memcpy(*c_idx, idx.data(), idx.size() * sizeof(u64));
    if (!val.empty()) std::memcpy(cv, val.data(), val.size() * sizeof(T));
    *c_val = cv;
    return 0;
  });
}

int oracle_dok_dense_mul(int dtype, uint64_t l, uint64_t m, uint64_t n, const void* a, const void* b, void* c) {
  return dispatch(dtype, [&](auto tag) {
    using T = decltype(tag);
    dok_dense_mul<T>(l, m, n, (const T*)a, (const T*)b, (T*)c);
    return 0;
  });
}

int oracle_spmv(int dtype, uint64_t rows, uint64_t cols, const uint64_t* off, const uint64_t* idx, const void* val,
                const void* x, void* y) {
  return dispatch(dtype, [&](auto tag) {
    using T = decltype(tag);
    CsrView<T> a{rows, cols, (const T*)val, idx, off};
    spmv_as_mul_hash<T>(a, (const T*)x, (T*)y);
    return 0;
  });
}

// linprobe probes for the hand-derived known-answer tests
uint64_t oracle_table_size_for(uint64_t capacity) { return table_size_for(capacity); }
uint64_t oracle_hash(uint32_t key) { return mul_hash_of(key); }
// insert keys into a fresh HashSet::new(), optionally after shrink_to(cap_hint) when
// cap_hint != 0; returns len, writes upper_bound and the slot array prefix.
uint64_t oracle_hashset_run(const uint32_t* keys, uint64_t n, uint64_t cap_hint, uint64_t* upper_bound_out,
                            uint32_t* slots_out, uint64_t slots_cap) {
  HashSet hs;
  if (cap_hint) hs.shrink_to(cap_hint);
  for (u64 i = 0; i < n; ++i) hs.insert(keys[i]);
  *upper_bound_out = hs.upper_bound;
  for (u64 i = 0; i < std::min<u64>(slots_cap, hs.upper_bound); ++i) slots_out[i] = hs.slots[i];
  return hs.len();
}

// a set that was allocated for `initial_capacity` keys (HashSet::with_capacity, set.rs:27-29) and then told
// shrink_to(shrink_cap) (0: not called), as the symbolic pass does with a set that grew on an earlier row
uint64_t oracle_hashset_run2(const uint32_t* keys, uint64_t n, uint64_t initial_capacity, uint64_t shrink_cap,
                             uint64_t* upper_bound_out, uint64_t* allocated_out, uint32_t* slots_out, uint64_t slots_cap) {
  HashSet hs(initial_capacity);
  if (shrink_cap) hs.shrink_to(shrink_cap);
  for (u64 i = 0; i < n; ++i) hs.insert(keys[i]);
  *upper_bound_out = hs.upper_bound;
  *allocated_out = hs.slots.size();
  for (u64 i = 0; i < std::min<u64>(slots_cap, hs.upper_bound); ++i) slots_out[i] = hs.slots[i];
  return hs.len();
}

void oracle_free(void* p) { std::free(p); }
unsigned oracle_hardware_threads(void) { return std::max(1u, std::thread::hardware_concurrency()); }

}  // extern "C"
