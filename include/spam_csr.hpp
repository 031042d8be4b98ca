// spam_csr.hpp — C++ host-side mirror of the reference's operator interface for the hot path, over
// the C ABI (include/spam_cuda.h).  The reference is compiled Rust and no Rust toolchain exists in
// the build image, so this header plays the part of the vendored spam_csr crate: same names, same
// argument meaning, same error behaviour (a failing status throws where the reference panics).
//
//   spam::CsrMatrix<T, IS_SORTED>          spam_csr/src/lib.rs:25-32 (rows, cols, vals, indices, offsets)
//   CsrMatrix::invariants()                spam_csr/src/lib.rs:47-81,152-160
//   CsrMatrix::mul_hash<B2>(rhs)           spam_csr/src/mul_hash.rs:13-36
//   operator*(const CsrMatrix&, ...)       impl Mul for &CsrMatrix, lib.rs:292-297 (Output: IS_SORTED=false)
//   CsrMatrix<T,true>::from(DokMatrix)     lib.rs:315-334
//   DokMatrix::set_element                 spam_dok/src/lib.rs:167-176
//   CsrMatrix::spmv(x)                     new API (no SpMV in the reference)
// All arithmetic runs in libspam_cuda.so; nothing here computes a product on the CPU.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_set>
#include <utility>
#include <vector>

#include "spam_cuda.h"

namespace spam {

struct IndexError : std::out_of_range {  // spam_matrix/src/lib.rs:13
  IndexError() : std::out_of_range("IndexError") {}
};

template <class T> struct device_scalar;  // sealed: f32, f64, i32, i64 (no CPU fallback for other T)
template <> struct device_scalar<float> { static constexpr int dtype = SPAM_F32; };
template <> struct device_scalar<double> { static constexpr int dtype = SPAM_F64; };
template <> struct device_scalar<int32_t> { static constexpr int dtype = SPAM_I32; };
template <> struct device_scalar<int64_t> { static constexpr int dtype = SPAM_I64; };

class Handle {
 public:
  explicit Handle(int device = 0) {
    int st = spam_cuda_create(&h_, device);
    if (st != SPAM_OK) throw std::runtime_error(std::string("spam_cuda_create: ") + spam_strerror(st));
  }
  ~Handle() { if (h_) spam_cuda_destroy(h_); }
  Handle(const Handle&) = delete;
  Handle& operator=(const Handle&) = delete;
  spam_handle* get() const { return h_; }
  void check(int st) const {
    if (st == SPAM_OK) return;
    std::string msg = std::string(spam_strerror(st)) + " (" + spam_last_error(h_) + ")";
    if (st == SPAM_EINDEX) throw IndexError();
    throw std::runtime_error(msg);  // the reference panics here
  }
  static Handle& thread_default() {  // the handle is not thread-safe: one per thread
    thread_local Handle h(0);
    return h;
  }
 private:
  spam_handle* h_ = nullptr;
};

template <class T>
class DokMatrix {  // spam_dok/src/lib.rs:32-36
 public:
  DokMatrix(uint64_t rows, uint64_t cols) : rows_(rows), cols_(cols) {
    if (!rows || !cols) throw std::invalid_argument("rows and cols are NonZeroUsize");
  }
  uint64_t rows() const { return rows_; }
  uint64_t cols() const { return cols_; }
  size_t nnz() const { return entries_.size(); }
  // zero removes, anything else inserts or replaces (lib.rs:167-176)
  void set_element(uint64_t i, uint64_t j, T t) {
    if (!(i < rows_ && j < cols_)) throw IndexError();
    if (t == T(0)) entries_.erase({i, j}); else entries_[{i, j}] = t;
  }
  const std::map<std::pair<uint64_t, uint64_t>, T>& entries() const { return entries_; }
 private:
  uint64_t rows_, cols_;
  std::map<std::pair<uint64_t, uint64_t>, T> entries_;  // BTreeMap order: (row, col)
};

template <class T, bool IS_SORTED>
class CsrMatrix {
 public:
  uint64_t rows, cols;
  std::vector<T> vals;
  std::vector<uint64_t> indices, offsets;

  CsrMatrix(uint64_t r, uint64_t c) : rows(r), cols(c), offsets(r + 1, 0) {
    if (!r || !c) throw std::invalid_argument("rows and cols are NonZeroUsize");
  }
  CsrMatrix(uint64_t r, uint64_t c, std::vector<T> v, std::vector<uint64_t> i, std::vector<uint64_t> o)
      : rows(r), cols(c), vals(std::move(v)), indices(std::move(i)), offsets(std::move(o)) {}

  static CsrMatrix identity(uint64_t n) {  // lib.rs:177-185
    CsrMatrix m(n, n);
    m.vals.assign(n, T(1));
    m.indices.resize(n);
    for (uint64_t i = 0; i < n; ++i) { m.indices[i] = i; m.offsets[i] = i; }
    m.offsets[n] = n;
    return m;
  }
  size_t nnz() const { return indices.size(); }

  bool invariants() const {  // lib.rs:47-81
    if (indices.size() != vals.size()) return false;                     // 1
    if (offsets.size() != rows + 1) return false;                        // 2
    for (uint64_t r = 0; r < rows; ++r) if (offsets[r + 1] < offsets[r]) return false;  // 3
    if (offsets[rows] != indices.size()) return false;                   // 4
    for (auto c : indices) if (c >= cols) return false;                  // 5
    if (offsets[0] != 0) return false;                                   // 7
    for (uint64_t r = 0; r < rows; ++r) {                                // 6
      if (IS_SORTED) {
        for (uint64_t e = offsets[r] + 1; e < offsets[r + 1]; ++e) if (indices[e - 1] >= indices[e]) return false;
      } else {
        std::unordered_set<uint64_t> seen(indices.begin() + offsets[r], indices.begin() + offsets[r + 1]);
        if (seen.size() != offsets[r + 1] - offsets[r]) return false;
      }
    }
    return true;
  }

  // mul_hash::<B1, B2>: rows come back sorted by column, valid for either B2; with B2 = false and
  // `reference_order` they come back in the reference's own B2 = false order (the slot order of its map,
  // mul_hash.rs:176-186), column for column.
  template <bool B2, bool B1>
  CsrMatrix<T, B2> mul_hash(const CsrMatrix<T, B1>& rhs, Handle& h = Handle::thread_default(),
                            bool reference_order = false) const {
    CsrMatrix<T, B2> c(rows, rhs.cols);
    uint64_t nnz = 0;
    h.check(spam_spgemm_symbolic(h.get(), device_scalar<T>::dtype, rows, cols, offsets.data(), indices.data(), vals.data(),
                                 rhs.rows, rhs.cols, rhs.offsets.data(), rhs.indices.data(), rhs.vals.data(),
                                 c.offsets.data(), &nnz));
    c.indices.resize(nnz);  // Vec::with_capacity(nnz), mul_hash.rs:119
    c.vals.resize(nnz);
    h.check(spam_spgemm_numeric(h.get(), c.indices.data(), c.vals.data(), (!B2 && reference_order) ? 0 : 1));
    return c;
  }

  std::vector<T> spmv(const std::vector<T>& x, Handle& h = Handle::thread_default()) const {
    if (x.size() != cols) throw std::runtime_error("dimension mismatch");
    std::vector<T> y(rows);
    h.check(spam_spmv(h.get(), device_scalar<T>::dtype, rows, cols, offsets.data(), indices.data(), vals.data(), x.data(),
                      y.data()));
    return y;
  }

  // impl Add / Sub for CsrMatrix -> apply_elementwise (spam_csr/src/lib.rs:83-149, 276-290)
  CsrMatrix<T, IS_SORTED> ewise(const CsrMatrix<T, IS_SORTED>& rhs, int op, Handle& h = Handle::thread_default()) const {
    if (rows != rhs.rows || cols != rhs.cols) throw std::runtime_error("matrices must have identical dimensions");
    CsrMatrix<T, IS_SORTED> c(rows, cols);
    uint64_t nnz = 0;
    h.check(spam_csr_ewise(h.get(), op | (IS_SORTED ? 0 : 2), device_scalar<T>::dtype, rows, cols, offsets.data(),
                           indices.data(), vals.data(), rhs.offsets.data(), rhs.indices.data(), rhs.vals.data(),
                           c.offsets.data(), &nnz));
    c.indices.resize(nnz);
    c.vals.resize(nnz);
    h.check(spam_csr_ewise_fetch(h.get(), c.indices.data(), c.vals.data()));
    return c;
  }
  CsrMatrix<T, IS_SORTED> operator+(const CsrMatrix<T, IS_SORTED>& rhs) const { return ewise(rhs, 0); }
  CsrMatrix<T, IS_SORTED> operator-(const CsrMatrix<T, IS_SORTED>& rhs) const { return ewise(rhs, 1); }

  // Matrix::transpose (spam_csr/src/lib.rs:256-264); rows of the result are sorted by column
  CsrMatrix<T, true> transpose(Handle& h = Handle::thread_default()) const {
    CsrMatrix<T, true> t(cols, rows);
    t.indices.resize(indices.size());
    t.vals.resize(vals.size());
    h.check(spam_csr_transpose(h.get(), device_scalar<T>::dtype, rows, cols, offsets.data(), indices.data(), vals.data(),
                               t.offsets.data(), t.indices.data(), t.vals.data()));
    return t;
  }

  // impl From<DokMatrix<T>> for CsrMatrix<T, true>; from_triplets replays a set_element stream
  static CsrMatrix<T, true> from_triplets(uint64_t r, uint64_t c, const std::vector<uint64_t>& tr,
                                          const std::vector<uint64_t>& tc, const std::vector<T>& tv,
                                          Handle& h = Handle::thread_default()) {
    CsrMatrix<T, true> m(r, c);
    uint64_t nnz = 0;
    h.check(spam_dok_to_csr(h.get(), device_scalar<T>::dtype, r, c, tv.size(), tr.data(), tc.data(), tv.data(),
                            m.offsets.data(), &nnz));
    m.indices.resize(nnz);
    m.vals.resize(nnz);
    h.check(spam_dok_to_csr_fetch(h.get(), m.indices.data(), m.vals.data()));
    return m;
  }
  static CsrMatrix<T, true> from(const DokMatrix<T>& d, Handle& h = Handle::thread_default()) {
    std::vector<uint64_t> tr, tc;
    std::vector<T> tv;
    for (auto& kv : d.entries()) { tr.push_back(kv.first.first); tc.push_back(kv.first.second); tv.push_back(kv.second); }
    return from_triplets(d.rows(), d.cols(), tr, tc, tv, h);
  }
};

// impl Mul for &CsrMatrix<T, B>: Output = CsrMatrix<T, false>  (lib.rs:292-297)
template <class T, bool B>
CsrMatrix<T, false> operator*(const CsrMatrix<T, B>& a, const CsrMatrix<T, B>& b) {
  return a.template mul_hash<false>(b);
}

// parse_matrix_market (spam_dok/src/lib.rs:282-478) followed by CsrMatrix::from(dok): T = int64_t for
// `integer` files (MatrixType::Integer), double for `real` ones (MatrixType::Real); anything else throws.
template <class T>
CsrMatrix<T, true> from_matrix_market(const std::string& text, Handle& h = Handle::thread_default()) {
  static_assert(std::is_same<T, int64_t>::value || std::is_same<T, double>::value, "MatrixType::Integer is i64, Real is f64");
  spam_mm mm;
  const int st = spam_mm_parse(text.data(), text.size(), &mm);
  if (st == SPAM_EINDEX) throw IndexError();
  if (st != SPAM_OK) throw std::runtime_error(std::string("FromMatrixMarketError: ") + mm.err);
  if (mm.kind != device_scalar<T>::dtype) { spam_mm_free(&mm); throw std::runtime_error("entry type of the file differs from T"); }
  std::vector<uint64_t> tr(mm.tri_rows, mm.tri_rows + mm.n), tc(mm.tri_cols, mm.tri_cols + mm.n);
  std::vector<T> tv((const T*)mm.tri_vals, (const T*)mm.tri_vals + mm.n);
  const uint64_t rows = mm.rows, cols = mm.cols;
  spam_mm_free(&mm);
  return CsrMatrix<T, true>::from_triplets(rows, cols, tr, tc, tv, h);
}

}  // namespace spam
