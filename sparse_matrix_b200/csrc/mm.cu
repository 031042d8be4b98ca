// mm.cu — MatrixMarket coordinate text -> triplet stream (host code).
//
// Restates the grammar of `parse_matrix_market` (spam_dok/src/lib.rs:282-478, a nom parser) so that the same
// files give the same DokMatrix; the triplets then go through spam_dok_to_csr (DOK -> CSR on the device), which
// is what the reference's bench does with its on-disk inputs (spam_csr/src/lib.rs:419-431).
//   header   "%%MatrixMarket matrix coordinate" ' ' (integer|real|complex|pattern) ' '
//            (general|symmetric|skew-symmetric|hermitian) EOL                          (lib.rs:352-368)
//            pattern / skew-symmetric / hermitian hit todo!() in the reference: unsupported here too;
//            complex parses in the reference but has no device scalar: unsupported.
//   comments zero or more lines starting with '%'                                      (lib.rs:375)
//   size     usize ' ' usize ' ' usize EOL; the third number (entry count) is not used  (lib.rs:316-331,376)
//   entries  usize ' ' usize ' ' value EOL, as many lines as match: the first line that does not match ends
//            the list silently (fold_many0), whatever follows is ignored                (lib.rs:380-425)
//            integer value: '-'? digits parsed as i64;  real value: nom's recognize_float parsed as f64
//   zeros are skipped, a later entry with the same key replaces an earlier one (BTreeMap::insert), indices
//   are 1-based, `symmetric` inserts (r, c) and (c, r)                                  (lib.rs:332-351)
//   rows == 0 or cols == 0 -> HasZeroDimension                                          (lib.rs:446-449)
// EOL is "\n" or "\r\n" (nom line_ending).  Single spaces only, no leading/trailing blanks.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/spam_cuda.h"

namespace {

struct Cur {
  const char* p;
  const char* end;
  bool eof() const { return p >= end; }
};

bool eat_tag(Cur& c, const char* tag) {
  const size_t n = std::strlen(tag);
  if ((size_t)(c.end - c.p) < n || std::memcmp(c.p, tag, n) != 0) return false;
  c.p += n;
  return true;
}
bool eat_char(Cur& c, char ch) {
  if (c.eof() || *c.p != ch) return false;
  ++c.p;
  return true;
}
bool eat_eol(Cur& c) {  // nom line_ending
  if (eat_char(c, '\n')) return true;
  if (c.end - c.p >= 2 && c.p[0] == '\r' && c.p[1] == '\n') { c.p += 2; return true; }
  return false;
}
size_t digits(const Cur& c, const char* from) {
  const char* q = from;
  while (q < c.end && *q >= '0' && *q <= '9') ++q;
  return (size_t)(q - from);
}

// recognize_int = '-'? digit1, then str::parse::<usize>: a '-' never parses as usize, overflow is an error
bool parse_usize(Cur& c, uint64_t* out) {
  const char* q = c.p;
  const bool neg = q < c.end && *q == '-';
  if (neg) ++q;
  const size_t n = digits(c, q);
  if (n == 0 || neg) return false;
  uint64_t v = 0;
  for (size_t i = 0; i < n; ++i) {
    const uint64_t d = (uint64_t)(q[i] - '0');
    if (v > (UINT64_MAX - d) / 10) return false;
    v = v * 10 + d;
  }
  *out = v;
  c.p = q + n;
  return true;
}

// recognize_int parsed as i64
bool parse_i64(Cur& c, int64_t* out) {
  const char* q = c.p;
  const bool neg = q < c.end && *q == '-';
  if (neg) ++q;
  const size_t n = digits(c, q);
  if (n == 0) return false;
  uint64_t v = 0;
  const uint64_t lim = neg ? (uint64_t)INT64_MAX + 1 : (uint64_t)INT64_MAX;
  for (size_t i = 0; i < n; ++i) {
    const uint64_t d = (uint64_t)(q[i] - '0');
    if (v > (lim - d) / 10) return false;
    v = v * 10 + d;
  }
  *out = neg ? (int64_t)(0 - v) : (int64_t)v;
  c.p = q + n;
  return true;
}

// nom recognize_float: [+-]? (digit1 ('.' digit0)? | '.' digit1) ([eE] [+-]? cut(digit1))?
// returns 0 no match, 1 ok, -1 hard failure (an exponent marker without digits: nom `cut`)
int parse_f64(Cur& c, double* out) {
  const char* q = c.p;
  if (q < c.end && (*q == '+' || *q == '-')) ++q;
  size_t n = digits(c, q);
  if (n) {
    q += n;
    if (q < c.end && *q == '.') { ++q; q += digits(c, q); }
  } else {
    if (!(q < c.end && *q == '.')) return 0;
    ++q;
    n = digits(c, q);
    if (n == 0) return 0;
    q += n;
  }
  if (q < c.end && (*q == 'e' || *q == 'E')) {
    ++q;
    if (q < c.end && (*q == '+' || *q == '-')) ++q;
    n = digits(c, q);
    if (n == 0) return -1;
    q += n;
  }
  const std::string s(c.p, q);
  const char* b = s.c_str();
  char* e = nullptr;
  *out = std::strtod(b[0] == '+' ? b + 1 : b, &e);  // decimal forms only: strtod and Rust's f64::from_str agree
  if (e == nullptr || *e != '\0') return 0;
  c.p = q;
  return 1;
}

int fail(spam_mm* out, int status, const char* msg) {
  std::snprintf(out->err, sizeof(out->err), "%s", msg);
  return status;
}

}  // namespace

extern "C" {

int spam_mm_parse(const char* text, uint64_t len, spam_mm* out) {
  if (!text || !out) return SPAM_EINVAL;
  std::memset(out, 0, sizeof(*out));
  Cur c{text, text + len};
  if (!eat_tag(c, "%%MatrixMarket matrix coordinate") || !eat_char(c, ' ')) return fail(out, SPAM_EINVAL, "bad header");
  int kind;
  if (eat_tag(c, "integer")) kind = SPAM_I64;
  else if (eat_tag(c, "real")) kind = SPAM_F64;
  else if (eat_tag(c, "complex")) return fail(out, SPAM_EDTYPE, "complex entries have no device scalar");
  else if (eat_tag(c, "pattern")) return fail(out, SPAM_EDTYPE, "entry type pattern unsupported (todo!() in the reference)");
  else return fail(out, SPAM_EINVAL, "bad entry type");
  if (!eat_char(c, ' ')) return fail(out, SPAM_EINVAL, "bad header");
  bool symmetric;
  if (eat_tag(c, "general")) symmetric = false;
  else if (eat_tag(c, "symmetric")) symmetric = true;
  else if (eat_tag(c, "skew-symmetric") || eat_tag(c, "hermitian"))
    return fail(out, SPAM_EDTYPE, "matrix shape unsupported (todo!() in the reference)");
  else return fail(out, SPAM_EINVAL, "bad matrix shape");
  if (!eat_eol(c)) return fail(out, SPAM_EINVAL, "bad header");
  for (;;) {  // many0(delimited(char('%'), not_line_ending, line_ending))
    Cur t = c;
    if (!eat_char(t, '%')) break;
    while (!t.eof() && *t.p != '\n' && *t.p != '\r') ++t.p;
    if (!eat_eol(t)) break;
    c = t;
  }
  uint64_t rows, cols, declared;
  if (!parse_usize(c, &rows) || !eat_char(c, ' ') || !parse_usize(c, &cols) || !eat_char(c, ' ') ||
      !parse_usize(c, &declared) || !eat_eol(c))
    return fail(out, SPAM_EINVAL, "bad size line");
  std::vector<uint64_t> tr, tc;
  std::vector<int64_t> vi;
  std::vector<double> vf;
  for (;;) {
    Cur t = c;
    uint64_t r, col;
    if (!parse_usize(t, &r) || !eat_char(t, ' ') || !parse_usize(t, &col) || !eat_char(t, ' ')) break;
    int64_t iv = 0;
    double fv = 0;
    if (kind == SPAM_I64) {
      if (!parse_i64(t, &iv)) break;
    } else {
      const int st = parse_f64(t, &fv);
      if (st < 0) return fail(out, SPAM_EINVAL, "exponent without digits");
      if (st == 0) break;
    }
    if (!eat_eol(t)) break;
    c = t;
    const bool zero = kind == SPAM_I64 ? iv == 0 : fv == 0.0;  // is_zero: -0.0 is zero, NaN is not
    if (zero) continue;
    if (r == 0 || col == 0) return fail(out, SPAM_EINDEX, "index 0 in a 1-based file (r - 1 underflows in the reference)");
    const int copies = symmetric ? 2 : 1;
    for (int k = 0; k < copies; ++k) {
      tr.push_back((k == 0 ? r : col) - 1);
      tc.push_back((k == 0 ? col : r) - 1);
      if (kind == SPAM_I64) vi.push_back(iv); else vf.push_back(fv);
    }
  }
  if (rows == 0 || cols == 0) return fail(out, SPAM_EDIM, "HasZeroDimension");
  const size_t n = tr.size();
  out->kind = kind; out->rows = rows; out->cols = cols; out->declared_entries = declared; out->n = n;
  out->tri_rows = (uint64_t*)std::malloc((n ? n : 1) * sizeof(uint64_t));
  out->tri_cols = (uint64_t*)std::malloc((n ? n : 1) * sizeof(uint64_t));
  out->tri_vals = std::malloc((n ? n : 1) * 8);
  if (!out->tri_rows || !out->tri_cols || !out->tri_vals) {
    std::free(out->tri_rows); std::free(out->tri_cols); std::free(out->tri_vals);
    out->tri_rows = out->tri_cols = nullptr; out->tri_vals = nullptr;
    return fail(out, SPAM_ENOMEM, "out of host memory");
  }
  if (n) {
    std::memcpy(out->tri_rows, tr.data(), n * 8);
    std::memcpy(out->tri_cols, tc.data(), n * 8);
    if (kind == SPAM_I64) std::memcpy(out->tri_vals, vi.data(), n * 8); else std::memcpy(out->tri_vals, vf.data(), n * 8);
  }
  return SPAM_OK;
}

void spam_mm_free(spam_mm* m) {
  if (!m) return;
  std::free(m->tri_rows); std::free(m->tri_cols); std::free(m->tri_vals);
  m->tri_rows = m->tri_cols = nullptr; m->tri_vals = nullptr; m->n = 0;
}

}  // extern "C"
