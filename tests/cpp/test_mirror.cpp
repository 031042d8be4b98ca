// C++ host-mirror test (GPU): the reference's own property test shape
// (spam_csr/src/tests.rs:356-371: mul_hash == dense DokMatrix product) through include/spam_csr.hpp.
#include <cstdio>
#include <random>

#include "../../include/spam_csr.hpp"

template <class T>
static int run(const char* name, std::mt19937_64& rng) {
  int fails = 0;
  for (int it = 0; it < 60; ++it) {
    const uint64_t l = 1 + rng() % 6, m = 1 + rng() % 6, n = 1 + rng() % 6;
    spam::DokMatrix<T> da(l, m), db(m, n);
    for (uint64_t k = rng() % (2 * l * m + 1); k-- > 0;) da.set_element(rng() % l, rng() % m, (T)((int)(rng() % 9) - 4));
    for (uint64_t k = rng() % (2 * m * n + 1); k-- > 0;) db.set_element(rng() % m, rng() % n, (T)((int)(rng() % 9) - 4));
    auto a = spam::CsrMatrix<T, true>::from(da);
    auto b = spam::CsrMatrix<T, true>::from(db);
    if (!a.invariants() || !b.invariants()) { ++fails; continue; }
    auto c = a * b;  // impl Mul for &CsrMatrix: CsrMatrix<T, false>
    auto cs = a.template mul_hash<true>(b);
    if (!c.invariants() || !cs.invariants()) { ++fails; continue; }
    // dense DokMatrix product (spam_dok/src/lib.rs:206-233), compared zero-insensitively
    std::vector<T> want(l * n, T(0)), got(l * n, T(0));
    for (auto& ea : da.entries())
      for (auto& eb : db.entries())
        if (ea.first.second == eb.first.first) want[ea.first.first * n + eb.first.second] += ea.second * eb.second;
    for (uint64_t r = 0; r < l; ++r)
      for (uint64_t e = cs.offsets[r]; e < cs.offsets[r + 1]; ++e) got[r * n + cs.indices[e]] = cs.vals[e];
    if (want != got || c.indices != cs.indices || c.vals != cs.vals) ++fails;
    // B2 = false in the reference's slot order: the same entries per row, permuted
    auto cu = a.template mul_hash<false>(b, spam::Handle::thread_default(), true);
    if (!cu.invariants() || cu.offsets != cs.offsets) { ++fails; continue; }
    std::vector<T> gotu(l * n, T(0));
    for (uint64_t r = 0; r < l; ++r)
      for (uint64_t e = cu.offsets[r]; e < cu.offsets[r + 1]; ++e) gotu[r * n + cu.indices[e]] = cu.vals[e];
    if (gotu != got) ++fails;
    // spmv against the same dense data
    std::vector<T> x(m);
    for (auto& v : x) v = (T)((int)(rng() % 7) - 3);
    auto y = a.spmv(x);
    std::vector<T> yw(l, T(0));
    for (auto& ea : da.entries()) yw[ea.first.first] += ea.second * x[ea.first.second];
    if (y != yw) ++fails;
    // transpose (lib.rs:256-264) against the DOK transpose (spam_dok/src/lib.rs:178-188): entry (i, j) -> (j, i)
    auto t = a.transpose();
    if (!t.invariants() || t.rows != m || t.cols != l || t.indices.size() != a.indices.size()) { ++fails; continue; }
    std::vector<T> dt(m * l, T(0)), dw(m * l, T(0));
    for (uint64_t r = 0; r < m; ++r)
      for (uint64_t e = t.offsets[r]; e < t.offsets[r + 1]; ++e) dt[r * l + t.indices[e]] = t.vals[e];
    for (auto& ea : da.entries()) dw[ea.first.second * l + ea.first.first] = ea.second;
    if (dt != dw) ++fails;
    // add / sub (apply_elementwise, lib.rs:83-149; the reference's test: tests.rs:334-354): a second l x m
    // operand, compared through the dense DOK sums
    spam::DokMatrix<T> da2(l, m);
    for (uint64_t k = rng() % (2 * l * m + 1); k-- > 0;) da2.set_element(rng() % l, rng() % m, (T)((int)(rng() % 9) - 4));
    auto a2 = spam::CsrMatrix<T, true>::from(da2);
    auto s = a + a2;
    auto d = a - a2;
    if (!s.invariants() || !d.invariants()) { ++fails; continue; }
    std::vector<T> ws(l * m, T(0)), wd(l * m, T(0)), gs(l * m, T(0)), gd(l * m, T(0));
    for (auto& ea : da.entries()) { ws[ea.first.first * m + ea.first.second] += ea.second; wd[ea.first.first * m + ea.first.second] += ea.second; }
    for (auto& ea : da2.entries()) { ws[ea.first.first * m + ea.first.second] += ea.second; wd[ea.first.first * m + ea.first.second] -= ea.second; }
    for (uint64_t r = 0; r < l; ++r) {
      for (uint64_t e = s.offsets[r]; e < s.offsets[r + 1]; ++e) gs[r * m + s.indices[e]] = s.vals[e];
      for (uint64_t e = d.offsets[r]; e < d.offsets[r + 1]; ++e) gd[r * m + d.indices[e]] = d.vals[e];
    }
    if (ws != gs || wd != gd) ++fails;
  }
  std::printf("%s: %s\n", name, fails ? "FAIL" : "ok");
  return fails;
}

int main() {
  std::mt19937_64 rng(42);
  int fails = run<double>("f64", rng) + run<float>("f32", rng) + run<int32_t>("i32", rng) + run<int64_t>("i64", rng);
  // error behaviour: dimension mismatch throws (the reference panics out of bounds, mul_hash.rs:46)
  bool threw = false;
  try { auto a = spam::CsrMatrix<double, true>::identity(3); auto b = spam::CsrMatrix<double, true>::identity(4); (void)(a * b); }
  catch (const std::runtime_error&) { threw = true; }
  if (!threw) { std::printf("EDIM not raised\n"); ++fails; }
  bool idx = false;
  try { spam::DokMatrix<double> d(2, 2); d.set_element(2, 0, 1.0); } catch (const spam::IndexError&) { idx = true; }
  if (!idx) ++fails;
  // MatrixMarket ingest: symmetric integer file -> CsrMatrix<i64, true> (parse_matrix_market + From<DokMatrix>)
  {
    auto m = spam::from_matrix_market<int64_t>(
        "%%MatrixMarket matrix coordinate integer symmetric\n% c\n3 3 3\n1 1 4\n2 1 -1\n3 3 0\n3 2 7\n");
    const std::vector<uint64_t> off{0, 2, 4, 5}, idx2{0, 1, 0, 2, 1};
    const std::vector<int64_t> val{4, -1, -1, 7, 7};
    if (!m.invariants() || m.offsets != off || m.indices != idx2 || m.vals != val) { std::printf("matrix market mismatch\n"); ++fails; }
    bool threw2 = false;
    try { (void)spam::from_matrix_market<double>("%%MatrixMarket matrix coordinate pattern general\n2 2 0\n"); }
    catch (const std::runtime_error&) { threw2 = true; }
    if (!threw2) ++fails;
  }
  // hand-derived known answer (tests/golden/linprobe_kat.json, map_slot_order): A = [[1,2,3,4]], B = I4 -> 16 slots,
  // columns 0,1,2,3 land in slots 0,11,6,1: the B2 = false drain yields columns 0,3,2,1
  {
    spam::DokMatrix<double> da(1, 4), db(4, 4);
    for (int j = 0; j < 4; ++j) { da.set_element(0, j, 1.0 + j); db.set_element(j, j, 1.0); }
    auto a = spam::CsrMatrix<double, true>::from(da);
    auto b = spam::CsrMatrix<double, true>::from(db);
    auto c = a.template mul_hash<false>(b, spam::Handle::thread_default(), true);
    const std::vector<uint64_t> want_idx{0, 3, 2, 1};
    const std::vector<double> want_val{1.0, 4.0, 3.0, 2.0};
    if (c.indices != want_idx || c.vals != want_val) { std::printf("slot-order KAT mismatch\n"); ++fails; }
  }
  std::printf(fails ? "FAILED\n" : "ALL OK\n");
  return fails ? 1 : 0;
}
