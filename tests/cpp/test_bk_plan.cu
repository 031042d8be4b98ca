// Host-only check of the bucket path's plan (sparse_matrix_b200/csrc/bucket.cuh: bk_plan, bk_part_smem, bk_build_smem):
// compiled with nvcc and run on the CPU by tests/test_host.py — no device code is launched.
#include <cstdio>
#include <cstdlib>
#include "../../sparse_matrix_b200/csrc/bucket.cuh"

static int fails = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL line %d: %s\n", __LINE__, #c); ++fails; } } while (0)

static void check_plan(u64 majors, u64 minors, u64 n) {
  BkPlan p;
  if (!bk_plan(majors, minors, n, &p)) return;
  CHECK(p.shift >= 0 && p.shift <= BK_SHIFT_MAX);
  CHECK(p.mbits >= 0 && p.mbits + p.shift <= 32);            // (major_local, minor) packs into one u32
  CHECK(minors <= (1ull << p.mbits));
  CHECK(p.nb >= 1 && p.nb <= BK_NB_MAX);
  CHECK(((u64)p.nb << p.shift) >= majors && (((u64)p.nb - 1) << p.shift) < majors);  // the buckets tile the majors
  CHECK(n * 100 <= (u64)BK_CAP * 85 * p.nb);                 // average fill <= 85 %
  CHECK(bk_build_smem(p.shift) * 3 + 3 * 1024 <= 228 * 1024);  // three build blocks per SM
  CHECK(bk_part_smem(p.nb) <= 227 * 1024);                   // one partition block always fits
  if (p.shift < BK_SHIFT_MAX && p.mbits + p.shift < 32) {    // a wider bucket would have been too full (or too few)
    const u64 nb2 = (majors + (2ull << p.shift) - 1) >> (p.shift + 1);
    CHECK(n * 100 > (u64)BK_CAP * 85 * nb2);
  }
}

int main() {
  BkPlan p;
  // C5: DOK -> CSR (rows 1 M, cols 4 M, 8.09 M triplets) and its transpose (majors = 4 M columns, minors = 1 M rows)
  CHECK(bk_plan(1000000, 4000000, 8087988, &p) && p.shift == 9 && p.nb == 1954 && p.mbits == 22);
  CHECK(bk_plan(4000000, 1000000, 7999990, &p) && p.shift == 11 && p.nb == 1954 && p.mbits == 20);
  // shapes the path does not take
  CHECK(!bk_plan(0, 10, 10, &p));
  CHECK(!bk_plan(10, 10, 0, &p));
  CHECK(bk_plan(1ull << 24, 1ull << 10, 1000, &p) && p.nb == 8192 && p.shift == 11);
  CHECK(!bk_plan((1ull << 24) + 1, 1ull << 10, 1000, &p));    // more than 8192 buckets even at 2048 majors each
  CHECK(!bk_plan(1000, 1ull << 33, 1000, &p));                // minors do not fit 32 bits
  CHECK(!bk_plan(4194304, 4194304, 67094166, &p));            // R-MAT 22: too many entries for 8192 buckets
  CHECK(bk_plan(3000, (1ull << 31) - 2, 13500, &p) && p.shift == 1);  // the fuzz target's column space: 2 rows per bucket
  CHECK(bk_plan(1, 70000, 300, &p) && p.nb == 1);
  unsigned long long x = 88172645463325252ull;
  auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
  for (int i = 0; i < 200000; ++i) {
    const u64 majors = 1 + rnd() % (1ull << (1 + rnd() % 26));
    const u64 minors = 1 + rnd() % (1ull << (1 + rnd() % 33));
    const u64 n = 1 + rnd() % (1ull << (1 + rnd() % 27));
    check_plan(majors, minors, n);
  }
  if (fails) return 1;
  std::printf("bk_plan ok\n");
  return 0;
}
