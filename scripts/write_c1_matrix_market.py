#!/usr/bin/env python
"""Writes BASELINE.json configs[0] (uniform random CSR 10k x 10k, ~10 nnz per row, f64, seed 1) as a MatrixMarket
`coordinate real general` file, so that the reference's own bench (spam_csr/benches/mul_hash.rs, which multiplies every
file of ./matrices/ by itself: spam_csr/src/lib.rs:419-431) can consume the same input the GPU path is measured on.

  python scripts/write_c1_matrix_market.py [matrices/uniform10k.mtx]

The writer is the restatement of `into_float_matrix_market` (spam_dok/src/lib.rs:480-489); tests/test_matrix_market.py
round-trips a file written by it through the parser and the device DOK -> CSR build.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import sparse_matrix_b200 as S  # noqa: E402
from sparse_matrix_b200 import generators as G  # noqa: E402


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join("matrices", "uniform10k.mtx")
    m = G.uniform_random(10_000, 10_000, 10, seed=1)
    a = S.CsrMatrix(m[0], m[1], m[4], m[3], m[2])
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    with open(out, "w") as f:
        f.write(S.into_float_matrix_market(a))
    print(f"{out}: {a.rows()} x {a.cols()}, {a.nnz()} entries")


if __name__ == "__main__":
    main()
