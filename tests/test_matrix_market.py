"""MatrixMarket ingest: the host parser spam_mm_parse against hand-derived known answers
for every rule of the reference's nom grammar (spam_dok/src/lib.rs:282-478), against scipy.io as an
independent reader, and a write/parse round trip of `into_float_matrix_market` (lib.rs:480-489).
The parser is host code inside libspam_cuda.so: these tests need the built library but no GPU."""
import io
import os

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

import sparse_matrix_b200 as S
from sparse_matrix_b200 import generators as G

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def as_dict(tr, tc, tv):
    d = {}
    for r, c, v in zip(tr.tolist(), tc.tolist(), tv.tolist()):
        d[(r, c)] = v          # BTreeMap::insert: a later entry replaces
    return d


def test_golden_general_real():
    kind, rows, cols, tr, tc, tv = S.parse_matrix_market(open(os.path.join(GOLD, "general_real.mtx"), "rb").read())
    assert (kind, rows, cols) == ("real", 4, 5)
    # comments skipped; zero entries (0, 0.0, -0.0, 0e5) skipped; 1-based -> 0-based; duplicate (2,3) kept in
    # stream order so that the later one wins; forms of nom's recognize_float: 1., .5, -2.5e-1, +3, 1E2
    assert list(zip(tr.tolist(), tc.tolist(), tv.tolist())) == [
        (0, 0, 1.0), (1, 2, 7.0), (3, 4, 0.5), (1, 2, -0.25), (2, 0, 3.0), (0, 4, 100.0)]
    assert as_dict(tr, tc, tv)[(1, 2)] == -0.25


def test_golden_symmetric_integer_crlf():
    kind, rows, cols, tr, tc, tv = S.parse_matrix_market(open(os.path.join(GOLD, "symmetric_integer_crlf.mtx"), "rb").read())
    assert (kind, rows, cols) == ("integer", 3, 3) and tv.dtype == np.int64
    # symmetric: (r, c) then (c, r) for every non-zero entry, the diagonal twice as well (same key, same value)
    assert list(zip(tr.tolist(), tc.tolist(), tv.tolist())) == [
        (0, 0, 4), (0, 0, 4), (1, 0, -1), (0, 1, -1), (2, 1, 9223372036854775807), (1, 2, 9223372036854775807)]


def test_entry_list_ends_silently_at_the_first_line_that_does_not_match():
    head = "%%MatrixMarket matrix coordinate real general\n3 3 9\n"
    for tail, n in (("1 1 1.0\n2 2 2.0", 1),                 # no EOL after the last entry: dropped
                    ("1 1 1.0\n 2 2 2.0\n3 3 3.0\n", 1),      # leading blank: stop, the rest is ignored
                    ("1 1 1.0\n2  2 2.0\n", 1),               # double space
                    ("1 1 1.0\n-2 2 2.0\n3 3 3.0\n", 1),      # '-' never parses as usize
                    ("1 1 1.0\n2 2 2.0 \n", 1),               # trailing blank
                    ("1 1 1.0\n2 2 nan\n", 1),                # recognize_float has no nan / inf
                    ("1 1 1.0\n2 2\n3 3 1\n", 1),             # value missing
                    ("1 1 1.0\n\n2 2 2.0\n", 1),              # empty line
                    ("", 0)):
        kind, rows, cols, tr, tc, tv = S.parse_matrix_market(head + tail)
        assert len(tv) == n and (rows, cols) == (3, 3), tail
    # integer files: a value that does not fit i64 ends the list; '+' is not part of recognize_int
    head = "%%MatrixMarket matrix coordinate integer general\n3 3 9\n"
    assert len(S.parse_matrix_market(head + "1 1 5\n2 2 9223372036854775808\n3 3 1\n")[5]) == 1
    assert S.parse_matrix_market(head + "1 1 -9223372036854775808\n")[5].tolist() == [-9223372036854775808]
    assert len(S.parse_matrix_market(head + "1 1 +5\n")[5]) == 0
    assert len(S.parse_matrix_market(head + "1 1 5.0\n")[5]) == 0
    # the declared entry count is not used, and indices beyond the shape are not the parser's business
    assert len(S.parse_matrix_market(head + "1 1 5\n2 2 6\n")[5]) == 2
    assert S.parse_matrix_market(head + "7 9 5\n")[3].tolist() == [6]


def test_errors():
    E = S.FromMatrixMarketError
    ok = "%%MatrixMarket matrix coordinate real general\n2 2 1\n1 1 1.0\n"
    assert len(S.parse_matrix_market(ok)[5]) == 1
    for bad in ("%MatrixMarket matrix coordinate real general\n2 2 1\n",           # header tag
                "%%MatrixMarket matrix array real general\n2 2\n",                 # only `coordinate`
                "%%MatrixMarket matrix coordinate  real general\n2 2 1\n",         # double space
                "%%MatrixMarket matrix coordinate real general \n2 2 1\n",         # trailing blank
                "%%MatrixMarket matrix coordinate double general\n2 2 1\n",        # entry type
                "%%MatrixMarket matrix coordinate real general\n2 2\n",            # size line needs three numbers
                "%%MatrixMarket matrix coordinate real general\n2 -2 1\n",
                "%%MatrixMarket matrix coordinate real general\n2 2 1",            # size line needs its EOL
                "%%MatrixMarket matrix coordinate real general\n\n2 2 1\n",        # blank line is not a comment
                "%%MatrixMarket matrix coordinate real general\n2 2 1\n1 1 1e\n",  # nom `cut`: exponent without digits
                "%%MatrixMarket matrix coordinate real general\n0 2 0\n",          # HasZeroDimension
                "%%MatrixMarket matrix coordinate real general\n2 0 0\n"):
        with pytest.raises(E):
            S.parse_matrix_market(bad)
    for todo in ("complex general", "pattern general", "real skew-symmetric", "integer hermitian"):
        with pytest.raises(E):     # todo!() in the reference (and complex has no device scalar)
            S.parse_matrix_market(f"%%MatrixMarket matrix coordinate {todo}\n2 2 0\n")
    with pytest.raises(IndexError):   # r - 1 underflows in the reference
        S.parse_matrix_market("%%MatrixMarket matrix coordinate real general\n2 2 1\n0 1 1.0\n")
    # a zero entry at index 0 is skipped before the index is touched (general(), lib.rs:332-340)
    assert len(S.parse_matrix_market("%%MatrixMarket matrix coordinate real general\n2 2 1\n0 1 0.0\n")[5]) == 0


@pytest.mark.filterwarnings("ignore::DeprecationWarning")
@pytest.mark.parametrize("symmetric", [False, True])
def test_against_scipy_reader(symmetric):
    rng = np.random.default_rng(8)
    for rows, cols, dens in ((7, 7, 0.4), (40, 40, 0.1), (1, 1, 1.0)):
        m = sp.random(rows, cols, density=dens, random_state=rng, format="coo")
        if symmetric:
            m = sp.coo_matrix(sp.tril(m + m.T))
        buf = io.BytesIO()
        scipy.io.mmwrite(buf, m, symmetry="symmetric" if symmetric else "general")
        text = buf.getvalue()
        kind, r, c, tr, tc, tv = S.parse_matrix_market(text)
        want = scipy.io.mmread(io.BytesIO(text)).todok()
        got = as_dict(tr, tc, tv)
        assert (kind, r, c) == ("real", rows, cols) and len(got) == want.nnz
        for (i, j), v in got.items():
            assert want[i, j] == v


def test_writer_round_trip():
    """into_float_matrix_market -> parse_matrix_market gives the entries back bit for bit: Rust's `{}` prints the
    shortest digits that round-trip, without exponent."""
    rng = np.random.default_rng(9)
    vals = np.concatenate([rng.uniform(-1, 1, 20), [1.0, -2.0, 1e21, 1e-7, 123456789.125, 5e-324, 1.7976931348623157e308]])
    n = len(vals)
    m = S.CsrMatrix(n, n, vals, np.arange(n, dtype=np.uint64), np.arange(n + 1, dtype=np.uint64))
    text = S.into_float_matrix_market(m)
    assert text.splitlines()[0] == "%%MatrixMarket matrix coordinate real general" and "e" not in text.split("\n", 2)[2]
    assert "\n21 21 1\n" in text and "\n23 23 1000000000000000000000\n" in text
    kind, r, c, tr, tc, tv = S.parse_matrix_market(text)
    assert (kind, r, c) == ("real", n, n)
    assert np.array_equal(tr, np.arange(n)) and np.array_equal(tc, np.arange(n)) and np.array_equal(tv, vals)


@pytest.mark.gpu
def test_file_to_device_csr(oracle, handle, tmp_path):
    """The whole ingest: file -> triplets -> DOK -> CSR on the device, against the oracle's BTreeMap build."""
    rng = np.random.default_rng(10)
    rows, cols = 300, 200
    lines = ["%%MatrixMarket matrix coordinate integer symmetric", "% generated", f"{rows} {rows} 0"]
    for _ in range(3000):
        r, c = int(rng.integers(1, rows + 1)), int(rng.integers(1, rows + 1))
        lines.append(f"{r} {c} {int(rng.integers(-3, 4))}")           # zeros and repeated keys included
    p = tmp_path / "m.mtx"
    p.write_text("\n".join(lines) + "\n")
    got = S.load_matrix_market(str(p), handle=handle)
    kind, r, c, tr, tc, tv = S.parse_matrix_market(p.read_bytes())
    off, idx, val = oracle.dok_to_csr(r, c, tr, tc, tv)
    assert kind == "integer" and got.invariants()
    assert np.array_equal(got.offsets, off) and np.array_equal(got.indices, idx) and np.array_equal(got.vals, val)
    assert all(got.get_element((j, i)) == v for (i, j), v in list(got.iter())[:200])   # symmetric
    _ = cols


def test_config1_written_for_the_reference_bench(tmp_path):
    """scripts/write_c1_matrix_market.py: BASELINE configs[0] as the file the reference's bench reads from ./matrices
    (spam_csr/src/lib.rs:419-431).  The file parses back to exactly the generated matrix (values round-trip through
    Rust's `{}` float format)."""
    import subprocess
    import sys
    out = tmp_path / "matrices" / "uniform10k.mtx"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call([sys.executable, os.path.join(root, "scripts", "write_c1_matrix_market.py"), str(out)])
    kind, r, c, tr, tc, tv = S.parse_matrix_market(out.read_bytes())
    m = G.uniform_random(10_000, 10_000, 10, seed=1)
    rows_of = np.repeat(np.arange(m[0], dtype=np.uint64), np.diff(m[2]).astype(np.int64))
    assert (kind, r, c) == ("real", 10_000, 10_000)
    assert np.array_equal(tr, rows_of) and np.array_equal(tc, m[3]) and np.array_equal(tv, m[4])
