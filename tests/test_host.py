"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header declares,
the host mirror's invariants / DokMatrix semantics, the generators, and the loud failure without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import sparse_matrix_b200 as S
from sparse_matrix_b200 import _lib
from sparse_matrix_b200 import generators as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "spam_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spam_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from sparse_matrix_b200 import build
    build.build()
    L = ctypes.CDLL(S.SO_PATH)
    declared = _header_functions()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/spam_cuda.h but not exported"
    assert sorted(_lib.EXPORTS) == declared, "python binding list and header disagree"
    assert S.load().spam_cuda_abi_version() == 6


def test_rust_binding_and_build_recipes_agree_with_the_header():
    """rust/spam_cuda: every `extern "C"` name is declared in include/spam_cuda.h (and exported by the library), and
    build.rs compiles every .cu of csrc/ — the r1 recipe had fallen behind build.py and would not have linked."""
    declared = set(_header_functions())
    rs = open(os.path.join(ROOT, "rust", "spam_cuda", "src", "lib.rs")).read()
    ext = re.search(r'extern "C" \{(.*?)\n\}', rs, flags=re.S).group(1)
    rust_names = set(re.findall(r"pub fn (spam_[a-z0-9_]+)\s*\(", ext))
    assert len(rust_names) >= 15 and rust_names <= declared, rust_names - declared
    L = ctypes.CDLL(S.SO_PATH)
    assert all(hasattr(L, n) for n in rust_names)
    from sparse_matrix_b200 import build
    csrc = os.path.join(ROOT, "sparse_matrix_b200", "csrc")
    on_disk = sorted(f for f in os.listdir(csrc) if f.endswith(".cu"))
    assert build.SOURCES == on_disk and len(on_disk) >= 10
    brs = open(os.path.join(ROOT, "rust", "spam_cuda", "build.rs")).read()
    assert "read_dir" in brs and 'Some("cu")' in brs        # globs the directory instead of listing files


def test_strerror_covers_every_status():
    L = S.load()
    seen = set()
    for code in range(0, 10):
        msg = L.spam_strerror(code).decode()
        assert msg and msg != "unknown status"
        seen.add(msg)
    assert len(seen) == 10
    assert L.spam_strerror(99).decode() == "unknown status"


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(S.SpamError):
        S.Handle(0)
    a = S.CsrMatrix.identity(4)
    with pytest.raises(S.SpamError):
        a.mul_hash(a)
    with pytest.raises(S.SpamError):
        a.spmv(np.ones(4))
    with pytest.raises(S.SpamError):
        a.transpose()
    with pytest.raises(S.SpamError):
        a + a
    with pytest.raises(S.SpamError):
        S.CsrMatrix.from_triplets(2, 2, [0], [1], np.array([1.0]))
    # the MatrixMarket parser is host code and works without a device; building the matrix does not
    kind, r, c, tr, tc, tv = S.parse_matrix_market("%%MatrixMarket matrix coordinate real general\n2 2 1\n1 2 3.5\n")
    assert (kind, r, c, tr.tolist(), tc.tolist(), tv.tolist()) == ("real", 2, 2, [0], [1], [3.5])


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sparse_matrix_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "spam_oracle" not in text and "libspam_oracle" not in text, f


def test_invariants_mirror():
    M = S.CsrMatrix
    ok = M(2, 3, [1.0, 2.0, 3.0], [0, 2, 1], [0, 2, 3])
    assert ok.invariants()
    assert not M(2, 3, [1.0, 2.0, 3.0], [2, 0, 1], [0, 2, 3], is_sorted=True).invariants()   # not increasing
    assert M(2, 3, [1.0, 2.0, 3.0], [2, 0, 1], [0, 2, 3], is_sorted=False).invariants()      # distinct is enough
    assert not M(2, 3, [1.0, 2.0, 3.0], [2, 2, 1], [0, 2, 3], is_sorted=False).invariants()  # duplicate column
    assert not M(2, 3, [1.0, 2.0, 3.0], [0, 3, 1], [0, 2, 3]).invariants()                   # column out of range
    assert not M(2, 3, [1.0, 2.0, 3.0], [0, 2, 1], [0, 2, 2]).invariants()                   # offsets[rows] != nnz
    assert not M(2, 3, [1.0, 2.0, 3.0], [0, 2, 1], [1, 2, 3]).invariants()                   # offsets[0] != 0
    assert not M(2, 3, [1.0, 2.0], [0, 2, 1], [0, 2, 3]).invariants()                        # vals/indices length
    i = M.identity(5)
    assert i.invariants() and i.nnz() == 5 and i.get_element((3, 3)) == 1.0 and i.get_element((3, 2)) is None
    with pytest.raises(IndexError):
        i.get_element((5, 0))
    with pytest.raises(TypeError):
        M(1, 1, np.array([1 + 2j]), [0], [0, 1])   # Complex is not a device scalar
    assert M.new((3, 4)).invariants() and M.new((3, 4)).nnz() == 0


def test_dok_matrix_semantics():
    d = S.DokMatrix.new((3, 3))
    assert d.set_element((1, 1), 2.0) is None
    assert d.set_element((1, 1), 3.0) == 2.0          # replace returns the old value
    assert d.set_element((1, 1), 0.0) == 3.0          # zero removes (spam_dok lib.rs:171-175)
    assert d.nnz() == 0 and d.get_element((1, 1)) is None
    assert d.set_element((0, 2), 0.0) is None
    with pytest.raises(IndexError):
        d.set_element((3, 0), 1.0)
    d.set_element((2, 0), 1.0)
    d.set_element((0, 1), 4.0)
    assert [k for k, _ in d.iter()] == [(0, 1), (2, 0)] and d.invariants()


def test_generators_match_survey_counts():
    u = G.uniform_random(10_000, 10_000, 10, seed=1)
    assert u[0] == 10_000 and 99_000 < len(u[3]) <= 100_000
    p = G.poisson2d(64)
    n = 64
    assert len(p[3]) == 5 * n * n - 4 * n
    flops, per_row = G.spgemm_counts(p, p)
    assert per_row.min() == 11 and per_row.max() == 25
    # closed forms behind SURVEY §8's C2 numbers (n = 2048 gives 20 963 328 / 104 783 880)
    assert 5 * 2048 * 2048 - 4 * 2048 == 20_963_328
    s = G.stencil27(12)
    assert len(s[3]) == (3 * 12 - 2) ** 3
    fl, pr = G.spgemm_counts(s, s)
    assert fl == (9 * 12 - 10) ** 3           # per-dimension 9n-10 two-step paths (n=160: 1430^3)
    for m in (u, p, s):
        assert S.CsrMatrix(m[0], m[1], m[4], m[3], m[2]).invariants()
    r = G.rmat(10, 8)
    assert r[0] == 1024 and S.CsrMatrix(r[0], r[1], r[4], r[3], r[2]).invariants()


def test_transpose_and_triplet_stream(oracle):
    a = G.uniform_random(300, 700, 6, seed=3, dtype=np.int64, int_range=1000)
    at = G.transpose(a)
    assert at[0] == 700 and at[1] == 300 and len(at[3]) == len(a[3])
    att = G.transpose(at)
    for x, y in zip(a, att):
        assert np.array_equal(x, y)
    tr, tc, tv = G.triplets_with_rewrites(a, seed=5, dup_frac=0.05, zero_frac=0.02)
    off, idx, val = oracle.dok_to_csr(300, 700, tr, tc, tv)
    nzero = int(len(a[3]) * 0.02)
    assert len(idx) == len(a[3]) - nzero           # the zero writes deleted exactly that many entries
    assert np.all(np.diff(off.astype(np.int64)) <= np.diff(a[2].astype(np.int64)))


def test_bucket_path_plan_on_the_host(tmp_path):
    """The plan of the DOK -> CSR / transpose bucket path (bucket.cuh: bk_plan — bucket width, bucket count, key packing,
    shared-memory budgets) is host code: compiled with nvcc and run on the CPU (no kernel is launched) over the C5 shapes,
    the shapes the path must refuse, and 200 000 random shapes checked against its invariants."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    src = os.path.join(os.path.dirname(__file__), "cpp", "test_bk_plan.cu")
    exe = str(tmp_path / "test_bk_plan")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-o", exe, src],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "bk_plan ok" in out.stdout, out.stdout + out.stderr
