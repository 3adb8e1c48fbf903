"""The C-ABI shared library loads and exports every symbol include/mft_b200.h declares (no compute calls: CPU box)."""
import ctypes as C
import os
import re

import cases


def _mft():
    import mft_b200

    return mft_b200


def _declared():
    txt = open(os.path.join(cases.ROOT, "include", "mft_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mft_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported():
    m = _mft()
    lib = C.CDLL(m._lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mft_b200.h but not exported"
    assert set(m._lib.EXPORTS) == set(names)


def test_signatures_have_no_torch_types():
    txt = open(os.path.join(cases.ROOT, "include", "mft_b200.h")).read()
    assert "torch" not in txt.lower() and "at::" not in txt
    assert 'extern "C"' in txt


def test_no_cpu_fallback_without_device():
    m = _mft()
    lib = m._lib.load()
    if lib.mft_device_count() > 0:
        return
    ctx = C.c_void_p()
    rc = lib.mft_ctx_create(C.byref(ctx), 0, 100, 0, 4, 2, 20)
    assert rc == -3 and b"no CPU fallback" in lib.mft_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(cases.ROOT, "meshfreetrixi.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "mft_oracle" not in src and "oracle/" not in src, f
                # nor the tests' host-emulation harness (comments may name the directory; nothing may load or include it)
                assert "libmft_emu" not in src and "import emu" not in src and "emu_" not in src, f


import pytest


@pytest.mark.parametrize("R", [1, 2, 4])
@pytest.mark.parametrize("layout", [0, 1, 2, 3, 7])   # bit 0: bank-coloured slots, bit 1: two copies of the records, bit 2: tuned copy
@pytest.mark.parametrize("with_perm", [0, 1])
def test_union_tile_format_replays_bit_exact(R, layout, with_perm):
    """Host logic of the union-tile operator layout (mft_tile_kernels.cuh): the builder's step words, row masks,
    compact weight cursors and slot table, replayed on the CPU, reproduce the plain row sums bit for bit -- ragged rows,
    a trailing halo, tiles that end mid-slice; colouring / two copies lower the simulated LDS.128 bank-conflict degree."""
    m = _mft()
    lib = m._lib.load()
    for n, k in ((33, 3), (130, 5), (1000, 7), (4100, 20)):
        st = (C.c_double * 4)()
        rc = lib.mft_debug_tile_selftest(n, k, R, layout, with_perm, 1234 + n, st)
        assert rc == 0, lib.mft_last_error()
        assert st[3] < 4095
        if n == 4100:
            plain = (C.c_double * 4)()
            assert lib.mft_debug_tile_selftest(n, k, R, 0, with_perm, 1234 + n, plain) == 0
            assert st[0] <= plain[0] + 1e-12
            if layout and not (R == 4 and layout == 2):   # R = 4 has no copy-select bit
                assert st[0] < 0.95 * plain[0]


def test_union_tile_format_at_the_edge_sizes_of_the_gpu_tests():
    """the host layout builder at the cloud sizes tests/test_zz_ee_edge_sizes_gpu.py runs on hardware (below one slice, below
    one tile, straddling the 32-row and 128-row boundaries) and the stencil widths of the sweep"""
    m = _mft()
    lib = m._lib.load()
    st = (C.c_double * 4)()
    for n in (21, 26, 33, 45, 96, 128, 129, 140, 165, 220, 257, 285, 302):
        for k in (13, 15, 20, 30, 42):
            if k > n:
                continue
            for layout in (0, 3, 7):
                for perm in (0, 1):
                    assert lib.mft_debug_tile_selftest(n, k, 1, layout, perm, 7 + n, st) == 0, (n, k, layout, perm, lib.mft_last_error())


def test_union_tile_layout_does_not_depend_on_the_host_thread_count():
    """the layout builder runs on all host cores (MFT_HOST_THREADS); offsets come from a sizing pass and the generator of
    the second-copy permutations is jumped to each tile's position, so the bytes are the same for any thread count"""
    import re
    import subprocess
    import sys

    code = ("import sys, ctypes as C; sys.path.insert(0, %r); import mft_b200 as m; lib = m._lib.load(); st = (C.c_double * 4)()\n"
            "for n, k, R, layout, perm in ((20000, 20, 1, 3, 1), (9000, 13, 2, 7, 0), (7000, 30, 4, 1, 1), (300, 20, 1, 7, 0)):\n"
            "    assert lib.mft_debug_tile_selftest(n, k, R, layout, perm, 11, st) == 0\n") % cases.ROOT
    seen = []
    for threads in ("1", "3", "8"):
        env = dict(os.environ, MFT_TRACE="1", MFT_HOST_THREADS=threads)
        res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stderr
        sums = re.findall(r"fnv ([0-9a-f]{16})", res.stderr)
        assert len(sums) == 4, res.stderr
        seen.append(sums)
    assert seen[0] == seen[1] == seen[2]


def test_tuned_second_copy_lowers_the_conflict_degree_on_a_real_stencil_structure():
    """MFT_OPT_TILE bit 4 (layout only): on the kNN structure of the fixture cloud (forward operator and its transpose, ragged
    rows) the builder's local search brings the simulated LDS.128 conflict degree of the two-copy layout close to 1, the
    phase-1 stores become conflict-free, and the replayed row sums stay bit-identical."""
    import numpy as np
    import scipy.sparse as sp

    m = _mft()
    L = m._lib
    lib = L.load()
    fx = cases.fixture_setup()
    pts = fx["points"]
    order = L.sfc_order(pts)
    rank = np.empty(len(pts), dtype=np.int64)
    rank[order] = np.arange(len(pts))
    nb = rank[fx["nb"][order]]                      # the cloud renumbered along the curve, as the device sees it
    n, k = nb.shape
    A = sp.csr_matrix((np.ones(n * k), (np.repeat(np.arange(n), k), nb.reshape(-1))), shape=(n, n))
    A.sort_indices()
    AT = A.T.tocsr()
    AT.sort_indices()
    for M in (A, AT):
        rp = np.ascontiguousarray(M.indptr, dtype=np.int64)
        ci = np.ascontiguousarray(M.indices, dtype=np.int32)
        deg = {}
        for R in (1, 2):
            for layout in (0, 3, 7):
                st = (C.c_double * 6)()
                assert lib.mft_debug_tile_selftest_csr(n, n, L.ptr(rp), L.ptr(ci), R, layout, 9, st) == 0, lib.mft_last_error()
                deg[R, layout] = st[0]
                assert st[3] < 4095
                if layout == 7:      # octets of the union list are permutations of the bank groups: conflict-free phase-1 stores
                    assert st[4] <= 1.06 and st[5] <= 1.06
                elif layout == 3:
                    assert st[4] > 1.5 and st[5] > 1.5
        assert deg[1, 7] < deg[1, 3] < deg[1, 0] and deg[1, 7] < 1.12
        assert deg[2, 7] < deg[2, 3] < deg[2, 0]
    # argument checks
    assert lib.mft_debug_tile_selftest_csr(n, n + 1, L.ptr(rp), L.ptr(ci), 1, 3, 9, None) != 0
    bad = ci.copy()
    bad[5] = n
    assert lib.mft_debug_tile_selftest_csr(n, n, L.ptr(rp), L.ptr(bad), 1, 3, 9, None) != 0


def test_union_tile_layout_property_random_structures():
    """hypothesis: for random ragged sparsity patterns (empty rows, rows longer than a slice is wide, duplicate-free columns in
    arbitrary summation order, rows < columns), every layout variant replays to bit-identical row sums"""
    import numpy as np
    from hypothesis import given, settings, strategies as st_

    m = _mft()
    L = m._lib
    lib = L.load()

    @settings(max_examples=40, deadline=None)
    @given(st_.integers(1, 400), st_.integers(0, 60), st_.integers(0, 2 ** 31 - 1), st_.sampled_from([1, 2, 4]),
           st_.sampled_from([0, 1, 3, 7]), st_.floats(0.0, 1.0))
    def run(n, kmax, seed, R, layout, frac_rows):
        rng = np.random.default_rng(seed)
        n_rows = max(0, min(n, int(round(frac_rows * n))))
        lens = rng.integers(0, min(kmax, n) + 1, size=n_rows)
        rowptr = np.zeros(n_rows + 1, dtype=np.int64)
        rowptr[1:] = np.cumsum(lens)
        cols = np.concatenate([rng.choice(n, size=int(k), replace=False) for k in lens]) if n_rows and lens.sum() else np.zeros(0)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        if len(cols) == 0:
            cols = np.zeros(1, dtype=np.int32)
        stats = (C.c_double * 6)()
        rc = lib.mft_debug_tile_selftest_csr(n, n_rows, L.ptr(rowptr), L.ptr(cols), R, layout, seed % 1000, stats)
        assert rc == 0, (n, n_rows, kmax, R, layout, lib.mft_last_error())

    run()


# ---- the portable per-tile builder (csrc/mft_tile_build.cuh: the code the device layout kernels run) ----------------------------

def test_bank_group_feasibility_equals_the_augmenting_path_matcher():
    """The recursion-free two-choice test (union-find over the 8 bank groups: no component with more points than groups) gives
    the answer of the host builder's bipartite matcher on every instance with <= 3 points and on 2M random ones with <= 8."""
    m = _mft()
    lib = m._lib.load()
    bad = C.c_longlong(-1)
    assert lib.mft_debug_matcher_compare(7, 2_000_000, C.byref(bad)) == 0, lib.mft_last_error()
    assert bad.value == 0


def test_portable_tile_builder_is_byte_identical_to_the_host_builder():
    """Every array of the layout (step words, weights, slot tables of both copies, union lists, offsets, meta data) for random
    ragged banded operators: all layout variants, with and without a device permutation, k = 5 ... 50, a partial last tile."""
    m = _mft()
    lib = m._lib.load()
    for n, k, layout, perm in ((5000, 20, 7, 0), (5000, 20, 7, 1), (3000, 13, 3, 1), (3000, 30, 1, 0), (2000, 20, 0, 1), (700, 50, 7, 1),
                               (100, 5, 7, 0), (129, 20, 7, 1), (5000, 42, 7, 1), (4000, 15, 6, 1), (1, 1, 7, 0)):
        assert lib.mft_debug_tile_build_compare(n, k, layout, perm, 11) == 0, (n, k, layout, perm, lib.mft_last_error())
    assert lib.mft_debug_tile_build_compare(0, 20, 7, 0, 1) != 0   # argument checks
    assert lib.mft_debug_tile_build_compare(10, 20, 7, 0, 1) != 0


def test_portable_tile_builder_on_a_real_stencil_structure():
    """the kNN structure of the fixture cloud along the curve (forward operator: k entries per row; its transpose: ragged rows)"""
    import numpy as np
    import scipy.sparse as sp

    m = _mft()
    L = m._lib
    lib = L.load()
    fx = cases.fixture_setup()
    pts = fx["points"]
    order = L.sfc_order(pts)
    rank = np.empty(len(pts), dtype=np.int64)
    rank[order] = np.arange(len(pts))
    nb = rank[fx["nb"][order]]
    n, k = nb.shape
    A = sp.csr_matrix((np.ones(n * k), (np.repeat(np.arange(n), k), nb.reshape(-1))), shape=(n, n))
    A.sort_indices()
    AT = A.T.tocsr()
    AT.sort_indices()
    for M in (A, AT):
        rp = np.ascontiguousarray(M.indptr, dtype=np.int64)
        ci = np.ascontiguousarray(M.indices, dtype=np.int32)
        for layout in (0, 1, 3, 7):
            assert lib.mft_debug_tile_build_compare_csr(n, n, L.ptr(rp), L.ptr(ci), layout, 9) == 0, (layout, lib.mft_last_error())
        # fewer rows than columns (a partition's owned rows over owned + halo columns)
        nr = n - n // 5
        assert lib.mft_debug_tile_build_compare_csr(n, nr, L.ptr(rp[:nr + 1].copy()), L.ptr(ci), 7, 9) == 0, lib.mft_last_error()


def test_portable_tile_builder_property_random_structures():
    """hypothesis: random ragged sparsity patterns (empty rows, rows longer than a slice is wide, arbitrary summation order)"""
    import numpy as np
    from hypothesis import given, settings, strategies as st_

    m = _mft()
    L = m._lib
    lib = L.load()

    @settings(max_examples=40, deadline=None)
    @given(st_.integers(1, 400), st_.integers(0, 60), st_.integers(0, 2 ** 31 - 1), st_.sampled_from([0, 1, 3, 7, 6]), st_.floats(0.0, 1.0))
    def run(n, kmax, seed, layout, frac_rows):
        rng = np.random.default_rng(seed)
        n_rows = max(0, min(n, int(round(frac_rows * n))))
        lens = rng.integers(0, min(kmax, n) + 1, size=n_rows)
        rowptr = np.zeros(n_rows + 1, dtype=np.int64)
        rowptr[1:] = np.cumsum(lens)
        cols = np.concatenate([rng.choice(n, size=int(k), replace=False) for k in lens]) if n_rows and lens.sum() else np.zeros(0)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        if len(cols) == 0:
            cols = np.zeros(1, dtype=np.int32)
        rc = lib.mft_debug_tile_build_compare_csr(n, n_rows, L.ptr(rowptr), L.ptr(cols), layout, seed % 1000)
        assert rc == 0, (n, n_rows, kmax, layout, lib.mft_last_error())

    run()
