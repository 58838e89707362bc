"""CPU tests of the C-ABI library: it loads, exports every symbol the header declares, and its host-side logic
(shard planning, geometry, generator, argument validation) agrees with the oracle.  No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import cuking_b200 as ck
from cuking_b200 import capi
from oracle import king_oracle as ko

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    with open(os.path.join(ROOT, "include", "cuking_b200.h")) as f:
        header = f.read()
    declared = sorted(set(re.findall(r"\b(ck_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    L = C.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/cuking_b200.h but not exported"
    assert sorted(capi.EXPORTED_SYMBOLS) == declared
    assert capi.load().ck_abi_version() == 3


def test_struct_layouts():
    assert capi.RESULT_DTYPE.itemsize == 24 and ko.RESULT_DTYPE == capi.RESULT_DTYPE  # cuking.cu:182-186
    assert C.sizeof(capi.Submatrix) == 16 and C.sizeof(capi.Counts) == 24


@pytest.mark.parametrize("n,k", [(1, 1), (10, 1), (10, 3), (1000, 4), (7, 7), (300_000, 4), (1_000_000, 8), (65, 2)])
def test_submatrix_matches_oracle(n, k):
    L = capi.load()
    size = -(-n // k)
    for shard in range(k * (k + 1) // 2):
        a = ck.submatrix(n, k, shard)
        b = ko.submatrix(n, k, shard)
        bi = next(i for i in range(k) if shard < sum(k - t for t in range(i + 1)))
        if bi * size > n or (b.j_begin > n):
            assert a.i_begin <= a.i_end and a.j_begin <= a.j_end  # guarded instead of underflowing
            continue
        assert (a.i_begin, a.i_end, a.j_begin, a.j_end) == (b.i_begin, b.i_end, b.j_begin, b.j_end)
        assert L.ck_submatrix_num_samples(C.byref(a)) == ko.lib().ko_num_samples(C.byref(b))
        for s in {a.i_begin, a.i_end - 1, a.j_begin, a.j_end - 1, 0, n - 1, n // 2}:
            if s < 0 or s >= n:
                continue
            assert bool(L.ck_submatrix_contains(C.byref(a), s)) == bool(ko.lib().ko_contains(C.byref(b), s))
            if L.ck_submatrix_contains(C.byref(a), s):
                assert L.ck_submatrix_sample_offset(C.byref(a), s) == ko.lib().ko_sample_offset(C.byref(b), s)


def test_flag_validation_errors():
    # cuking.cu:454-462
    with pytest.raises(ck.CukingError, match="Invalid split factor") as e:
        ck.submatrix(10, 0, 0)
    assert e.value.code == capi.CK_ERR_INVALID_ARGUMENT
    with pytest.raises(ck.CukingError, match="Invalid shard index"):
        ck.submatrix(10, 4, 10)
    assert ck.num_shards(4) == 10 and ck.num_shards(1) == 1 and ck.num_shards(8) == 36


@pytest.mark.parametrize("s", [1, 31, 32, 33, 63, 64, 65, 10_000, 100_000, 1_000_000])
def test_words_per_sample(s):
    assert ck.words_per_sample(s) == ko.words_per_sample(s)


def test_synth_generator_is_a_pure_function():
    a = ck.synth_genotypes_host(42, 0.02, 0, 24, 0, 500)
    b = ck.synth_genotypes_host(42, 0.02, 8, 16, 100, 300)
    assert np.array_equal(a[8:16, 100:300], b)  # any cell can be regenerated anywhere
    c = ck.synth_genotypes_host(42, 0.02, 5, 11, 7, 90)  # ranges that cut pedigree blocks
    assert np.array_equal(a[5:11, 7:90], c)
    assert set(np.unique(a)) <= {-1, 0, 1, 2}
    miss = np.mean(ck.synth_genotypes_host(42, 0.05, 0, 64, 0, 4000) < 0)
    assert 0.04 < miss < 0.06
    assert not np.array_equal(a, ck.synth_genotypes_host(43, 0.02, 0, 24, 0, 500))


def test_synth_pedigree_has_planted_relatedness():
    g = ck.synth_genotypes_host(42, 0.0, 0, 16, 0, 20_000)
    bs, sm = ko.pack_dense(g)
    kin = lambda i, j: ko.pair_counts(bs, g.shape[1], i, j)[1]
    assert 0.2 < kin(0, 2) < 0.3     # parent-child
    assert 0.2 < kin(2, 3) < 0.3     # full sibs
    assert 0.08 < kin(0, 5) < 0.17   # grandparent
    assert 0.03 < kin(0, 7) < 0.10   # great-grandparent
    assert abs(kin(0, 1)) < 0.04     # founders
    assert abs(kin(0, 8)) < 0.04     # different pedigree blocks


def test_missing_library_or_device_fails_loudly():
    L = capi.load()
    n = C.c_int(-1)
    rc = L.ck_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(ck.CukingError):
        ck.Context(0)  # no CPU fallback
