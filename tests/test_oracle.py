"""CPU tests of the oracle (oracle/king_oracle.c) against the known-answer vectors and the reference's stated
semantics (/root/reference/cuking.cu:129-179 Submatrix, :675-703 pack, :191-314 kernel, :761-765 sort)."""
import json
import os
import struct

import numpy as np
import pytest

from oracle import king_oracle as ko
from tests.helpers import random_genotypes, triples_of, oracle_bitset

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load_kats():
    with open(os.path.join(GOLDEN, "kat_vectors.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("kat", load_kats(), ids=lambda k: k["id"])
def test_known_answer_vectors(kat):
    g = np.array([kat["genotypes_i"], kat["genotypes_j"]], dtype=np.int8)
    bs, sm = ko.pack_dense(g)
    c, kin = ko.pair_counts(bs, g.shape[1], 0, 1)
    for f in ("het_i", "het_j", "both_het", "opposing_hom", "concordant_hom", "shared_sites"):
        assert c[f] == kat[f], f
    if kat["kin_f32_hex"] == "nan":
        assert np.isnan(kin)
    else:
        assert struct.pack(">f", kin).hex() == kat["kin_f32_hex"]
    res, count, ovf = ko.king(bs, g.shape[1], sm, -1.0, 4)
    assert not ovf
    if kat["emitted_at_threshold_minus_1"]:
        assert count == 1
        assert (res[0]["sample_i"], res[0]["sample_j"]) == (0, 1)
        assert (res[0]["ibs0"], res[0]["ibs1"], res[0]["ibs2"]) == (kat["ibs0"], kat["ibs1"], kat["ibs2"])
    else:
        assert count == 0  # NaN / -inf never pass `kin > threshold` (cuking.cu:297)


def test_submatrix_enumerates_upper_triangle():
    # cuking.cu:136-144: row-major walk over block pairs (bi <= bj); every pair i<j is covered exactly once.
    for n, k in [(10, 1), (10, 3), (1000, 4), (7, 7), (64, 5)]:
        seen = np.zeros((n, n), dtype=np.int32)
        shards = k * (k + 1) // 2
        size = -(-n // k)
        expect = [(bi, bj) for bi in range(k) for bj in range(bi, k)]
        for shard in range(shards):
            sm = ko.submatrix(n, k, shard)
            bi, bj = expect[shard]
            if bi * size > n or bj * size > n:
                continue  # reference underflows here (SURVEY §8a4 latent bug); not part of the contract
            assert sm.i_begin == min(bi * size, n) and sm.j_begin == min(bj * size, n)
            for i in range(sm.i_begin, sm.i_end):
                for j in range(sm.j_begin, sm.j_end):
                    if i < j:
                        seen[i, j] += 1
        iu = np.triu_indices(n, 1)
        assert np.all(seen[iu] == 1)
        assert seen.sum() == len(iu[0])


def test_sample_offset_rows_then_cols():
    sm = ko.submatrix(1000, 4, 1)  # block (0, 1): rows [0,250), cols [250,500)
    L = ko.lib()
    import ctypes as C
    assert (sm.i_begin, sm.i_end, sm.j_begin, sm.j_end) == (0, 250, 250, 500)
    assert L.ko_num_samples(C.byref(sm)) == 500
    assert L.ko_sample_offset(C.byref(sm), 10) == 10
    assert L.ko_sample_offset(C.byref(sm), 260) == 250 + 10
    d = ko.submatrix(1000, 4, 4)  # block (1, 1) diagonal
    assert L.ko_num_samples(C.byref(d)) == 250


def test_geometry_padding():
    # cuking.cu:498-500 pads sites to x32; :513 words = 2 * ceil(S32 / 64)
    assert ko.words_per_sample(10_000) == 2 * 157
    assert ko.words_per_sample(100_000) == 2 * 1563
    assert ko.words_per_sample(1) == 2
    assert ko.words_per_sample(33) == 2
    assert ko.words_per_sample(65) == 4


def test_pack_and_semantics():
    # AND-accumulation (cuking.cu:687-697): duplicates are idempotent, conflicting triples AND together.
    sm = ko.submatrix(2)
    bs = ko.new_bitset(sm, 40)
    assert np.all(bs == np.uint64(0xFFFFFFFFFFFFFFFF))
    row = np.array([0, 0, 1, 1, 2, 3, 3], dtype=np.int64)
    col = np.array([0, 0, 0, 0, 0, 0, 1], dtype=np.int64)
    alt = np.array([1, 1, 1, 2, 0, 2, 2], dtype=np.int32)
    assert ko.pack(bs, 40, sm, row, col, alt) == -1
    wps = ko.words_per_sample(40)
    het0, alt0 = int(bs[0]), int(bs[wps // 2])
    assert (het0 >> 0) & 1 == 1 and (alt0 >> 0) & 1 == 0      # site 0: het (twice)
    assert (het0 >> 1) & 1 == 0 and (alt0 >> 1) & 1 == 0      # site 1: het AND hom-alt => hom-ref
    assert (het0 >> 2) & 1 == 0 and (alt0 >> 2) & 1 == 0      # site 2: hom-ref
    assert (het0 >> 3) & 1 == 0 and (alt0 >> 3) & 1 == 1      # site 3: hom-alt
    assert (het0 >> 4) & 1 == 1 and (alt0 >> 4) & 1 == 1      # site 4: untouched = missing
    # invalid genotype value -> index of the offending triple (cuking.cu:698-701)
    assert ko.pack(bs, 40, sm, np.array([0, 1]), np.array([0, 1]), np.array([0, 3])) == 1
    # samples outside the shard are skipped before validation (cuking.cu:677-679)
    assert ko.pack(bs, 40, sm, np.array([0]), np.array([5]), np.array([7])) == -1


def test_pack_order_independent():
    rng = np.random.default_rng(3)
    g = random_genotypes(rng, 9, 130)
    sm = ko.submatrix(9)
    site, sample, alt = triples_of(g)
    a = ko.new_bitset(sm, 130)
    ko.pack(a, 130, sm, site, sample, alt)
    perm = rng.permutation(site.size)
    b = ko.new_bitset(sm, 130)
    ko.pack(b, 130, sm, site[perm], sample[perm], alt[perm])
    assert np.array_equal(a, b)


def brute_counts(gi, gj):
    d = (gi >= 0) & (gj >= 0)
    return dict(
        het_i=int(np.sum((gi == 1) & d)), het_j=int(np.sum((gj == 1) & d)), both_het=int(np.sum((gi == 1) & (gj == 1))),
        opposing_hom=int(np.sum(d & (((gi == 0) & (gj == 2)) | ((gi == 2) & (gj == 0))))),
        concordant_hom=int(np.sum(d & (((gi == 0) & (gj == 0)) | ((gi == 2) & (gj == 2))))), shared_sites=int(d.sum()))


@pytest.mark.parametrize("n_sites", [1, 31, 32, 33, 63, 64, 65, 200, 1000])
def test_counts_match_genotype_level_definition(n_sites):
    rng = np.random.default_rng(n_sites)
    g = random_genotypes(rng, 6, n_sites, missing=0.1)
    bs, sm = ko.pack_dense(g)
    for i in range(6):
        for j in range(i + 1, 6):
            c, kin = ko.pair_counts(bs, n_sites, i, j)
            assert c == brute_counts(g[i], g[j])
            # counter identity the CUDA kernel relies on
            assert c["shared_sites"] == c["opposing_hom"] + c["concordant_hom"] + c["het_i"] + c["het_j"] - c["both_het"]


def test_king_threshold_overflow_and_sort():
    rng = np.random.default_rng(11)
    g = random_genotypes(rng, 40, 500)
    sm = ko.submatrix(40)
    bs = oracle_bitset(g, sm)
    allp, count, ovf = ko.king(bs, 500, sm, -1.0, 1000)
    assert count == 40 * 39 // 2 and not ovf
    assert np.all(allp["sample_i"] < allp["sample_j"])
    keys = allp["sample_i"].astype(np.int64) * 100 + allp["sample_j"]
    assert np.all(np.diff(keys) > 0)
    some, count2, _ = ko.king(bs, 500, sm, 0.1, 1000)
    assert 0 < count2 < count
    assert np.all(some["kin"] > np.float32(0.1))
    want = allp[allp["kin"] > np.float32(0.1)]
    assert np.array_equal(some, want)
    _, count3, ovf3 = ko.king(bs, 500, sm, -1.0, 10)
    assert ovf3 and count3 == count  # counter keeps counting past the capacity (cuking.cu:299-311)


def test_shards_union_equals_full():
    rng = np.random.default_rng(5)
    n, s, k = 37, 300, 3
    g = random_genotypes(rng, n, s)
    full, _, _ = ko.king(oracle_bitset(g, ko.submatrix(n)), s, ko.submatrix(n), 0.0, 4000)
    parts = []
    for shard in range(k * (k + 1) // 2):
        sm = ko.submatrix(n, k, shard)
        res, _, _ = ko.king(oracle_bitset(g, sm), s, sm, 0.0, 4000)
        parts.append(res)
    merged = np.concatenate(parts)
    ko.lib().ko_sort(merged.ctypes.data, merged.size)
    assert np.array_equal(merged, full)


def test_oracle_synth_cohort_equals_the_product_generator():
    # bench.py's --impl reference arm builds its inputs with the oracle's own restatement of the synthetic cohort so that
    # the reference process never loads the product library; the two generators must agree cell by cell
    import cuking_b200 as ck

    for seed, miss, s0, s1, r0, r1 in [(42, 0.01, 0, 64, 0, 500), (42, 0.05, 13, 150, 1000, 1300), (7, 0.0, 8, 16, 99_990, 100_010),
                                       (42, 1.0, 3, 5, 0, 40)]:
        a = ko.synth_genotypes(seed, miss, s0, s1, r0, r1)
        b = ck.synth_genotypes_host(seed, miss, s0, s1, r0, r1)
        assert np.array_equal(a, b), (seed, miss, s0, s1, r0, r1)
    g = ko.synth_genotypes(42, 0.02, 24, 77, 0, 333)
    bs = ko.synth_bitset(42, 0.02, 24, 77, 333)
    want, _ = ko.pack_dense(g)
    assert np.array_equal(bs, want)
