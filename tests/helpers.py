"""Shared test utilities: seeded genotype matrices, triples, oracle runs."""
from __future__ import annotations

import ctypes as C

import numpy as np

from oracle import king_oracle as ko


def random_genotypes(rng, n_samples, n_sites, missing=0.03, af_lo=0.05, af_hi=0.5, related_blocks=True):
    """int8 [samples, sites] in {0,1,2} / -1; optional planted duplicates + parent/child so thresholds keep pairs."""
    af = rng.uniform(af_lo, af_hi, size=n_sites)
    h0 = rng.random((n_samples, n_sites)) < af
    h1 = rng.random((n_samples, n_sites)) < af
    if related_blocks and n_samples >= 4:
        for c in range(2, n_samples, 4):  # sample c is a child of (c-2, c-1)
            sel0 = rng.random(n_sites) < 0.5
            sel1 = rng.random(n_sites) < 0.5
            pa0, pa1 = h0[c - 2].copy(), h1[c - 2].copy()
            pb0, pb1 = h0[c - 1].copy(), h1[c - 1].copy()
            h0[c] = np.where(sel0, pa0, pa1)
            h1[c] = np.where(sel1, pb0, pb1)
    g = (h0.astype(np.int8) + h1.astype(np.int8)).astype(np.int8)
    g[rng.random((n_samples, n_sites)) < missing] = -1
    return g


def triples_of(g: np.ndarray, order="site"):
    """(row_idx=site, col_idx=sample, n_alt) of the non-missing entries; Hail order = site-major, sample-minor."""
    if order == "site":
        site, sample = np.nonzero(g.T >= 0)
    else:
        sample, site = np.nonzero(g >= 0)
    return site.astype(np.int64), sample.astype(np.int64), g[sample, site].astype(np.int32)


def oracle_bitset(g: np.ndarray, sm: ko.Submatrix):
    site, sample, alt = triples_of(g)
    bs = ko.new_bitset(sm, g.shape[1])
    assert ko.pack(bs, g.shape[1], sm, site, sample, alt) == -1
    return bs


def ko_sm(ck_sm) -> ko.Submatrix:
    return ko.Submatrix(ck_sm.i_begin, ck_sm.i_end, ck_sm.j_begin, ck_sm.j_end)


def bits_equal_f32(a: np.ndarray, b: np.ndarray) -> bool:
    return np.array_equal(np.asarray(a, dtype=np.float32).view(np.uint32), np.asarray(b, dtype=np.float32).view(np.uint32))


def assert_results_equal(got: np.ndarray, want: np.ndarray):
    assert got.shape == want.shape, (got.shape, want.shape)
    for f in ("sample_i", "sample_j", "ibs0", "ibs1", "ibs2"):
        assert np.array_equal(got[f], want[f]), f
    assert bits_equal_f32(got["kin"], want["kin"]), "kin differs bitwise"
