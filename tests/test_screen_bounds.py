"""CPU check of the inequalities behind variant 5's screens (king_screen1_kernel.cu, king_screen_kernel.cu),
independent of any kernel: for every pair of random cohorts - with heavy, lopsided missingness, duplicates, relatives, samples
without hets - the oracle's kinship (cuking.cu:289-294 in fp32) never exceeds what either bound allows, under exactly the
fp32 test the kernels evaluate.  If this held only approximately the screens could drop a retained pair."""
import numpy as np
import pytest

from oracle import king_oracle as ko
from tests.helpers import random_genotypes, oracle_bitset


def oracle_kin_matrix(g):
    n, s = g.shape
    sm = ko.submatrix(n)
    res, count, _ = ko.king(oracle_bitset(g, sm), s, sm, -1.0, n * n)
    kin = np.full((n, n), -np.inf, dtype=np.float32)  # pairs the reference never emits (NaN / -inf) stay at -inf
    kin[res["sample_i"], res["sample_j"]] = res["kin"]
    return kin


def bounds(g):
    """half_d3 = D / 2 exactly (three products); lower1 = the one-product lower bound on D / 2; het = het totals."""
    called = g >= 0
    x = np.where(g == 2, 1.0, np.where(g == 0, -1.0, 0.0))          # +1 hom-alt, -1 hom-ref
    y = (called & (g != 1)).astype(np.float64)                      # hom
    h = 0.5 * (g == 1)                                              # het, stored as 0.5
    w = y + h
    half_d3 = y @ w.T + h @ y.T - x @ x.T                           # = (2 YY + YH + HY - 2 xx) / 2
    het = (g == 1).sum(axis=1).astype(np.float64)
    hom = y.sum(axis=1)
    defined = het + hom
    s = float(g.shape[1])
    t = np.maximum(0.0, hom[:, None] + defined[None, :] - s)        # >= 0 part of  Y_i + Def_j - S
    lower1 = 0.5 * (t + t.T) - x @ x.T
    return half_d3, lower1, het


@pytest.mark.parametrize("n,s,missing", [(60, 500, 0.0), (80, 900, 0.03), (70, 1500, 0.15), (50, 700, 0.4), (40, 64, 0.1)])
def test_no_retained_pair_escapes_either_bound(n, s, missing):
    rng = np.random.default_rng(n * s)
    g = random_genotypes(rng, n, s, missing=missing)
    g[1] = g[0]                                   # a duplicate: kin = 0.5 exactly where het counts allow
    g[5][g[5] == 1] = 0                           # no hets
    g[7][rng.random(s) < 0.7] = -1                # mostly missing
    kin = oracle_kin_matrix(g)
    half_d3, lower1, het = bounds(g)
    # the exact distance: D = het_i' + het_j' - 2 both_het + 4 opp >= 0, and lower1 really is a lower bound
    assert np.all(half_d3 >= -1e-9) and np.all(lower1 <= half_d3 + 1e-9)
    min_het = np.minimum(het[:, None], het[None, :])
    iu = np.triu_indices(n, 1)
    for thr in (-0.3, 0.0, 0.0442, 0.0884, 0.2, 0.35, 0.49, 0.5, 0.6):
        bound2 = np.float32(2.0) * (np.float32(0.5) - np.float32(thr)) * np.float32(1.0001)
        rhs = np.float32(bound2) * min_het.astype(np.float32) + np.float32(1.0)   # fmaf(bound2, min(Het), 1.f) up to one rounding
        cand3 = half_d3.astype(np.float32) < rhs
        cand1 = lower1.astype(np.float32) < rhs
        retained = kin > np.float32(thr)          # the reference's strict fp32 comparison (cuking.cu:297)
        assert not np.any(retained[iu] & ~cand3[iu]), ("three-product screen would drop a pair", thr)
        assert not np.any(retained[iu] & ~cand1[iu]), ("one-product screen would drop a pair", thr)
        assert not np.any(cand3[iu] & ~cand1[iu])  # the one-product bound is the looser of the two
