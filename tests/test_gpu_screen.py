"""Variant 5: the screen kernels (king_screen1_kernel.cu: one product, 128 x 160 tiles; king_screen_kernel.cu: three
products) in front of the mxf4 kernel.  Every record still comes out of the five-product kernel, so the output must be the
oracle's bit for bit whichever screen runs and whatever it lets through; what these cases stress is that the screens never
drop a pair: thresholds right at a pair's kinship, heavy and lopsided missingness, samples without hets, relatedness
scattered over the whole matrix, tile ranges that start and end anywhere."""
import os

import numpy as np
import pytest

import cuking_b200 as ck
from oracle import king_oracle as ko
from tests.helpers import random_genotypes, triples_of, oracle_bitset, ko_sm, assert_results_equal

pytestmark = pytest.mark.gpu


def screen_ctx(level):
    """A ctx whose variant 5 always runs the given screen (CUKING_SCREEN_LEVEL is read at ck_ctx_create); None = automatic."""
    old = os.environ.pop("CUKING_SCREEN_LEVEL", None)
    if level is not None:
        os.environ["CUKING_SCREEN_LEVEL"] = str(level)
    try:
        return ck.Context(0, king_variant=5)
    finally:
        os.environ.pop("CUKING_SCREEN_LEVEL", None)
        if old is not None:
            os.environ["CUKING_SCREEN_LEVEL"] = old


def scattered_relatives(rng, n, s, missing):
    """Unrelated founders plus duplicates / parent-child pairs placed far apart, so that candidates sit in many tiles."""
    g = random_genotypes(rng, n, s, missing=0.0, related_blocks=False)
    for _ in range(n // 6):
        a, b = rng.integers(0, n, 2)
        if a == b:
            continue
        if rng.random() < 0.5:
            g[b] = g[a]  # duplicate / twin
        else:  # child of a and a random other sample: one allele from each
            c = int(rng.integers(0, n))
            g[b] = (rng.random(s) < g[a] / 2).astype(np.int8) + (rng.random(s) < g[c] / 2).astype(np.int8)
    miss = rng.random((n, s)) < missing
    if n > 8:
        miss[7] = rng.random(s) < 0.6      # a sample that is mostly missing
        g[5][g[5] == 1] = 0                # no hets at all: min_hets == 0 where it matters
    g[miss] = -1
    return g


@pytest.mark.parametrize("level", [1, 3, None])
@pytest.mark.parametrize("n,s,k,shard,missing", [(300, 2000, 1, 0, 0.01), (700, 1200, 1, 0, 0.10), (900, 640, 2, 1, 0.03),
                                                 (1300, 333, 3, 4, 0.0), (260, 4100, 1, 0, 0.30)])
def test_screens_never_drop_a_pair(level, n, s, k, shard, missing):
    rng = np.random.default_rng(n + s + shard)
    g = scattered_relatives(rng, n, s, missing)
    sm = ck.submatrix(n, k, shard)
    osm = ko_sm(sm)
    bs = oracle_bitset(g, osm)
    with screen_ctx(level) as ctx, ctx.planes(sm, s) as pl:
        pl.pack(*triples_of(g))
        assert pl.king_variant() == 5
        all_pairs, _, _ = ko.king(bs, s, osm, -1.0, 1 << 22)
        kins = np.sort(all_pairs["kin"][np.isfinite(all_pairs["kin"])])
        # thresholds: the usual ones, and values sitting exactly on / next to kinships that occur (strict comparison)
        thrs = [0.0884, 0.0442, 0.2, 0.0, -0.05, 0.49, 0.5, float(kins[-1]), float(np.nextafter(kins[-1], np.float32(-1))),
                float(kins[len(kins) // 2]), float(kins[-min(len(kins), 40)])]
        for thr in thrs:
            want, count, _ = ko.king(bs, s, osm, thr, 1 << 22)
            got = pl.king(thr, 1 << 22)
            assert pl.last_count == count, (thr, level)
            assert_results_equal(got, want)


@pytest.mark.parametrize("level", [1, 3])
def test_screen_on_arbitrary_tile_ranges(level):
    """ck_king_tiles slices the band-ordered tile grid anywhere: a 160-wide screen tile may lose either half to the cut."""
    rng = np.random.default_rng(77)
    n, s, thr = 1500, 700, 0.06
    g = scattered_relatives(rng, n, s, 0.01)
    sm = ck.submatrix(n)
    want, _, _ = ko.king(oracle_bitset(g, ko_sm(sm)), s, ko_sm(sm), thr, 1 << 22)
    with screen_ctx(level) as ctx, ctx.planes(sm, s) as pl:
        pl.pack(*triples_of(g))
        total = pl.num_tiles()
        cuts = sorted({0, total, *[int(x) for x in rng.integers(1, total, 9)]})
        parts = [pl.king(thr, 1 << 22, tiles=(a, b)).copy() for a, b in zip(cuts[:-1], cuts[1:])]
        got = np.sort(np.concatenate(parts), order=["sample_i", "sample_j"])
        assert_results_equal(got, want)


def test_automatic_level_follows_the_call_rate():
    """Same records whichever screen the library picks from the cohort's call rate (0.5 % and 12 % missing here)."""
    rng = np.random.default_rng(3)
    n, s = 400, 3000
    for missing, thr in ((0.005, 0.0884), (0.12, 0.0884)):
        g = random_genotypes(rng, n, s, missing=missing)
        sm = ck.submatrix(n)
        want, count, _ = ko.king(oracle_bitset(g, ko_sm(sm)), s, ko_sm(sm), thr, 1 << 22)
        with screen_ctx(None) as ctx, ctx.planes(sm, s) as pl:
            pl.pack(*triples_of(g))
            got = pl.king(thr, 1 << 22)
            assert_results_equal(got, want)
            assert len(got) == count
