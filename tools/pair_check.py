#!/usr/bin/env python3
"""Checks a tensor-core kernel variant (argv[1]; default 4, the CTA-pair mxf4 kernel; 5 = screen + mxf4) against variant 3 and the oracle on shapes that stress its tile pairing:
odd / even numbers of row tiles, ragged edges, off-diagonal shards, views, parts, dense output."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import cuking_b200 as ck  # noqa: E402
from oracle import king_oracle as ko  # noqa: E402
from tests.helpers import random_genotypes, triples_of, oracle_bitset, assert_results_equal  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ok = True
with ck.Context(0) as ctx:
    rng = np.random.default_rng(5)
    for n, s in [(100, 700), (129, 333), (257, 1999), (300, 3000), (640, 512), (1100, 900), (2600, 300)]:
        g = random_genotypes(rng, n, s)
        for k, shard in [(1, 0), (2, 1), (3, 4)]:
            sm = ck.submatrix(n, k, shard)
            osm = ko.submatrix(n, k, shard)
            want, count, _ = ko.king(oracle_bitset(g, osm), s, osm, 0.03, 1 << 22)
            res = {}
            for v in (3, V):
                ctx.set_king_variant(v)
                with ctx.planes(sm, s) as pl:
                    pl.pack(*triples_of(g))
                    assert pl.king_variant() == v
                    res[v] = pl.king(0.03, 1 << 22).copy()
                    dense = pl.king(-1.0, 1 << 22).copy()
                    if v == 3:
                        dense3 = dense
                    else:
                        try:
                            assert_results_equal(dense, dense3)
                        except AssertionError as exc:
                            ok = False
                            print("DENSE MISMATCH", n, s, k, shard, exc)
            try:
                assert_results_equal(res[V], want)
                print("ok", n, s, k, shard, count)
            except AssertionError as exc:
                ok = False
                print("MISMATCH", n, s, k, shard, len(res[V]), len(res[3]), count, exc)
    # views and parts of a cohort
    g = random_genotypes(rng, 2600, 260)
    ctx.set_king_variant(V)
    with ctx.planes(ck.submatrix(2600), 260) as pl:
        pl.pack(*triples_of(g))
        for k in (2, 3):
            for shard in range(ck.num_shards(k)):
                osm = ko.submatrix(2600, k, shard)
                want, _, _ = ko.king(oracle_bitset(g, osm), 260, osm, 0.05, 1 << 22)
                parts = [pl.king_view(ck.submatrix(2600, k, shard), 0.05, 1 << 22, part=(p, 3)).copy() for p in range(3)]
                got = np.sort(np.concatenate(parts), order=["sample_i", "sample_j"])
                try:
                    assert_results_equal(got, want)
                except AssertionError as exc:
                    ok = False
                    print("VIEW MISMATCH", k, shard, exc)
    print("views/parts checked")
print("PAIR CHECK", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
