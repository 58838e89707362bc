"""Repeats small pack + pairwise runs 150 times per shape and compares every one with the oracle: a race in the kernels' barrier
protocol would show up as an occasional mismatch.  usage: tools/flaky_check.py [kernel variant]"""
import sys; sys.path.insert(0, "/root/repo")
import numpy as np, ctypes as C
import cuking_b200 as ck
from oracle import king_oracle as ko
from tests.helpers import random_genotypes, triples_of, oracle_bitset, ko_sm
ctx = ck.Context(0, king_variant=int(sys.argv[1]) if len(sys.argv) > 1 else 3)  # 3 = mxf4 kernel (default), 2 = int8
bad_king = bad_counts = 0
for n, s in [(3, 31), (17, 33), (129, 511), (300, 1000)]:
    rng = np.random.default_rng(n * 31 + s * 7)
    g = random_genotypes(rng, n, s, missing=0.04)
    sm = ck.submatrix(n); osm = ko_sm(sm)
    bs = oracle_bitset(g, osm)
    want, count, _ = ko.king(bs, s, osm, -1.0, 1 << 20)
    ii, jj = np.triu_indices(n, 1)
    sel = rng.choice(len(ii), min(64, len(ii)), replace=False)
    ii, jj = ii[sel], jj[sel]
    ref = [ko.pair_counts(bs, s, int(a), int(b))[0] for a, b in zip(ii, jj)]
    site, sample, alt = triples_of(g)
    for it in range(150):
        with ctx.planes(sm, s) as pl:
            pl.pack(site, sample, alt)
            got = pl.king(-1.0, 1 << 20)
            if len(got) != len(want) or not np.array_equal(got["ibs0"], want["ibs0"]) or not np.array_equal(got["ibs2"], want["ibs2"]):
                bad_king += 1
            counts, kin = pl.counts(ii, jj)
            for q in range(len(ii)):
                if {f: int(counts[q][f]) for f in ref[q]} != ref[q]:
                    bad_counts += 1
                    break
    print(n, s, "bad king", bad_king, "bad counts", bad_counts, flush=True)
