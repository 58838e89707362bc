"""GPU parity tests of the round-2 entry points: shard views into cohort planes, snake-dealt parts, the dense
(sort-free) output path, chunked delivery to a sink, the kind::mxf4 self-test with its int8 fallback, adversarial
accumulation patterns through the kernel itself, and the multi-GPU plane exchange.  Checker: the CPU oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import cuking_b200 as ck
from cuking_b200 import capi
from oracle import king_oracle as ko
from tests.helpers import random_genotypes, triples_of, oracle_bitset, ko_sm, assert_results_equal, bits_equal_f32

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = ck.Context(0)
    yield c
    c.close()


def packed(ctx, g, sm):
    pl = ctx.planes(sm, g.shape[1])
    pl.pack(*triples_of(g))
    return pl


def oracle_shard(g, n, k, shard, thr, cap=1 << 22):
    osm = ko.submatrix(n, k, shard)
    want, count, ovf = ko.king(oracle_bitset(g, osm), g.shape[1], osm, thr, cap)
    assert not ovf
    return want


def device_count():
    n = C.c_int(0)
    capi.load().ck_device_count(C.byref(n))
    return n.value


# ---- views ---------------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("variant", [3, 2, 5])
@pytest.mark.parametrize("n,s", [(700, 500), (1300, 333)])
def test_every_shard_as_a_view_of_cohort_planes(ctx, variant, n, s):
    # one plane set for the whole cohort; each shard of --split_factor k is evaluated as a view (cuking.cu:129-152
    # builds one bit set per shard instead).  Shard edges (ceil(n / k)) are not multiples of the 64-sample plane blocks.
    rng = np.random.default_rng(n + s)
    g = random_genotypes(rng, n, s)
    ctx.set_king_variant(variant)
    try:
        with packed(ctx, g, ck.submatrix(n)) as pl:
            for k in (1, 2, 3, 5):
                for shard in range(ck.num_shards(k)):
                    view = ck.submatrix(n, k, shard)
                    for thr in (0.05, -1.0):
                        got = pl.king_view(view, thr, 1 << 20)
                        assert_results_equal(got, oracle_shard(g, n, k, shard, thr))
    finally:
        ctx.set_king_variant(-1)


def test_view_validation(ctx):
    rng = np.random.default_rng(1)
    g = random_genotypes(rng, 200, 100)
    with packed(ctx, g, ck.submatrix(200)) as pl:
        bad = capi.Submatrix(0, 100, 50, 150)  # rows and columns overlap without being identical
        with pytest.raises(ck.CukingError):
            pl.king_view(bad, 0.0)
        with pytest.raises(ck.CukingError):
            pl.king_view(capi.Submatrix(0, 100, 100, 201), 0.0)  # outside the planes
        with pytest.raises(ck.CukingError):
            pl.king_view(None, 0.0, part=(2, 2))
        ctx.set_king_variant(0)
        try:
            with pytest.raises(ck.CukingError, match="tensor-core"):
                pl.king_view(ck.submatrix(200, 2, 1), 0.0)
        finally:
            ctx.set_king_variant(-1)
    with packed(ctx, g, ck.submatrix(200, 2, 1)) as pl:  # planes of an off-diagonal shard hold two ranges: no views
        with pytest.raises(ck.CukingError, match="contiguous"):
            pl.king_view(ck.submatrix(200, 2, 1), 0.0)
        assert_results_equal(pl.king_view(None, 0.0, 1 << 16), oracle_shard(g, 200, 2, 1, 0.0))


def test_cta_pair_kernel_variant_matches_oracle(ctx):
    # variant 4: the mxf4 kernel on CTA pairs (tcgen05 cta_group::2, 256 x 64 tiles sharing the B operand); odd and even
    # numbers of 128-row tiles, ragged edges, off-diagonal shards, views, parts, dense output
    rng = np.random.default_rng(44)
    ctx.set_king_variant(4)
    try:
        for n, s in [(129, 333), (300, 1999), (1100, 700)]:
            g = random_genotypes(rng, n, s)
            for k, shard in [(1, 0), (2, 1), (3, 4)]:
                with packed(ctx, g, ck.submatrix(n, k, shard)) as pl:
                    assert pl.king_variant() == 4
                    for thr in (0.03, -1.0):
                        assert_results_equal(pl.king(thr, 1 << 21), oracle_shard(g, n, k, shard, thr))
        g = random_genotypes(rng, 2600, 260)
        with packed(ctx, g, ck.submatrix(2600)) as pl:
            for shard in range(3):
                parts = [pl.king_view(ck.submatrix(2600, 2, shard), 0.05, 1 << 21, part=(p, 3)).copy() for p in range(3)]
                got = np.sort(np.concatenate(parts), order=["sample_i", "sample_j"])
                assert_results_equal(got, oracle_shard(g, 2600, 2, shard, 0.05))
    finally:
        ctx.set_king_variant(-1)


# ---- parts ---------------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("variant", [3, 2, 1])
def test_parts_are_disjoint_and_their_union_is_the_shard(ctx, variant):
    rng = np.random.default_rng(99)
    n, s = 2600, 260  # three bands of 1024 rows
    g = random_genotypes(rng, n, s)
    ctx.set_king_variant(variant)
    try:
        with packed(ctx, g, ck.submatrix(n)) as pl:
            views = [None] if variant < 2 else [None, ck.submatrix(n, 2, 1), ck.submatrix(n, 2, 2)]
            for view in views:
                for thr in (0.1, -1.0):
                    full = pl.king_view(view, thr, 1 << 22)
                    for parts in (2, 3, 8):
                        got = [pl.king_view(view, thr, 1 << 22, part=(p, parts)).copy() for p in range(parts)]
                        for a in got:  # every part is sorted by itself
                            key = (a["sample_i"].astype(np.uint64) << np.uint64(32)) | a["sample_j"]
                            assert np.all(np.diff(key.astype(np.int64)) > 0)
                        merged = np.sort(np.concatenate(got), order=["sample_i", "sample_j"])
                        assert_results_equal(merged, full)
    finally:
        ctx.set_king_variant(-1)


# ---- dense output --------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("variant", [3, 2, 5])
@pytest.mark.parametrize("k,shard", [(1, 0), (2, 1)])
def test_dense_output_with_and_without_holes(ctx, variant, k, shard):
    # thr = -1 keeps (nearly) every pair: records are written straight to their sorted slot; thr = -0.02 leaves
    # below-threshold holes that are squeezed out; the result must be the oracle's either way
    rng = np.random.default_rng(7 + k)
    n, s = 1500, 300
    g = random_genotypes(rng, n, s)
    ctx.set_king_variant(variant)
    try:
        view = ck.submatrix(n, k, shard)
        with packed(ctx, g, ck.submatrix(n)) as pl:
            for thr in (-1.0, -0.02, -1e-6):
                want = oracle_shard(g, n, k, shard, thr)
                got = pl.king_view(view, thr, 1 << 21)
                assert_results_equal(got, want)
            total_pairs = n * (n - 1) // 2 if k == 1 else (n // 2) * (n - n // 2)
            assert 0 < len(oracle_shard(g, n, k, shard, -0.02)) < total_pairs  # the hole path really ran
            # a buffer without room for every pair takes the append + sort path and yields the same records
            want = oracle_shard(g, n, k, shard, -0.02)
            got = pl.king_view(view, -0.02, len(want) + 5)
            assert_results_equal(got, want)
    finally:
        ctx.set_king_variant(-1)


def test_dense_output_into_page_locked_memory_and_device_memory(ctx):
    import torch

    rng = np.random.default_rng(17)
    n, s = 1200, 200
    g = random_genotypes(rng, n, s)
    want = oracle_shard(g, n, 1, 0, -1.0)
    with packed(ctx, g, ck.submatrix(n)) as pl:
        pinned = torch.empty(len(want) * 24 + 240, dtype=torch.uint8).pin_memory()
        out = pinned.numpy().view(capi.RESULT_DTYPE)
        got = pl.king_view(None, -1.0, len(out), out=out)
        assert_results_equal(got, want)
        dev = torch.empty(len(want) * 24, dtype=torch.uint8, device="cuda")
        cnt = pl.king_view(None, -1.0, len(want), out=dev)  # device destination: append + sort on the device
        assert cnt == len(want)
        assert_results_equal(dev.cpu().numpy().view(capi.RESULT_DTYPE), want)
        cnt = pl.king_view(None, -1.0, len(want), out=dev, sort=False)
        assert cnt == len(want)
        assert_results_equal(np.sort(dev.cpu().numpy().view(capi.RESULT_DTYPE), order=["sample_i", "sample_j"]), want)


def test_host_bitset_seam_dense_output(ctx):
    # the pipelined host-buffer seam (rows delivered last chunk first) with dense output: regions complete in
    # descending order and are copied out as they do
    rng = np.random.default_rng(23)
    n, s = 4300, 150
    g = random_genotypes(rng, n, s, related_blocks=False)
    osm = ko.submatrix(n, 1, 0)
    bs = oracle_bitset(g, osm)
    for thr in (-1.0, -0.01):
        want, count, _ = ko.king(bs, s, osm, thr, 1 << 24)
        got = ctx.king_host_bitset(n, 1, 0, s, bs, thr, 1 << 24)
        assert_results_equal(got, want)
        parts = [ctx.king_host_bitset(n, 1, 0, s, bs, thr, 1 << 24, part=(p, 3)).copy() for p in range(3)]
        assert_results_equal(np.sort(np.concatenate(parts), order=["sample_i", "sample_j"]), want)


def test_host_bitset_seam_off_diagonal_shard_is_pipelined_too(ctx):
    # off-diagonal shard with >= 4 bands of rows: the column samples are uploaded first, the row samples in chunks behind
    # the kernels of the chunks before them; parts and dense output included
    rng = np.random.default_rng(29)
    n, s = 9100, 120
    g = random_genotypes(rng, n, s, related_blocks=False)
    g[4600:4700] = g[100:200]  # duplicates across the two blocks so that the threshold keeps some pairs
    osm = ko.submatrix(n, 2, 1)
    assert osm.i_end - osm.i_begin >= 4096
    bs = oracle_bitset(g, osm)
    for thr, cap in ((0.2, 1 << 20), (-1.0, 1 << 25)):
        want, count, _ = ko.king(bs, s, osm, thr, cap)
        assert count > 50
        got = ctx.king_host_bitset(n, 2, 1, s, bs, thr, cap)
        assert_results_equal(got, want)
    want, _, _ = ko.king(bs, s, osm, 0.2, 1 << 20)
    parts = [ctx.king_host_bitset(n, 2, 1, s, bs, 0.2, 1 << 20, part=(p, 3)).copy() for p in range(3)]
    assert_results_equal(np.sort(np.concatenate(parts), order=["sample_i", "sample_j"]), want)


# ---- chunked delivery ----------------------------------------------------------------------------------------------


@pytest.mark.parametrize("thr", [0.0, -1.0, -0.02])
def test_sink_receives_the_sorted_records_in_bounded_chunks(ctx, thr):
    rng = np.random.default_rng(31)
    n, s = 1100, 256
    g = random_genotypes(rng, n, s)
    with packed(ctx, g, ck.submatrix(n)) as pl:
        for view, k, shard in [(None, 1, 0), (ck.submatrix(n, 2, 1), 2, 1)]:
            want = oracle_shard(g, n, k, shard, thr)
            chunks = []
            delivered = pl.king_view_sink(view, thr, lambda r: chunks.append(r.copy()), 1 << 21, chunk_records=1024)
            assert delivered == len(want)
            assert all(0 < len(c) <= 1024 for c in chunks)
            assert len(chunks) >= len(want) // 1024
            assert_results_equal(np.concatenate(chunks) if chunks else np.empty(0, capi.RESULT_DTYPE), want)

        def boom(_records):
            raise RuntimeError("sink failure")

        with pytest.raises(ck.CukingError, match="sink"):
            pl.king_view_sink(None, -1.0, boom, 1 << 21, chunk_records=1024)
        assert_results_equal(pl.king_view(None, 0.05), oracle_shard(g, n, 1, 0, 0.05))  # the ctx is still usable


# ---- kind::mxf4 exactness guard ------------------------------------------------------------------------------------


def test_fp4_selftest_passes_on_this_gpu(ctx):
    exact, report = ctx.fp4_selftest()
    assert exact, report
    assert "0 differ" in report
    checked = int(report.split("self-test:")[1].split("of")[0])
    assert checked >= 1500, report  # most of the 2048 accumulators stay representable and are compared


def test_failed_selftest_routes_variant_3_to_the_int8_kernel():
    rng = np.random.default_rng(41)
    g = random_genotypes(rng, 300, 700)
    os.environ["CUKING_FP4_SELFTEST_FAIL"] = "1"
    try:
        with ck.Context(0) as c2, packed(c2, g, ck.submatrix(300)) as pl:
            assert pl.king_variant() == 2
            assert_results_equal(pl.king(0.05, 1 << 16), oracle_shard(g, 300, 1, 0, 0.05))
            exact, report = c2.fp4_selftest()
            assert not exact and "int8" in report
    finally:
        del os.environ["CUKING_FP4_SELFTEST_FAIL"]
    with ck.Context(0) as c3, packed(c3, g, ck.submatrix(300)) as pl:
        assert pl.king_variant() == 5


def test_adversarial_accumulation_patterns_through_the_kernel(ctx):
    # Hand-built genotype vectors at the mxf4 kernel's largest site count: counters that end one below a power of two
    # (odd low bits next to a large accumulator), with the odd site first or last, sign-alternating products, and a
    # single non-zero product per 64-site instruction.  Every counter must equal the oracle's popcounts.
    s = 1 << 23
    HET, REF, ALT, MISS = 1, 0, 2, -1
    even = np.arange(s) % 2 == 0
    first_of_64 = np.arange(s) % 64 == 0
    g = np.empty((12, s), dtype=np.int8)
    g[0] = HET
    g[1] = HET; g[1, 0] = REF                      # both_het(0,1) = 2^23 - 1: hh = 2^21 - 1/4, odd site first
    g[2] = HET; g[2, s - 1] = REF                  # the odd site last: the 1/4 meets an accumulator near 2^21
    g[3] = np.where(even, HET, ALT)                # het/hom alternate: hy and yh grow by halves
    g[4] = np.where(even, ALT, HET)
    g[5] = ALT; g[5, 0] = MISS                     # xx = +-(2^23 - 1)
    g[6] = REF; g[6, s - 1] = HET
    g[7] = np.where(even, ALT, REF)                # x products alternate in sign inside every instruction
    g[8] = np.where(first_of_64, HET, MISS)        # one non-zero operand per 64-site step
    g[9] = np.where(first_of_64, ALT, MISS)
    g[10] = np.where(np.arange(s) % 64 == 63, REF, MISS)
    g[11] = ALT
    bs, osm = ko.pack_dense(g)
    ii, jj = np.triu_indices(12, 1)
    ctx.set_king_variant(-1)
    with ctx.planes(ck.submatrix(12), s) as pl:
        pl.import_bitset(bs)
        assert pl.king_variant() == 5  # the default; the count dump runs the five-product kernel itself
        counts, kin = pl.counts(ii, jj)
        for q in range(len(ii)):
            c, kq = ko.pair_counts(bs, s, int(ii[q]), int(jj[q]))
            assert {f: int(counts[q][f]) for f in c} == c, (int(ii[q]), int(jj[q]))
            assert bits_equal_f32([kin[q]], [kq]) or (np.isnan(kin[q]) and np.isnan(kq))
        c01, _ = ko.pair_counts(bs, s, 0, 1)
        assert c01["both_het"] == s - 1
        c57, _ = ko.pair_counts(bs, s, 5, 7)
        assert c57["concordant_hom"] == s // 2 - 1 and c57["opposing_hom"] == s // 2
        got = pl.king(-1.0, 1 << 10)
        want, _, _ = ko.king(bs, s, osm, -1.0, 1 << 10)
        assert_results_equal(got, want)


def test_live_peak_measurements_are_sane(ctx):
    # the denominators bench.py quotes its rooflines against, measured on this GPU now
    burst = ctx.measure_fp4_peak()
    sustained = ctx.measure_fp4_peak_sustained(0.3)
    ints = ctx.measure_int_peaks()
    assert 5e15 < burst < 1.1e16, burst              # nominal dense fp4: 9e15 ops/s
    assert 4e15 < sustained <= burst * 1.02, (sustained, burst)
    assert 2e12 < ints["popc_lane_ops_per_s"] < 6e12 and ints["lop3_lane_ops_per_s"] > 3 * ints["popc_lane_ops_per_s"]


# ---- multi-GPU plane exchange --------------------------------------------------------------------------------------


@pytest.mark.skipif(device_count() < 2, reason="needs two GPUs with peer access")
@pytest.mark.parametrize("gpus", [2, 4, 8])
def test_and_reduce_of_dealt_triples_equals_packing_everything(gpus):
    if device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    rng = np.random.default_rng(gpus)
    n, s = 1500, 700
    g = random_genotypes(rng, n, s)
    site, sample, alt = triples_of(g)
    want_bits = oracle_bitset(g, ko.submatrix(n))
    want = oracle_shard(g, n, 1, 0, 0.05)
    ctxs = [ck.Context(d) for d in range(gpus)]
    try:
        planes = [c.planes(ck.submatrix(n), s) for c in ctxs]
        cuts = np.linspace(0, len(site), 5 * gpus + 1).astype(int)  # chunks of the triple stream, dealt round-robin
        for q in range(5 * gpus):
            lo, hi = cuts[q], cuts[q + 1]
            planes[q % gpus].pack(site[lo:hi], sample[lo:hi], alt[lo:hi])
        ck.and_reduce(planes)
        for d, pl in enumerate(planes):
            assert np.array_equal(pl.export_bitset(), want_bits), f"planes of GPU {d} differ after the exchange"
        parts = [pl.king_view(None, 0.05, 1 << 20, part=(d, gpus)).copy() for d, pl in enumerate(planes)]
        assert_results_equal(np.sort(np.concatenate(parts), order=["sample_i", "sample_j"]), want)
        for pl in planes:
            pl.close()
    finally:
        for c in ctxs:
            c.close()
