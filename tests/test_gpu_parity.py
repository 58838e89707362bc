"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and, where present, the reference's own
ComputeKingKernel (oracle/_ref).  Integer fields bit-exact; kin bit-exact (same fp32 operation order)."""
import json
import os

import numpy as np
import pytest

import cuking_b200 as ck
from cuking_b200 import capi
from oracle import king_oracle as ko
from oracle import ref_kernel
from tests.helpers import random_genotypes, triples_of, oracle_bitset, ko_sm, assert_results_equal, bits_equal_f32

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ctx():
    c = ck.Context(0)
    yield c
    c.close()


def device_planes(ctx, g, sm, order="site"):
    pl = ctx.planes(sm, g.shape[1])
    site, sample, alt = triples_of(g, order)
    pl.pack(site, sample, alt)
    return pl


# ---- pack / layout --------------------------------------------------------------------------------------------


@pytest.mark.parametrize("n,s,k,shard", [(9, 130, 1, 0), (70, 1000, 1, 0), (130, 517, 3, 1), (130, 517, 3, 3),
                                          (257, 64, 2, 1), (5, 1, 1, 0), (64, 32, 1, 0), (65, 33, 1, 0)])
def test_pack_matches_oracle_bitset(ctx, n, s, k, shard):
    rng = np.random.default_rng(n * 1000 + s)
    g = random_genotypes(rng, n, s, missing=0.05)
    sm = ck.submatrix(n, k, shard)
    want = oracle_bitset(g, ko_sm(sm))
    for order in ("site", "sample"):
        with device_planes(ctx, g, sm, order) as pl:
            assert np.array_equal(pl.export_bitset(), want)


def test_pack_duplicates_conflicts_and_filter(ctx):
    # AND semantics (cuking.cu:687-697) incl. the out-of-shard filter (:677) applied before validation (:698)
    rng = np.random.default_rng(7)
    n, s = 50, 300
    sm = ck.submatrix(n, 2, 1)  # rows [0,25) cols [25,50)
    m = 20_000
    row = rng.integers(0, s, m).astype(np.int64)
    col = rng.integers(0, n + 10, m).astype(np.int64)  # some samples beyond the cohort: skipped
    alt = rng.integers(0, 3, m).astype(np.int32)
    alt[col >= n] = 9                                   # invalid but filtered out
    want = ko.new_bitset(ko_sm(sm), s)
    assert ko.pack(want, s, ko_sm(sm), row, col, alt) == -1
    with ctx.planes(sm, s) as pl:
        pl.pack(row[: m // 2], col[: m // 2], alt[: m // 2])  # incremental packing accumulates
        pl.pack(row[m // 2:], col[m // 2:], alt[m // 2:])
        assert np.array_equal(pl.export_bitset(), want)
        pl.reset()
        assert np.all(pl.export_bitset() == np.uint64(0xFFFFFFFFFFFFFFFF))


def test_pack_device_pointers_and_unaligned(ctx):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(8)
    g = random_genotypes(rng, 40, 200)
    sm = ck.submatrix(40)
    site, sample, alt = triples_of(g)
    want = oracle_bitset(g, ko_sm(sm))
    dev = torch.device("cuda:0")
    for off in (0, 1):  # off=1: views that start 8 / 4 bytes into an allocation -> not 16-byte aligned -> scalar path
        r = torch.from_numpy(np.concatenate([np.zeros(off, np.int64), site])).to(dev)[off:]
        c = torch.from_numpy(np.concatenate([np.zeros(off, np.int64), sample])).to(dev)[off:]
        a = torch.from_numpy(np.concatenate([np.zeros(off, np.int32), alt])).to(dev)[off:]
        assert (r.data_ptr() % 16 == 0) == (off == 0)
        with ctx.planes(sm, 200) as pl:
            pl.pack(r, c, a)
            assert np.array_equal(pl.export_bitset(), want)


def test_pack_narrow_equals_pack(ctx):
    # ck_pack_triples_narrow: uint32 / uint32 / uint8 triples (9 bytes instead of 20) from pageable host memory, page-locked
    # host memory (read in place by the kernel) and device memory; same planes, same errors
    import torch

    rng = np.random.default_rng(31)
    g = random_genotypes(rng, 203, 1500)
    for k, shard in [(1, 0), (3, 1), (3, 3)]:
        sm = ck.submatrix(203, k, shard)
        want = oracle_bitset(g, ko_sm(sm))
        site, sample, alt = triples_of(g)
        r32, c32, a8 = site.astype(np.uint32), sample.astype(np.uint32), alt.astype(np.uint8)
        sources = {
            "pageable": (r32, c32, a8),
            "pinned": tuple(torch.from_numpy(x.view(np.int32 if x.dtype == np.uint32 else np.uint8)).pin_memory() for x in (r32, c32, a8)),
            "device": tuple(torch.from_numpy(x.view(np.int32 if x.dtype == np.uint32 else np.uint8)).cuda() for x in (r32, c32, a8)),
        }
        for name, (r, c, a) in sources.items():
            with ctx.planes(sm, 1500) as pl:
                pl.pack_narrow(r, c, a)
                assert np.array_equal(pl.export_bitset(), want), (name, k, shard)
    with ctx.planes(ck.submatrix(203), 1500) as pl:
        for bad in (3, 255):
            a_bad = a8.copy()
            a_bad[777] = bad
            with pytest.raises(ck.CukingError, match=rf"n_alt_alleles \({bad}\) encountered at triple 777") as e:
                pl.pack_narrow(r32, c32, a_bad)
            assert e.value.code == capi.CK_ERR_INVALID_GENOTYPE
        r_bad = r32.copy()
        r_bad[5] = 1500
        with pytest.raises(ck.CukingError) as e:
            pl.pack_narrow(r_bad, c32, a8)
        assert e.value.code == capi.CK_ERR_OUT_OF_RANGE


def test_pack_rejects_bad_genotype_and_site(ctx):
    sm = ck.submatrix(4)
    with ctx.planes(sm, 100) as pl:
        with pytest.raises(ck.CukingError, match=r"Invalid value for n_alt_alleles \(3\)") as e:
            pl.pack(np.array([1, 2, 3]), np.array([0, 1, 2]), np.array([0, 3, 1]))
        assert e.value.code == capi.CK_ERR_INVALID_GENOTYPE  # cuking.cu:698-701
        with pytest.raises(ck.CukingError) as e:
            pl.pack(np.array([100]), np.array([0]), np.array([0]))
        assert e.value.code == capi.CK_ERR_OUT_OF_RANGE


@pytest.mark.parametrize("n,s,k,shard", [(100, 777, 1, 0), (100, 777, 3, 2), (64, 2048, 1, 0)])
def test_import_export_round_trip(ctx, n, s, k, shard):
    rng = np.random.default_rng(n + s)
    g = random_genotypes(rng, n, s, missing=0.1)
    sm = ck.submatrix(n, k, shard)
    bs = oracle_bitset(g, ko_sm(sm))
    with ctx.planes(sm, s) as pl:
        pl.import_bitset(bs)
        assert np.array_equal(pl.export_bitset(), bs)


# ---- pairwise kernel ---------------------------------------------------------------------------------------------


def load_kats():
    with open(os.path.join(GOLDEN, "kat_vectors.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 5])
@pytest.mark.parametrize("kat", load_kats(), ids=lambda k: k["id"])
def test_known_answer_vectors_on_gpu(ctx, kat, variant):
    ctx.set_king_variant(variant)
    g = np.array([kat["genotypes_i"], kat["genotypes_j"]], dtype=np.int8)
    sm = ck.submatrix(2)
    with device_planes(ctx, g, sm) as pl:
        counts, kin = pl.counts([0], [1])
        for f in ("het_i", "het_j", "both_het", "opposing_hom", "concordant_hom", "shared_sites"):
            assert counts[0][f] == kat[f], f
        if kat["kin_f32_hex"] == "nan":
            assert np.isnan(kin[0])
        else:
            assert kin.view(np.uint32)[0] == int(kat["kin_f32_hex"], 16)
        res = pl.king(-1.0, 4)
        assert len(res) == (1 if kat["emitted_at_threshold_minus_1"] else 0)
        if len(res):
            assert (res[0]["ibs0"], res[0]["ibs1"], res[0]["ibs2"]) == (kat["ibs0"], kat["ibs1"], kat["ibs2"])
    ctx.set_king_variant(-1)


CASES = [
    # n, sites, split, shard, threshold
    (2, 1, 1, 0, -1.0), (3, 31, 1, 0, -1.0), (10, 32, 1, 0, -1.0), (17, 33, 1, 0, -1.0), (64, 63, 1, 0, -1.0),
    (65, 64, 1, 0, -1.0), (66, 65, 1, 0, 0.05), (129, 511, 1, 0, 0.05), (200, 513, 1, 0, 0.0884),
    (300, 1000, 2, 0, 0.05), (300, 1000, 2, 1, -1.0), (300, 1000, 2, 2, 0.05), (301, 999, 4, 5, -1.0),
    (301, 999, 4, 9, -1.0), (1000, 1600, 3, 1, 0.05), (130, 10_017, 1, 0, 0.0442),
]


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 5])
@pytest.mark.parametrize("n,s,k,shard,thr", CASES)
def test_king_matches_oracle(ctx, n, s, k, shard, thr, variant):
    ctx.set_king_variant(variant)
    rng = np.random.default_rng(n * 31 + s * 7 + shard)
    g = random_genotypes(rng, n, s, missing=0.04)
    if n > 6:
        g[3] = -1                      # an all-missing sample
        g[5][g[5] == 1] = 0            # a sample without hets: min_hets == 0 -> NaN / -inf, never emitted
    sm = ck.submatrix(n, k, shard)
    osm = ko_sm(sm)
    want, count, _ = ko.king(oracle_bitset(g, osm), s, osm, thr, 1 << 20)
    with device_planes(ctx, g, sm) as pl:
        got = pl.king(thr, 1 << 20)
        assert pl.last_count == count
        assert_results_equal(got, want)
        # raw counters for a sample of pairs (includes pairs below the threshold)
        ii = rng.integers(sm.i_begin, sm.i_end, 64)
        jj = rng.integers(sm.j_begin, sm.j_end, 64)
        keep = ii < jj
        if keep.any():
            ii, jj = ii[keep], jj[keep]
            counts, kin = pl.counts(ii, jj)
            L = ko.lib()
            import ctypes as C
            bs = oracle_bitset(g, osm)
            for q in range(len(ii)):
                c, kq = ko.pair_counts(bs, s, L.ko_sample_offset(C.byref(osm), int(ii[q])),
                                       L.ko_sample_offset(C.byref(osm), int(jj[q])))
                assert {f: int(counts[q][f]) for f in c} == c
                assert bits_equal_f32([kin[q]], [kq]) or (np.isnan(kin[q]) and np.isnan(kq))
    ctx.set_king_variant(-1)


def _long_vector_case(ctx, n, s, expect_variant):
    rng = np.random.default_rng(s)
    g = rng.integers(0, 3, size=(n, s), dtype=np.int8)   # uniform genotypes: cheap to draw at 8e6 sites per sample
    g[rng.random((n, s), dtype=np.float32) < 0.03] = -1
    g[0] = 1; g[1] = 1          # both_het = num_sites: the largest count an accumulator can hold
    g[2] = 2; g[3] = 0          # opposing homozygotes at every site: xx = -num_sites
    g[4] = 2                    # concordant with 2 everywhere
    sm = ck.submatrix(n)
    osm = ko_sm(sm)
    bs, _ = ko.pack_dense(g)
    want, count, _ = ko.king(bs, s, osm, -1.0, 1 << 16)
    ctx.set_king_variant(-1)
    with ctx.planes(sm, s) as pl:
        pl.import_bitset(bs)
        assert pl.king_variant() == expect_variant
        got = pl.king(-1.0, 1 << 16)
        assert pl.last_count == count
        assert_results_equal(got, want)
        counts, _ = pl.counts([0, 2, 2], [1, 3, 4])
        assert counts[0]["both_het"] == s and counts[1]["opposing_hom"] == s and counts[2]["concordant_hom"] == s


def test_fp4_kernel_is_exact_at_its_largest_site_count(ctx):
    # 2^23 sites is the largest count the mxf4 probe verified the fp32 accumulation for (kFp4MaxSites)
    _long_vector_case(ctx, 7, 1 << 23, expect_variant=5)


def test_longer_genotype_vectors_take_the_int8_kernel(ctx):
    _long_vector_case(ctx, 6, (1 << 23) + 37, expect_variant=2)


def test_king_unsorted_is_a_permutation(ctx):
    rng = np.random.default_rng(21)
    g = random_genotypes(rng, 150, 400)
    sm = ck.submatrix(150)
    with device_planes(ctx, g, sm) as pl:
        a = pl.king(0.0, 1 << 16, sort=True)
        b = pl.king(0.0, 1 << 16, sort=False)
        assert np.array_equal(np.sort(b, order=["sample_i", "sample_j"]), a)


def test_overflow_reports_resource_exhausted(ctx):
    rng = np.random.default_rng(22)
    g = random_genotypes(rng, 80, 300)
    sm = ck.submatrix(80)
    _, count, ovf = ko.king(oracle_bitset(g, ko_sm(sm)), 300, ko_sm(sm), -1.0, 100)
    assert ovf
    with device_planes(ctx, g, sm) as pl:
        with pytest.raises(ck.CukingError, match="--max_results") as e:
            pl.king(-1.0, 100)
        assert e.value.code == capi.CK_ERR_RESULT_OVERFLOW  # cuking.cu:747-751
        assert pl.last_count == count                       # the counter keeps counting (cuking.cu:299)
        assert len(pl.king(-1.0, count)) == count           # exactly enough room works


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_tile_slices_union_equals_full(ctx, variant):
    # the multi-GPU partition: contiguous slices of the tile grid, no exchange between slices
    rng = np.random.default_rng(23)
    ctx.set_king_variant(variant)
    for n, k, shard in [(333, 1, 0), (333, 2, 1), (700, 1, 0)]:
        g = random_genotypes(rng, n, 500)
        sm = ck.submatrix(n, k, shard)
        with device_planes(ctx, g, sm) as pl:
            full = pl.king(0.02, 1 << 18)
            tiles = pl.num_tiles()
            for parts in (2, 3, 8):
                cuts = [tiles * p // parts for p in range(parts + 1)]
                got = np.concatenate([pl.king(0.02, 1 << 18, tiles=(cuts[p], cuts[p + 1])) for p in range(parts)])
                got = np.sort(got, order=["sample_i", "sample_j"])
                assert_results_equal(got, full)
    ctx.set_king_variant(-1)


def test_host_bitset_seam_matches_oracle(ctx):
    rng = np.random.default_rng(24)
    n, s = 260, 900
    g = random_genotypes(rng, n, s)
    for k, shard in [(1, 0), (2, 1)]:
        osm = ko.submatrix(n, k, shard)
        bs = oracle_bitset(g, osm)
        want, _, _ = ko.king(bs, s, osm, 0.05, 1 << 16)
        got = ctx.king_host_bitset(n, k, shard, s, bs, 0.05, 1 << 16)
        assert_results_equal(got, want)


def test_host_bitset_pipelined_upload_matches_oracle(ctx):
    # >= 4096 samples on a diagonal shard: ck_king_host_bitset uploads the bit set last chunk first and launches each
    # chunk's bands while the next chunk is still in flight; the result must not depend on that schedule
    rng = np.random.default_rng(77)
    n, s = 4400, 300
    g = random_genotypes(rng, n, s)
    osm = ko.submatrix(n, 1, 0)
    bs = oracle_bitset(g, osm)
    want, count, _ = ko.king(bs, s, osm, 0.1, 1 << 20)
    assert count > 0
    for _ in range(2):  # second call: cached buffers and tile table
        got = ctx.king_host_bitset(n, 1, 0, s, bs, 0.1, 1 << 20)
        assert_results_equal(got, want)
    with pytest.raises(ck.CukingError):
        ctx.king_host_bitset(n, 1, 0, s, bs, 0.1, count - 1)  # overflow is still the reference's error
    # the multi-GPU form: disjoint parts whose union is the whole result (here all on one GPU)
    for parts in (2, 3):
        got = np.concatenate([ctx.king_host_bitset(n, 1, 0, s, bs, 0.1, 1 << 20, part=(p, parts)) for p in range(parts)])
        assert len(got) == count
        assert_results_equal(np.sort(got, order=["sample_i", "sample_j"]), want)
    small = ko.submatrix(300, 2, 1)  # off-diagonal and small: the plain path, contiguous tile slices
    g2 = random_genotypes(rng, 300, s)
    bs2 = oracle_bitset(g2, small)
    want2, _, _ = ko.king(bs2, s, small, 0.05, 1 << 16)
    got2 = np.concatenate([ctx.king_host_bitset(300, 2, 1, s, bs2, 0.05, 1 << 16, part=(p, 2)) for p in range(2)])
    assert_results_equal(np.sort(got2, order=["sample_i", "sample_j"]), want2)


def test_streaming_delivery_matches_oracle(ctx):
    # ck_king_stream_*: the bit set arrives in descending row ranges (host memory, device memory, several parts)
    import torch
    from cuking_b200.distributed import stream_chunks

    rng = np.random.default_rng(5)
    n, s = 2500, 400
    g = random_genotypes(rng, n, s)
    osm = ko.submatrix(n, 1, 0)
    bs = oracle_bitset(g, osm)
    want, count, _ = ko.king(bs, s, osm, 0.1, 1 << 20)
    w = ck.words_per_sample(s)
    gran = 1024
    chunks = stream_chunks(n, gran, target_chunks=3)
    assert chunks[0][1] == n and chunks[-1][0] == 0 and all(b % gran == 0 for b, _ in chunks)
    dev_bits = torch.from_numpy(bs.view(np.int64)).cuda()
    for source in ("host", "device"):
        for parts in (1, 2):
            got = []
            for p in range(parts):
                with ctx.planes(ck.submatrix(n), s) as pl:
                    pl.stream_begin(0.1, 1 << 20, part=(p, parts))
                    for b, e in chunks:
                        rows = bs[b * w: e * w] if source == "host" else dev_bits[b * w: e * w]
                        pl.stream_rows(rows, b, e)
                    got.append(pl.stream_end(1 << 20))
            got = np.concatenate(got)
            assert_results_equal(np.sort(got, order=["sample_i", "sample_j"]), want)
    ctx.set_king_variant(2)  # the int8 tensor kernel shares the band-ordered tiles, hence the seam
    with ctx.planes(ck.submatrix(n), s) as pl:
        pl.stream_begin(0.1, 1 << 20)
        for b, e in chunks:
            pl.stream_rows(dev_bits[b * w: e * w], b, e)
        assert_results_equal(pl.stream_end(1 << 20), want)
    ctx.set_king_variant(0)
    with ctx.planes(ck.submatrix(n), s) as pl:
        with pytest.raises(ck.CukingError):
            pl.stream_begin(0.1, 1 << 20)                     # LOP3+POPC kernels: other tile order
    ctx.set_king_variant(-1)
    with ctx.planes(ck.submatrix(n), s) as pl:  # protocol errors are reported, not executed
        pl.stream_begin(0.1, 1 << 20)
        with pytest.raises(ck.CukingError):
            pl.stream_rows(bs[: gran * w], 0, gran)          # not the last rows first
        pl.stream_rows(bs[chunks[0][0] * w: n * w], chunks[0][0], n)
        with pytest.raises(ck.CukingError):
            pl.stream_end(1 << 20)                            # rows [0, chunks[0][0]) never arrived
    with ctx.planes(ck.submatrix(n, 2, 1), s) as pl:
        with pytest.raises(ck.CukingError):
            pl.stream_begin(0.1, 1 << 20)                     # off-diagonal shard


def test_allgather_fed_seam_single_rank(ctx):
    # cuking_b200.distributed.king_host_bitset_allgather with a one-rank NCCL group: pinned host bit set -> per-chunk
    # upload on a side stream -> ck_king_stream_rows; the N > 1 form differs only by the all-gather between the two
    import torch
    import torch.distributed as dist
    from cuking_b200.distributed import king_host_bitset_allgather

    rng = np.random.default_rng(11)
    n, s = 3100, 350
    g = random_genotypes(rng, n, s)
    osm = ko.submatrix(n, 1, 0)
    bs = oracle_bitset(g, osm)
    want, count, _ = ko.king(bs, s, osm, 0.1, 1 << 20)
    host = torch.from_numpy(bs.view(np.int64)).pin_memory()
    created = not dist.is_initialized()
    if created:
        # a GPU box that cannot create a one-rank NCCL group is broken for every multi-GPU path: fail, do not skip
        # (a free port is chosen so that a stale listener cannot be the reason)
        import socket

        with socket.socket() as sock:
            sock.bind(("127.0.0.1", 0))
            port = sock.getsockname()[1]
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
    try:
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            with ck.Context(0, stream=stream.cuda_stream) as c2, c2.planes(ck.submatrix(n), s) as pl:
                got = king_host_bitset_allgather(pl, host, ck.words_per_sample(s), 0.1, 1 << 20)
                assert_results_equal(got, want)
            # a ctx on its own stream: the function binds it to the stream it orders the uploads against
            with ck.Context(0) as c3, c3.planes(ck.submatrix(n), s) as pl:
                got = king_host_bitset_allgather(pl, host, ck.words_per_sample(s), 0.1, 1 << 20)
                assert_results_equal(got, want)
                assert c3.stream is None  # restored
            # whole-cohort import with the upload shared between the ranks (here: one)
            from cuking_b200.distributed import import_bitset_allgather

            with ck.Context(0) as c4, c4.planes(ck.submatrix(n), s) as pl:
                import_bitset_allgather(pl, host, ck.words_per_sample(s))
                assert np.array_equal(pl.export_bitset(), bs)
                assert_results_equal(pl.king(0.1, 1 << 20), want)
    finally:
        if created:
            dist.destroy_process_group()


# ---- synthetic cohort ------------------------------------------------------------------------------------------


@pytest.mark.parametrize("n,s,k,shard", [(100, 700, 1, 0), (203, 333, 3, 1), (64, 10_000, 1, 0)])
def test_synth_planes_equal_packed_triples_equal_host_generator(ctx, n, s, k, shard):
    torch = pytest.importorskip("torch")
    sm = ck.submatrix(n, k, shard)
    g = ck.synth_genotypes_host(42, 0.02, 0, n, 0, s)
    want = oracle_bitset(g, ko_sm(sm))
    with ctx.planes(sm, s) as a:
        a.synthesize(42, 0.02)
        assert np.array_equal(a.export_bitset(), want)
    r, c, al, cnt = ctx.synth_triples_device(42, 0.02, 0, n, 0, s)
    assert cnt == int(np.sum(g >= 0))
    with ctx.planes(sm, s) as b:
        b.pack_device_ptrs(r, c, al, cnt)
        assert np.array_equal(b.export_bitset(), want)


def test_synth_triples_are_in_hail_order(ctx):
    torch = pytest.importorskip("torch")
    import ctypes as C
    n, s = 70, 90
    r, c, al, cnt = ctx.synth_triples_device(42, 0.1, 3, n, 5, s)
    g = ck.synth_genotypes_host(42, 0.1, 3, n, 5, s)
    site, sample, alt = triples_of(g)
    libcudart = C.CDLL("libcudart.so")
    def fetch(ptr, dtype):
        out = np.empty(cnt, dtype=dtype)
        assert libcudart.cudaMemcpy(C.c_void_p(out.ctypes.data), C.c_void_p(ptr), C.c_size_t(out.nbytes), 2) == 0
        return out
    assert np.array_equal(fetch(r, np.int64), site + 5)
    assert np.array_equal(fetch(c, np.int64), sample + 3)
    assert np.array_equal(fetch(al, np.int32), alt)


# ---- three-way: reference kernel itself -------------------------------------------------------------------------


@pytest.mark.skipif(not ref_kernel.available(), reason="oracle/_ref/libcuking_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("n,s,k,shard,thr", [(1000, 10_000, 1, 0, 0.05), (400, 3000, 2, 1, -1.0), (257, 1999, 3, 3, -1.0)])
def test_three_way_reference_kernel_oracle_product(ctx, n, s, k, shard, thr):
    g = ck.synth_genotypes_host(42, 0.02, 0, n, 0, s)
    sm = ck.submatrix(n, k, shard)
    osm = ko_sm(sm)
    bs = oracle_bitset(g, osm)
    cap = 1 << 20
    want, count, _ = ko.king(bs, s, osm, thr, cap)
    ref, ref_count, ref_ovf, _ = ref_kernel.king(bs, n, k, shard, ko.words_per_sample(s), thr, cap)
    assert ref_count == count and not ref_ovf
    assert_results_equal(ref, want)           # oracle == the reference's own kernel
    for variant in (0, 1, 2, 3):
        ctx.set_king_variant(variant)
        with ctx.planes(sm, s) as pl:
            pl.import_bitset(bs)
            assert_results_equal(pl.king(thr, cap), ref)  # product == the reference's own kernel
    ctx.set_king_variant(-1)


# ---- scale: dense output and the BASELINE single-GPU shape, checked through size-independent properties ---------------


def test_dense_output_every_pair_matches_oracle(ctx):
    # BASELINE.json configs[4] in miniature: threshold -1 emits every pair with a finite kinship
    n, s = 3000, 5000
    g = ck.synth_genotypes_host(42, 0.01, 0, n, 0, s)
    sm = ck.submatrix(n)
    osm = ko_sm(sm)
    want, count, ovf = ko.king(oracle_bitset(g, osm), s, osm, -1.0, n * (n - 1) // 2)
    assert not ovf and count == n * (n - 1) // 2
    with ctx.planes(sm, s) as pl:
        pl.synthesize(42, 0.01)
        got = pl.king(-1.0, count)
        assert_results_equal(got, want)
        with pytest.raises(ck.CukingError):
            pl.king(-1.0, count - 1)  # one slot short -> ResourceExhausted


def test_baseline_shape_properties(ctx):
    # BASELINE.json configs[1]: 100,000 samples x 100,000 sites, threshold 0.0884 — too large for the oracle as a
    # whole, so: structural properties of the full result + exact oracle values for a sample of retained pairs and
    # for one whole 256 x 256 block of the matrix.
    n, s, thr, seed, miss = 100_000, 100_000, 0.0884, 42, 0.01
    with ctx.planes(ck.submatrix(n), s) as pl:
        pl.synthesize(seed, miss)
        res = pl.king(thr, 10 << 20)
    key = res["sample_i"].astype(np.int64) * n + res["sample_j"]
    assert np.all(np.diff(key) > 0)                                  # sorted by (i, j), no duplicates
    assert np.all(res["sample_i"] < res["sample_j"]) and np.all(res["kin"] > np.float32(thr))
    assert np.all(res["sample_i"] // 8 == res["sample_j"] // 8)      # only planted pedigrees are related
    pairs = set(zip(res["sample_i"].tolist(), res["sample_j"].tolist()))
    first_degree = [(0, 2), (0, 3), (1, 2), (1, 3), (2, 3), (2, 5), (4, 5), (5, 7), (6, 7)]
    for b in (0, 1, 777, n // 8 - 1):
        for a, c in first_degree:
            assert (8 * b + a, 8 * b + c) in pairs
    assert 12 * (n // 8) <= len(res) <= 16 * (n // 8)

    def oracle_pairs(samples_i, samples_j):
        ids = sorted(set(samples_i) | set(samples_j))
        pos = {x: q for q, x in enumerate(ids)}
        g = np.stack([ck.synth_genotypes_host(seed, miss, x, x + 1, 0, s)[0] for x in ids])
        bs, _ = ko.pack_dense(g)
        return [ko.pair_counts(bs, s, pos[a], pos[b]) for a, b in zip(samples_i, samples_j)]

    rng = np.random.default_rng(0)
    pick = rng.choice(len(res), 48, replace=False)
    for q, (c, kin) in zip(pick, oracle_pairs(res["sample_i"][pick].tolist(), res["sample_j"][pick].tolist())):
        r = res[q]
        assert (int(r["ibs0"]), int(r["ibs2"])) == (c["opposing_hom"], c["concordant_hom"] + c["both_het"])
        assert int(r["ibs1"]) == c["shared_sites"] - int(r["ibs0"]) - int(r["ibs2"])
        assert bits_equal_f32([r["kin"]], [kin])

    # one whole off-diagonal 256 x 256 block, far from the origin: exact retained set vs the oracle
    i0, j0, w = 40_000, 40_192, 256  # overlaps the diagonal band so that related pairs exist in it
    ids = list(range(i0, i0 + w)) + list(range(j0, j0 + w))
    ids = sorted(set(ids))
    g = ck.synth_genotypes_host(seed, miss, ids[0], ids[-1] + 1, 0, s)
    osm = ko.submatrix(g.shape[0])
    want, _, _ = ko.king(oracle_bitset(g, osm), s, osm, thr, 1 << 20)
    sel = (res["sample_i"] >= ids[0]) & (res["sample_j"] <= ids[-1])
    got = res[sel].copy()
    got["sample_i"] -= ids[0]
    got["sample_j"] -= ids[0]
    assert_results_equal(got, want)


# ---- the other BASELINE.json configs at their real shard shapes -------------------------------------------------------
# One shard of each (what one GPU holds), synthetic cohort of SURVEY.md §8d generated on the device; the whole result
# is checked through structure, the records of a sample of retained pairs and ONE whole rectangle of the pair matrix
# against the oracle run on the same cohort regenerated on the host.


def _check_shard_against_oracle(res, n_sites, thr, seed, miss, block_rows, block_cols, sampled=32):
    assert np.all(res["sample_i"] < res["sample_j"]) and np.all(res["kin"] > np.float32(thr))
    key = res["sample_i"].astype(np.int64) << 32 | res["sample_j"].astype(np.int64)
    assert np.all(np.diff(key) > 0)  # sorted by (i, j), no duplicates

    def genotypes(a, b):
        return ck.synth_genotypes_host(seed, miss, a, b, 0, n_sites)

    # (1) records of a sample of retained pairs, recomputed by the oracle
    rng = np.random.default_rng(len(res))
    for q in rng.choice(len(res), min(sampled, len(res)), replace=False):
        r = res[q]
        g = np.concatenate([genotypes(int(r["sample_i"]), int(r["sample_i"]) + 1), genotypes(int(r["sample_j"]), int(r["sample_j"]) + 1)])
        bs, _ = ko.pack_dense(g)
        c, kin = ko.pair_counts(bs, n_sites, 0, 1)
        assert (int(r["ibs0"]), int(r["ibs2"])) == (c["opposing_hom"], c["concordant_hom"] + c["both_het"])
        assert int(r["ibs1"]) == c["shared_sites"] - int(r["ibs0"]) - int(r["ibs2"])
        assert bits_equal_f32([r["kin"]], [kin])
    # (2) one whole rectangle rows x cols (disjoint ranges, rows before cols): exact retained set and records
    (r0, r1), (c0, c1) = block_rows, block_cols
    assert r1 <= c0
    g = np.concatenate([genotypes(r0, r1), genotypes(c0, c1)])
    nr = r1 - r0
    osm = ko.Submatrix(0, nr, nr, nr + (c1 - c0))
    want, _, ovf = ko.king(oracle_bitset(g, osm), n_sites, osm, thr, 1 << 22)
    assert not ovf
    sel = (res["sample_i"] >= r0) & (res["sample_i"] < r1) & (res["sample_j"] >= c0) & (res["sample_j"] < c1)
    got = res[sel].copy()
    got["sample_i"] -= r0
    got["sample_j"] -= c0 - nr
    assert_results_equal(got, want)
    return len(want)


def test_cfg3_shapes_300k_samples_split_factor_4(ctx):
    # BASELINE.json configs[2]: 300,000 samples x 100,000 sites, split_factor 4 -> 75,000-sample blocks.
    n, s, seed, miss = 300_000, 100_000, 42, 0.01
    # diagonal shard 4 (block 1 x block 1), the config's threshold
    sm = ck.submatrix(n, 4, 4)
    assert (sm.i_begin, sm.i_end, sm.j_begin, sm.j_end) == (75_000, 150_000, 75_000, 150_000)
    with ctx.planes(sm, s) as pl:
        pl.synthesize(seed, miss)
        res = pl.king(0.05, 10 << 20)
    assert np.all(res["sample_i"] // 8 == res["sample_j"] // 8) and len(res) >= 12 * (75_000 // 8)
    assert _check_shard_against_oracle(res, s, 0.05, seed, miss, (100_000, 100_124), (100_124, 100_256)) > 0  # splits a pedigree
    # off-diagonal shard 1 (block 0 x block 1): no planted relatives there, so a threshold four standard deviations
    # above the unrelated mean keeps ~1e5 chance pairs whose set and records must still be exact
    sm = ck.submatrix(n, 4, 1)
    assert (sm.i_begin, sm.i_end, sm.j_begin, sm.j_end) == (0, 75_000, 75_000, 150_000)
    with ctx.planes(sm, s) as pl:
        pl.synthesize(seed, miss)
        res = pl.king(0.011, 10 << 20)
    assert 1_000 < len(res) < (10 << 20)
    assert res["sample_i"].max() < 75_000 <= res["sample_j"].min()
    _check_shard_against_oracle(res, s, 0.011, seed, miss, (31_000, 31_256), (140_000, 140_256))


def test_cfg4_shape_1m_samples_5pct_missing(ctx):
    # BASELINE.json configs[3]: 1,000,000 samples x 100,000 sites, 5 % missing, threshold 0.0442; with 8 x 8 blocks of
    # 125,000 samples a GPU holds one block pair at a time: the last diagonal shard (ragged against nothing, but at the
    # far end of the index range) and one off-diagonal shard.
    n, s, seed, miss, thr = 1_000_000, 100_000, 42, 0.05, 0.0442
    last = ck.num_shards(8) - 1
    sm = ck.submatrix(n, 8, last)
    assert (sm.i_begin, sm.i_end, sm.j_begin, sm.j_end) == (875_000, 1_000_000, 875_000, 1_000_000)
    with ctx.planes(sm, s) as pl:
        pl.synthesize(seed, miss)
        res = pl.king(thr, 10 << 20)
    assert np.all(res["sample_i"] // 8 == res["sample_j"] // 8) and len(res) >= 12 * (125_000 // 8)
    assert _check_shard_against_oracle(res, s, thr, seed, miss, (999_744, 999_868), (999_868, 1_000_000)) > 0  # splits a pedigree


def test_cfg5_shape_dense_output_1m_sites(ctx):
    # BASELINE.json configs[4] (50,000 samples x 1,000,000 sites, threshold -1: every pair is emitted), one diagonal
    # shard of split_factor 16: 3,125 samples -> 4.9e6 records, each the sum of 10^6 sites.
    n, s, seed, miss = 50_000, 1_000_000, 42, 0.01
    sm = ck.submatrix(n, 16, ck.num_shards(16) - 1)
    rows = sm.i_end - sm.i_begin
    assert rows == 3_125
    with ctx.planes(sm, s) as pl:
        pl.synthesize(seed, miss)
        assert pl.king_variant() == 5  # the default; dense output bypasses its screens
        res = pl.king(-1.0, rows * (rows - 1) // 2)
    assert len(res) == rows * (rows - 1) // 2  # the synthetic cohort has hets everywhere: every kinship is finite
    assert np.all(res["ibs0"].astype(np.int64) + res["ibs1"] + res["ibs2"] <= s)
    _check_shard_against_oracle(res, s, -1.0, seed, miss, (sm.i_begin + 1_000, sm.i_begin + 1_032), (sm.i_begin + 2_000, sm.i_begin + 2_032), sampled=8)
