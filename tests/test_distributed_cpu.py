"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: band partition, shard scheduling, result gather, merge,
overflow rule, record checksums.  The per-part evaluation is stood in for by the oracle restricted to the part's
bands; on a GPU box the same functions drive Planes.king_view(part=...)
(tests/test_gpu_views.py::test_parts_are_disjoint_and_their_union_is_the_shard)."""
import os
import socket

import numpy as np
import pytest

from cuking_b200 import distributed as ckd
from cuking_b200.capi import CukingError
from oracle import king_oracle as ko
from tests.helpers import random_genotypes, oracle_bitset


def test_band_owner_deals_bands_in_snake_order_and_balances_triangular_work():
    for parts in (1, 2, 3, 8):
        owners = [ckd.band_owner(b, parts) for b in range(4 * parts)]
        assert owners[: 2 * parts] == list(range(parts)) + list(range(parts - 1, -1, -1))
        assert owners[2 * parts:] == owners[: 2 * parts]
    # triangular shard: band b holds ~ (num_bands - b) units of work; the parts differ by less than one band's work
    for bands, parts in [(977, 8), (49, 8), (98, 2), (293, 4)]:
        load = [0] * parts
        for b in range(bands):
            load[ckd.band_owner(b, parts)] += bands - b
        assert max(load) - min(load) <= bands, (bands, parts, load)
        assert sum(load) / parts / max(load) > 0.97 or bands < 100


def test_record_checksum_is_order_independent_and_sensitive():
    rng = np.random.default_rng(3)
    rec = np.zeros(5000, dtype=ckd.RESULT_DTYPE)
    for f in ("sample_i", "sample_j", "ibs0", "ibs1", "ibs2"):
        rec[f] = rng.integers(0, 1 << 20, len(rec))
    rec["kin"] = rng.random(len(rec), dtype=np.float32)
    whole = ckd.record_checksum(rec, chunk=777)
    assert whole == ckd.record_checksum(rec[rng.permutation(len(rec))])
    cut = [0, 100, 100, 3210, 5000]
    assert ckd.combine_checksums([ckd.record_checksum(rec[a:b]) for a, b in zip(cut, cut[1:])]) == whole
    for f in ckd.RESULT_DTYPE.names:  # one changed field of one record changes the checksum
        other = rec.copy()
        other[f][1234] = other[f][1234] + 1
        assert ckd.record_checksum(other) != whole
    swapped = rec.copy()
    swapped["ibs0"][7], swapped["ibs1"][7] = rec["ibs1"][7], rec["ibs0"][7]
    assert ckd.record_checksum(swapped) != whole or rec["ibs0"][7] == rec["ibs1"][7]
    assert ckd.record_checksum(rec[:0]) == (0, 0, 0)


def test_work_plan_covers_every_shard_and_balances_the_gpus():
    import cuking_b200 as ck

    for n, k, gpus in [(300_000, 4, 8), (1_000_000, 1, 8), (100_000, 2, 8), (100_000, 1, 1), (50_000, 1, 8), (5_000, 5, 3)]:
        items = ck.plan_work(n, k, gpus)
        seen = {}
        load = [0.0] * gpus
        for it in items:
            assert it.gpu < gpus and it.part_index < it.num_parts
            seen.setdefault(it.shard_index, set()).add(it.part_index)
            load[it.gpu] += it.pairs / it.num_parts
        assert sorted(seen) == list(range(ck.num_shards(k)))
        for it in items:
            assert seen[it.shard_index] == set(range(it.num_parts))  # every part of every shard exactly once
        assert [it.gpu for it in items] == sorted(it.gpu for it in items)  # grouped by GPU, in execution order
        if n >= 50_000:
            assert sum(load) / gpus / max(load) > 0.95, (n, k, gpus, load)
    # BASELINE configs[2]: 300k samples, split_factor 4 -> 10 shards on 8 GPUs: six off-diagonal shards alone, the four
    # (half-cost) diagonal ones in pairs
    items = ck.plan_work(300_000, 4, 8)
    assert len(items) == 10 and all(it.num_parts == 1 for it in items)
    per_gpu = {}
    for it in items:
        per_gpu.setdefault(it.gpu, []).append(it.shard_index)
    assert sorted(len(v) for v in per_gpu.values()) == [1] * 6 + [2] * 2
    # a sub-range of the shards (one task per shard, cloud_batch_submit.py:45) and validation
    assert [it.shard_index for it in ck.plan_work(1000, 3, 1, first_shard=4, num_run=1)] == [4]
    with pytest.raises(ck.CukingError):
        ck.plan_work(1000, 3, 1, first_shard=5, num_run=2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, sites, k, shard, thr, cap, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = random_genotypes(np.random.default_rng(99), n, sites)
        sm = ko.submatrix(n, k, shard)
        full, _, _ = ko.king(oracle_bitset(g, sm), sites, sm, thr, 1 << 20)

        def evaluate_part(part, parts):  # what Planes.king_view(part=(part, parts)) returns on a GPU
            band = (full["sample_i"].astype(np.int64) - sm.i_begin) // ckd.BAND_ROWS
            owner = np.array([ckd.band_owner(int(b), parts) for b in band], dtype=np.int64)
            return full[owner == part]

        try:
            merged = ckd.king_distributed(evaluate_part, cap)
            if rank == 0:
                np.save(os.path.join(out_dir, "merged.npy"), merged)
                np.save(os.path.join(out_dir, "full.npy"), full)
        except CukingError as exc:
            if rank == 0:
                with open(os.path.join(out_dir, "error.txt"), "w") as f:
                    f.write(str(exc))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("k,shard", [(1, 0), (2, 1)])
def test_two_rank_gloo_union_equals_single(tmp_path, k, shard):
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(2, _free_port(), 2300, 200, k, shard, 0.02, 1 << 20, str(tmp_path)), nprocs=2, join=True)
    merged, full = np.load(tmp_path / "merged.npy"), np.load(tmp_path / "full.npy")
    assert len(full) > 10
    assert np.array_equal(merged, full)


def test_two_rank_overflow_is_global(tmp_path):
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(2, _free_port(), 200, 300, 1, 0, -1.0, 50, str(tmp_path)), nprocs=2, join=True)
    assert "max_results" in (tmp_path / "error.txt").read_text()


def test_stream_chunks_tile_the_shard_in_descending_order():
    # delivery order of ck_king_stream_rows: last rows first, boundaries at the stream granularity, nothing missing
    from cuking_b200.distributed import stream_chunks

    for n, gran, target in [(1, 1024, 24), (1024, 1024, 24), (1025, 1024, 24), (100_000, 1024, 24), (282_816, 1024, 24),
                            (5_000, 1024, 3), (2_500, 1024, 1)]:
        chunks = stream_chunks(n, gran, target)
        assert chunks[0][1] == n and chunks[-1][0] == 0
        assert all(b % gran == 0 and b < e for b, e in chunks)
        assert all(chunks[q][0] == chunks[q + 1][1] for q in range(len(chunks) - 1))  # contiguous, descending
        assert len(chunks) <= max(target, 1) + 1


def test_allgather_pieces_cover_every_chunk_in_rank_order():
    from cuking_b200.distributed import chunk_piece, stream_chunks

    for n, world in [(100_000, 2), (282_816, 8), (5_000, 3), (1_030, 4)]:
        for b, e in stream_chunks(n, 1024):
            pieces = [chunk_piece(b, e, r, world) for r in range(world)]
            per = pieces[0][0]
            assert all(p[0] == per for p in pieces) and per * world >= e - b
            covered = b
            for r, (_, pb, pe) in enumerate(pieces):
                assert pb == min(b + r * per, e) and pb <= pe <= e  # piece r sits at offset r * per of the gathered buffer
                assert pb == covered or pb == e
                covered = max(covered, pe)
            assert covered == e
