"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: tile partition, result gather, merge, overflow rule.
The per-slice evaluation is stood in for by the oracle restricted to the slice's tiles; on a GPU box the same
functions drive Planes.king(tiles=...) (tests/test_gpu_parity.py::test_tile_slices_union_equals_full)."""
import os
import socket

import numpy as np
import pytest

from cuking_b200 import distributed as ckd
from cuking_b200.capi import CukingError
from oracle import king_oracle as ko
from tests.helpers import random_genotypes, oracle_bitset


@pytest.mark.parametrize("n,tri", [(1, True), (5, True), (37, True), (1563, True)])
def test_tile_coords_enumerates_upper_triangle(n, tri):
    total = n * (n + 1) // 2
    probe = range(total) if total < 2000 else list(range(0, total, 997)) + [total - 1]
    seen = set()
    for t in probe:
        bi, bj = ckd.tile_coords(t, n, n, True)
        assert 0 <= bi <= bj < n
        assert ckd.tile_of_pair(bi * 64, bj * 64, n, n, True) == t
        seen.add((bi, bj))
    assert len(seen) == len(list(probe))
    assert ckd.tile_coords(0, n, n, True) == (0, 0) and ckd.tile_coords(total - 1, n, n, True) == (n - 1, n - 1)


def test_tile_slices_partition_the_grid():
    for tiles in (0, 1, 7, 1222266):
        for world in (1, 2, 3, 8):
            cuts = [ckd.tile_slice(tiles, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == tiles
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1


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
        rows, cols = sm.i_end - sm.i_begin, sm.j_end - sm.j_begin
        tri = sm.i_begin == sm.j_begin
        rb, cb = -(-rows // 64), -(-cols // 64)
        tile = np.array([ckd.tile_of_pair(int(r["sample_i"]) - sm.i_begin, int(r["sample_j"]) - sm.j_begin, rb, cb, tri)
                         for r in full], dtype=np.int64)

        def evaluate_slice(b, e):  # what Planes.king(tiles=(b, e)) returns on a GPU
            return full[(tile >= b) & (tile < e)]

        try:
            merged = ckd.king_distributed(evaluate_slice, ckd.num_tiles(rows, cols, tri), cap)
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

    mp.spawn(_worker, args=(2, _free_port(), 300, 400, k, shard, 0.02, 1 << 20, str(tmp_path)), nprocs=2, join=True)
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
