"""Multi-GPU execution of one shard: one process per GPU (torch.distributed), every rank holds the shard's planes and
evaluates a contiguous slice of its 64x64 tile grid; the sparse results are gathered on rank 0.

Pairs are independent (/root/reference/cuking.cu:197-201) and no count is ever combined across devices, so the
pairwise stage itself needs no collective; torch.distributed carries the result gather (a few MB), barriers and -
in king_host_bitset_allgather - the replication of the input planes over NVLink.  The reference scales the same way
but with one OS process per shard on separate VMs (README.md:94-102, cloud_batch_submit.py:45,73).
"""
from __future__ import annotations

import math

import numpy as np

from .capi import RESULT_DTYPE, CukingError, CK_ERR_RESULT_OVERFLOW

TILE = 64  # samples per tile edge (cuking_b200/csrc/layout.cuh: kTileSamples)


def tile_slice(num_tiles: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced slice [begin, end) of the linear tile index space for `rank` of `world`."""
    return num_tiles * rank // world, num_tiles * (rank + 1) // world


def num_tiles(num_rows: int, num_cols: int, triangular: bool) -> int:
    rb, cb = -(-num_rows // TILE), -(-num_cols // TILE)
    return rb * (rb + 1) // 2 if triangular else rb * cb


def tile_coords(t: int, num_row_blocks: int, num_col_blocks: int, triangular: bool) -> tuple[int, int]:
    """Linear tile index -> (row block, column block); mirror of tile_coords() in csrc/king_kernel.cu.  Triangular
    grids (rows == columns) enumerate bj >= bi row-major."""
    if not triangular:
        return t // num_col_blocks, t % num_col_blocks
    n = num_row_blocks
    b = int((2 * n + 1 - math.isqrt((2 * n + 1) ** 2 - 8 * t)) // 2)
    off = lambda x: x * n - x * (x - 1) // 2
    while b > 0 and off(b) > t:
        b -= 1
    while b + 1 < n and off(b + 1) <= t:
        b += 1
    return b, b + (t - off(b))


def tile_of_pair(i_local: int, j_local: int, num_row_blocks: int, num_col_blocks: int, triangular: bool) -> int:
    """Linear tile index that evaluates the pair at (row offset, column offset) of the sub-matrix."""
    bi, bj = i_local // TILE, j_local // TILE
    if not triangular:
        return bi * num_col_blocks + bj
    return bi * num_row_blocks - bi * (bi - 1) // 2 + (bj - bi)


def merge_sorted(parts: list[np.ndarray]) -> np.ndarray:
    """Merges per-rank results (each sorted by (sample_i, sample_j)) into the global order of cuking.cu:761-765."""
    parts = [p for p in parts if len(p)]
    if not parts:
        return np.empty(0, dtype=RESULT_DTYPE)
    allp = np.concatenate(parts)
    key = (allp["sample_i"].astype(np.uint64) << np.uint64(32)) | allp["sample_j"].astype(np.uint64)
    return allp[np.argsort(key, kind="stable")]


def gather_results(local: np.ndarray, dst: int = 0, group=None):
    """Gathers variable-length result arrays on rank `dst` (None elsewhere).  Works with gloo (CPU tensors) and nccl
    (CUDA tensors); sizes travel first, payloads are padded to the largest."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    cuda = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes[rank] = len(local)
    dist.all_reduce(sizes, group=group)
    max_n = int(sizes.max().item())
    payload = torch.zeros(max(max_n, 1) * RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    if len(local):
        raw = torch.from_numpy(np.ascontiguousarray(local).view(np.uint8).copy())
        payload[: raw.numel()] = raw.to(dev)
    bucket = [torch.empty_like(payload) for _ in range(world)] if rank == dst else None
    dist.gather(payload, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    return [b.cpu().numpy()[: int(sizes[r].item()) * RESULT_DTYPE.itemsize].view(RESULT_DTYPE).copy()
            for r, b in enumerate(bucket)]


def king_distributed(evaluate_slice, total_tiles: int, max_results: int, group=None):
    """Runs `evaluate_slice(tile_begin, tile_end) -> sorted results` on this rank's slice, gathers and merges on rank 0.
    Applies the reference's overflow rule to the TOTAL count (cuking.cu:747-751).  Returns the merged array on rank 0,
    None elsewhere."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    begin, end = tile_slice(total_tiles, rank, world)
    local = evaluate_slice(begin, end)
    parts = gather_results(local, 0, group)
    if rank != 0:
        return None
    merged = merge_sorted(parts)
    if len(merged) > max_results:
        raise CukingError(CK_ERR_RESULT_OVERFLOW, "Could not store all results: try increasing the --max_results parameter.")
    return merged


def stream_chunks(num_samples: int, granularity: int, target_chunks: int = 24) -> list[tuple[int, int]]:
    """Descending sample ranges [(begin, end), ...] that tile [0, num_samples) at multiples of `granularity` - the
    delivery order of ck_king_stream_rows (the kernel starts on the bottom rows, which need no other sample)."""
    bands = -(-num_samples // granularity)
    per = max(1, -(-bands // target_chunks))
    out, hi = [], bands
    while hi > 0:
        lo = max(0, hi - per)
        out.append((lo * granularity, min(hi * granularity, num_samples)))
        hi = lo
    return out


def chunk_piece(begin: int, end: int, rank: int, world: int) -> tuple[int, int, int]:
    """Rows of chunk [begin, end) that `rank` uploads before the all-gather: (rows per piece, my_begin, my_end).  Pieces
    are equal-sized (the all-gather needs that), in rank order, so the gathered buffer holds the chunk's rows in order;
    the last pieces may be short or empty."""
    per = -(-(end - begin) // world)
    return per, min(begin + rank * per, end), min(begin + (rank + 1) * per, end)


def king_host_bitset_allgather(planes, host_bits, words_per_sample: int, kin_threshold: float, max_results: int,
                               out: np.ndarray | None = None, group=None, side_stream=None):
    """The host-buffer seam on the G GPUs of one box with the planes replicated over NVLink instead of G times over
    PCIe: every rank holds the bit set in pinned host memory (e.g. one shared mapping), uploads only ITS 1/G of every
    chunk through its own PCIe link, and an NCCL all-gather hands every GPU the whole chunk; ck_king_stream_rows then
    launches the rank's bands among those rows.  Chunks travel last rows first on a side stream, so upload and
    all-gather of chunk c+1 overlap the pairwise kernel of chunk c.  Measured on 8 x B200 (bench.py plane_exchange):
    eight concurrent full uploads are host-limited at 23 GB/s per GPU, while 1/8 per link + NVLink moves the 7 GB in a
    few tens of ms - north_star's "NCCL broadcast only if it beats per-GPU host loading".
    `planes`: this rank's Planes of the (diagonal) shard; `host_bits`: pinned CPU uint64/int64 tensor of the reference
    layout, [num_samples * words_per_sample].  Returns this rank's sorted records (parts are disjoint)."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    hb = host_bits.view(torch.int64)
    n = hb.numel() // words_per_sample
    dev = torch.device("cuda", torch.cuda.current_device())
    main = torch.cuda.current_stream(dev)
    side = side_stream if side_stream is not None else torch.cuda.Stream(dev)
    gran = int(planes._lib.ck_king_stream_granularity())
    chunks = stream_chunks(n, gran)
    # chunk rows are padded to a multiple of `world` so that the all-gather pieces are equal
    max_rows = max(e - b for b, e in chunks)
    piece_rows = -(-max_rows // world)
    stage = [torch.empty(piece_rows * world * words_per_sample, dtype=torch.int64, device=dev) for _ in range(len(chunks))]
    side.wait_stream(main)
    planes.stream_begin(kin_threshold, max_results, part=(rank, world))
    for c, (b, e) in enumerate(chunks):
        pr, my_b, my_e = chunk_piece(b, e, rank, world)
        buf = stage[c][: pr * world * words_per_sample]
        with torch.cuda.stream(side):
            mine = buf[rank * pr * words_per_sample: (rank + 1) * pr * words_per_sample]
            if my_e > my_b:
                mine[: (my_e - my_b) * words_per_sample].copy_(hb[my_b * words_per_sample: my_e * words_per_sample], non_blocking=True)
            if world > 1:
                dist.all_gather_into_tensor(buf, mine, group=group)
            ready = torch.cuda.Event()
            ready.record(side)
        main.wait_event(ready)
        planes.stream_rows(buf.data_ptr(), b, e)
    res = planes.stream_end(max_results, out=out)
    del stage
    return res
