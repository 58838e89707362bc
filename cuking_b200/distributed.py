"""Multi-GPU execution of one shard: one process per GPU (torch.distributed), every rank holds the shard's planes and
evaluates a contiguous slice of its 64x64 tile grid; the sparse results are gathered on rank 0.

There is NO data-path collective: pairs are independent (/root/reference/cuking.cu:197-201) and no count is ever
combined across devices.  torch.distributed carries only the result gather (a few MB) and barriers.  The reference
scales the same way but with one OS process per shard on separate VMs (README.md:94-102, cloud_batch_submit.py:45,73).
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
