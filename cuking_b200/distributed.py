"""Multi-GPU execution: one process per GPU (torch.distributed).  Every rank holds the planes and evaluates its PART of a
shard - the bands of 1024 rows are dealt to the parts in snake order (ck_king_view) - or its share of the shards of a
--split_factor run (ck_plan_work); the sparse results are gathered on rank 0.

Pairs are independent (/root/reference/cuking.cu:197-201) and no count is ever combined across devices, so the
pairwise stage itself needs no collective; torch.distributed carries the result gather (a few MB), barriers and -
in king_host_bitset_allgather - the replication of the input planes over NVLink.  The reference scales the same way
but with one OS process per shard on separate VMs (README.md:94-102, cloud_batch_submit.py:45,73).
"""
from __future__ import annotations

import numpy as np

from .capi import RESULT_DTYPE, CukingError, CK_ERR_RESULT_OVERFLOW

BAND_ROWS = 1024  # rows per band of the tensor-core kernels' tile enumeration (csrc/band_tiles.cu) = the partition unit


def band_owner(band: int, num_parts: int) -> int:
    """Part that evaluates band `band` of a shard split into `num_parts` parts: snake order 0..P-1, P-1..0, ... (the
    work of a band falls linearly with its index, so every pair (g, 2P-1-g) carries the same work).  Mirror of
    band_owner() in csrc/king_api.cu."""
    g = band % (2 * num_parts)
    return g if g < num_parts else 2 * num_parts - 1 - g


def part_of_pair(i_local: int, num_parts: int) -> int:
    """Part that evaluates the pairs of row `i_local` (row index inside the shard)."""
    return band_owner(i_local // BAND_ROWS, num_parts)


_MIX = (np.uint64(0x9E3779B97F4A7C15), np.uint64(0xC2B2AE3D27D4EB4F), np.uint64(0x165667B19E3779F9))


def record_checksum(records: np.ndarray, chunk: int = 1 << 22) -> tuple[int, int, int]:
    """Order-independent checksum of KingResult records over all six fields: (count, sum, xor) of a 64-bit mix of each
    24-byte record.  Sums over disjoint sets of records add (mod 2^64) and xors xor, so the checksum of a result
    scattered over ranks is all_reduce(sum) / all_reduce(xor) of the per-rank values."""
    total, acc_sum, acc_xor = len(records), np.uint64(0), np.uint64(0)
    with np.errstate(over="ignore"):
        for lo in range(0, total, chunk):
            w = np.ascontiguousarray(records[lo: lo + chunk]).view(np.uint64).reshape(-1, 3)
            h = (w[:, 0] * _MIX[0]) ^ ((w[:, 1] + _MIX[2]) * _MIX[1]) ^ ((w[:, 2] ^ _MIX[0]) * _MIX[2])
            h ^= h >> np.uint64(29)
            acc_sum = acc_sum + h.sum(dtype=np.uint64)
            acc_xor = acc_xor ^ np.bitwise_xor.reduce(h) if len(h) else acc_xor
    return total, int(acc_sum), int(acc_xor)


def combine_checksums(parts: list[tuple[int, int, int]]) -> tuple[int, int, int]:
    n, s, x = 0, 0, 0
    for pn, ps, px in parts:
        n, s, x = n + pn, (s + ps) & 0xFFFFFFFFFFFFFFFF, x ^ px
    return n, s, x


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


def king_distributed(evaluate_part, max_results: int, group=None):
    """Runs `evaluate_part(part_index, num_parts) -> sorted results` (Planes.king_view(part=...)) for this rank's part,
    gathers and merges on rank 0.  Applies the reference's overflow rule to the TOTAL count (cuking.cu:747-751).
    Returns the merged array on rank 0, None elsewhere."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    local = evaluate_part(rank, world)
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
    # The uploads and all-gathers are ordered against `main` (torch's current stream) with events; the library queues
    # its transpose / code / pairwise kernels on the ctx's stream.  Those must be the same stream, or the kernels could
    # read a chunk before it has landed: bind the ctx to `main` for the duration of the call.
    ctx = planes.ctx
    prev_stream = ctx.stream
    ctx.set_stream(main.cuda_stream)
    try:
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
    finally:
        ctx.set_stream(prev_stream)
    del stage
    return res


def import_bitset_allgather(planes, host_bits, words_per_sample: int, group=None):
    """Loads a whole-cohort bit set (reference layout, pinned host memory, the same on every rank) into this rank's planes
    with the upload shared between the ranks: every rank copies 1/G of the samples through its own PCIe link, an NCCL
    all-gather over NVLink hands every GPU the whole bit set, ck_planes_import_bitset transposes it on the device.  Eight
    concurrent full uploads are host-limited (bench.py `plane_exchange`); this moves 1/8 of the bytes per link."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    hb = host_bits.view(torch.int64)
    n = hb.numel() // words_per_sample
    per = -(-n // world)  # samples per rank, the last ranks may be short or empty
    dev = torch.device("cuda", torch.cuda.current_device())
    full = torch.empty(per * world * words_per_sample, dtype=torch.int64, device=dev)
    b, e = min(rank * per, n), min((rank + 1) * per, n)
    mine = full[rank * per * words_per_sample: (rank + 1) * per * words_per_sample]
    ctx = planes.ctx
    main = torch.cuda.current_stream(dev)
    prev_stream = ctx.stream
    ctx.set_stream(main.cuda_stream)  # the import kernels must run behind the copies and the all-gather queued on `main`
    try:
        if e > b:
            mine[: (e - b) * words_per_sample].copy_(hb[b * words_per_sample: e * words_per_sample], non_blocking=True)
        if world > 1:
            dist.all_gather_into_tensor(full, mine, group=group)
        planes.import_bitset(full[: n * words_per_sample])
    finally:
        ctx.set_stream(prev_stream)
    del full
