"""Object wrappers over the C ABI: Context (one GPU + stream) and Planes (packed genotype bit planes of one shard).

Mirrors the host side of the reference's Run() around its kernel launch (/root/reference/cuking.cu:505-765):
plan the shard, allocate the bit set, pack triples into it, run the pairwise kernel, check overflow, sort.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import RESULT_DTYPE, COUNTS_DTYPE, Submatrix, SynthParams, Timings, check


def submatrix(num_samples: int, split_factor: int = 1, shard_index: int = 0) -> Submatrix:
    """Submatrix(num_samples, split_factor, shard_index), cuking.cu:130-152 (+ flag validation :454-462)."""
    sm = Submatrix()
    check(capi.load().ck_submatrix_init(num_samples, split_factor, shard_index, C.byref(sm)))
    return sm


def num_shards(split_factor: int) -> int:
    """k(k+1)/2 upper-triangular shards, cloud_batch_submit.py:73."""
    return int(capi.load().ck_num_shards(split_factor))


def plan_work(num_samples: int, split_factor: int, num_gpus: int, first_shard: int = 0, num_run: int | None = None):
    """Work items (shard, part, num_parts, gpu) of a multi-shard run on the GPUs of one box, LPT-scheduled
    (ck_plan_work; the local counterpart of cloud_batch_submit.py:45,73)."""
    L = capi.load()
    if num_run is None:
        num_run = int(L.ck_num_shards(split_factor)) - first_shard
    n = C.c_uint32(0)
    check(L.ck_plan_work(num_samples, split_factor, first_shard, num_run, num_gpus, None, 0, C.byref(n)))
    items = (capi.WorkItem * max(1, n.value))()
    check(L.ck_plan_work(num_samples, split_factor, first_shard, num_run, num_gpus, items, n.value, C.byref(n)))
    return list(items[: n.value])


def words_per_sample(num_sites: int) -> int:
    """uint64 words per sample of the reference bit set, cuking.cu:498-500,:513."""
    return int(capi.load().ck_words_per_sample(num_sites))


def sample_offset(sm: Submatrix, sample: int) -> int:
    return int(capi.load().ck_submatrix_sample_offset(C.byref(sm), sample))


def num_samples(sm: Submatrix) -> int:
    return int(capi.load().ck_submatrix_num_samples(C.byref(sm)))


def synth_genotypes_host(seed: int, missing_rate: float, sample_begin: int, sample_end: int, site_begin: int,
                         site_end: int) -> np.ndarray:
    """Dense int8 genotypes (-1 = missing) of the synthetic cohort, shape [samples, sites]."""
    out = np.empty((sample_end - sample_begin, site_end - site_begin), dtype=np.int8)
    p = SynthParams(seed, missing_rate)
    check(capi.load().ck_synth_genotypes_host(C.byref(p), sample_begin, sample_end, site_begin, site_end,
                                              out.ctypes.data))
    return out


def rle_scan(data, bit_width: int, num_values: int, first_value: int = 0, payload_base: int = 0) -> np.ndarray:
    """ck_rle_scan: the run table (capi.RUN_DTYPE, no sentinel) of one Parquet RLE / bit-packed hybrid stream.  Host code of
    the library; needs no GPU."""
    buf = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
    runs = np.zeros(buf.size + 1, dtype=capi.RUN_DTYPE)
    n = C.c_uint32(0)
    check(capi.load().ck_rle_scan(buf.ctypes.data if buf.size else None, buf.size, bit_width, num_values, first_value, payload_base,
                                  runs.ctypes.data, len(runs), C.byref(n)))
    return runs[: n.value].copy()


def _ptr(x) -> tuple[int, bool]:
    """(address, on_device) of a numpy array or a torch tensor."""
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data, False
    # torch tensor (duck-typed so that torch stays optional for the pure-ABI tests)
    assert x.is_contiguous()
    return x.data_ptr(), bool(x.is_cuda)


class Context:
    """One GPU and one stream (ck_ctx)."""

    def __init__(self, device: int = 0, stream: int | None = None, king_variant: int | None = None):
        self._lib = capi.load()
        self._h = C.c_void_p()
        check(self._lib.ck_ctx_create(device, C.byref(self._h)))
        self.device = device
        self.stream = None  # cudaStream_t the ctx is bound to (None = its own)
        if stream is not None:
            self.set_stream(stream)
        if king_variant is not None:
            self.set_king_variant(king_variant)

    def set_stream(self, cuda_stream: int | None) -> None:
        check(self._lib.ck_ctx_set_stream(self._h, C.c_void_p(cuda_stream or 0)))
        self.stream = cuda_stream or None

    def fp4_selftest(self) -> tuple[bool, str]:
        """Runs the kind::mxf4 accumulation self-test on this GPU now: (exact, one-line report).  A failure routes
        variant 3 to the int8 kernel on this ctx (ck_ctx_fp4_selftest)."""
        exact = C.c_int(0)
        check(self._lib.ck_ctx_fp4_selftest(self._h, C.byref(exact)))
        return bool(exact.value), (self._lib.ck_last_error() or b"").decode()

    def screen_stats(self) -> dict:
        """ck_ctx_screen_stats: tiles the variant-5 screens ran over / flagged since the ctx was created, and the last level."""
        t, f, lv = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
        check(self._lib.ck_ctx_screen_stats(self._h, C.byref(t), C.byref(f), C.byref(lv)))
        return {"tiles_screened": int(t.value), "tiles_flagged": int(f.value), "level": int(lv.value)}

    def set_king_variant(self, variant: int) -> None:
        check(self._lib.ck_ctx_set_king_variant(self._h, variant))

    def synchronize(self) -> None:
        check(self._lib.ck_ctx_synchronize(self._h))

    def timings(self) -> dict:
        t = Timings()
        check(self._lib.ck_ctx_get_timings(self._h, C.byref(t)))
        return t.as_dict()

    def measure_int_peaks(self) -> dict:
        """Live POPC.32 / LOP3 issue-rate peaks of this GPU in lane-ops/s (ck_measure_int_peaks)."""
        popc, lop3 = C.c_double(), C.c_double()
        check(self._lib.ck_measure_int_peaks(self._h, C.byref(popc), C.byref(lop3)))
        return {"popc_lane_ops_per_s": popc.value, "lop3_lane_ops_per_s": lop3.value}

    def measure_fp4_peak(self) -> float:
        """Live dense kind::mxf4 tensor rate of this GPU in ops/s (ck_measure_fp4_peak)."""
        v = C.c_double()
        check(self._lib.ck_measure_fp4_peak(self._h, C.byref(v)))
        return v.value

    def measure_fp4_peak_sustained(self, seconds: float = 2.0) -> float:
        """The same rate sustained for `seconds` (second half timed): the board's power-limited clock included."""
        v = C.c_double()
        check(self._lib.ck_measure_fp4_peak_sustained(self._h, C.c_double(seconds), C.byref(v)))
        return v.value

    def planes(self, sm: Submatrix, num_sites: int) -> "Planes":
        return Planes(self, sm, num_sites)

    def king_host_bitset(self, n_samples: int, split_factor: int, shard_index: int, num_sites: int,
                         bit_set, kin_threshold: float, max_results: int = 10 << 20, out: np.ndarray | None = None,
                         part: tuple[int, int] = (0, 1)):
        """The reference seam in one call with host buffers (ck_king_host_bitset[_part]): bit set in the reference
        layout in, sorted KingResult records out.  `bit_set` may be a numpy array or a (pinned) CPU torch tensor;
        `part = (index, count)` evaluates one of `count` disjoint parts of the shard (one per GPU of a box)."""
        addr, on_device = _ptr(bit_set)
        assert not on_device
        res = out if out is not None else np.empty(max_results, dtype=RESULT_DTYPE)
        n = C.c_uint32(0)
        check(self._lib.ck_king_host_bitset_part(self._h, n_samples, split_factor, shard_index, num_sites, addr,
                                                 C.c_float(kin_threshold), max_results, res.ctypes.data, C.byref(n),
                                                 part[0], part[1]))
        return res[: n.value]

    def synth_triples_device(self, seed: int, missing_rate: float, sample_begin: int, sample_end: int,
                             site_begin: int, site_end: int):
        """Device pointers (row_idx, col_idx, n_alt_alleles, count) of the synthetic cohort's triples in Hail order."""
        p = SynthParams(seed, missing_rate)
        r, c, a, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_size_t()
        check(self._lib.ck_synth_triples_device(self._h, C.byref(p), sample_begin, sample_end, site_begin, site_end,
                                                C.byref(r), C.byref(c), C.byref(a), C.byref(n)))
        return r.value, c.value, a.value, n.value

    def close(self) -> None:
        if self._h:
            self._lib.ck_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def and_reduce(planes: list["Planes"]) -> None:
    """AND-all-reduce of the raw planes of one plane set per GPU over NVLink peer memory (ck_planes_and_reduce): after
    it every GPU holds the planes of all the triples that were dealt among them."""
    arr = (C.c_void_p * len(planes))(*[p._h for p in planes])
    check(capi.load().ck_planes_and_reduce(arr, len(planes)))


class Planes:
    """Device-resident bit planes of one shard (ck_planes) — the reference's `bit_set`, cuking.cu:513-523."""

    def __init__(self, ctx: Context, sm: Submatrix, num_sites: int):
        self._lib = ctx._lib
        self.ctx = ctx
        self.sm = sm
        self.num_sites = num_sites
        self._h = C.c_void_p()
        check(self._lib.ck_planes_create(ctx._h, C.byref(sm), num_sites, C.byref(self._h)))

    # -- filling ------------------------------------------------------------------------------------------------
    def reset(self) -> None:
        check(self._lib.ck_planes_reset(self._h))

    def pack(self, row_idx, col_idx, n_alt_alleles) -> None:
        """cuking.cu:675-703 on the GPU.  Arguments: int64, int64, int32 arrays (numpy = host, torch.cuda = device)."""
        if isinstance(row_idx, np.ndarray):
            row_idx = np.ascontiguousarray(row_idx, dtype=np.int64)
            col_idx = np.ascontiguousarray(col_idx, dtype=np.int64)
            n_alt_alleles = np.ascontiguousarray(n_alt_alleles, dtype=np.int32)
        (r, dev_r), (c, dev_c), (a, dev_a) = _ptr(row_idx), _ptr(col_idx), _ptr(n_alt_alleles)
        assert dev_r == dev_c == dev_a
        n = int(row_idx.shape[0])
        assert int(col_idx.shape[0]) == n and int(n_alt_alleles.shape[0]) == n
        check(self._lib.ck_pack_triples(self._h, r, c, a, n, int(dev_r)))

    def pack_narrow(self, row_idx, col_idx, n_alt_alleles) -> None:
        """The same for narrowed triples: uint32, uint32, uint8 arrays (ck_pack_triples_narrow; 9 bytes per triple)."""
        if isinstance(row_idx, np.ndarray):
            row_idx = np.ascontiguousarray(row_idx, dtype=np.uint32)
            col_idx = np.ascontiguousarray(col_idx, dtype=np.uint32)
            n_alt_alleles = np.ascontiguousarray(n_alt_alleles, dtype=np.uint8)
        (r, dev_r), (c, dev_c), (a, dev_a) = _ptr(row_idx), _ptr(col_idx), _ptr(n_alt_alleles)
        assert dev_r == dev_c == dev_a
        n = int(row_idx.shape[0])
        assert int(col_idx.shape[0]) == n and int(n_alt_alleles.shape[0]) == n
        check(self._lib.ck_pack_triples_narrow(self._h, r, c, a, n, int(dev_r)))

    def pack_encoded(self, columns, num_rows: int) -> None:
        """ck_pack_encoded: Parquet page payloads decoded and packed in one kernel.  `columns` = three dicts (row_idx,
        col_idx, n_alt_alleles) with `bytes` (uint8 array), `runs` (RUN_DTYPE array, sentinel included), `dict` (int64 /
        int32 array or None) and `skip`."""
        from .capi import RUN_DTYPE, EncodedColumn

        cols = (EncodedColumn * 3)()
        keep = []
        for c, col in zip(cols, columns):
            data = np.ascontiguousarray(col["bytes"], dtype=np.uint8)
            runs = np.ascontiguousarray(col["runs"], dtype=RUN_DTYPE)
            d = col.get("dict")
            width = int(col.get("value_width", d.dtype.itemsize if d is not None else 8))
            if d is not None:
                d = np.ascontiguousarray(d)
                assert d.dtype.itemsize == width
            keep += [data, runs, d]
            c.bytes, c.num_bytes = data.ctypes.data, data.size
            c.runs, c.num_runs = runs.ctypes.data, len(runs) - 1
            c.dict, c.dict_len = (d.ctypes.data, len(d)) if d is not None and len(d) else (None, 0)
            c.value_width, c.skip = width, int(col.get("skip", 0))
        check(self._lib.ck_pack_encoded(self._h, cols, int(num_rows)))

    def pack_device_ptrs(self, row_ptr: int, col_ptr: int, alt_ptr: int, n: int) -> None:
        check(self._lib.ck_pack_triples(self._h, row_ptr, col_ptr, alt_ptr, n, 1))

    def import_bitset(self, bit_set) -> None:
        addr, on_device = _ptr(bit_set)
        check(self._lib.ck_planes_import_bitset(self._h, addr, int(on_device)))

    def export_bitset(self) -> np.ndarray:
        out = np.empty(words_per_sample(self.num_sites) * num_samples(self.sm), dtype=np.uint64)
        check(self._lib.ck_planes_export_bitset(self._h, out.ctypes.data, 0))
        return out

    def synthesize(self, seed: int, missing_rate: float) -> None:
        p = SynthParams(seed, missing_rate)
        check(self._lib.ck_planes_synthesize(self._h, C.byref(p)))

    def finalize(self) -> None:
        check(self._lib.ck_planes_finalize(self._h))

    def device_bytes(self) -> int:
        b = C.c_uint64()
        check(self._lib.ck_planes_device_bytes(self._h, C.byref(b)))
        return int(b.value)

    # -- pairwise -----------------------------------------------------------------------------------------------
    def num_tiles(self) -> int:
        t = C.c_uint64()
        check(self._lib.ck_king_num_tiles(self._h, C.byref(t)))
        return int(t.value)

    # -- streaming delivery (ck_king_stream_*) -------------------------------------------------------------------
    def stream_begin(self, kin_threshold: float, max_results: int = 10 << 20, part: tuple[int, int] = (0, 1)) -> None:
        check(self._lib.ck_king_stream_begin(self._h, C.c_float(kin_threshold), max_results, part[0], part[1]))

    def stream_rows(self, rows, sample_begin: int, sample_end: int) -> None:
        """Reference-layout rows of the shard-local samples [sample_begin, sample_end) (numpy, torch CPU or CUDA tensor,
        or a raw device address as int); ranges must arrive in descending order."""
        addr, on_device = (rows, True) if isinstance(rows, int) else _ptr(rows)
        check(self._lib.ck_king_stream_rows(self._h, addr, 1 if on_device else 0, sample_begin, sample_end))

    def stream_end(self, max_results: int, out: np.ndarray | None = None):
        res = out if out is not None else np.empty(max_results, dtype=RESULT_DTYPE)
        n = C.c_uint32(0)
        check(self._lib.ck_king_stream_end(self._h, res.ctypes.data, C.byref(n)))
        self.last_count = int(n.value)
        return res[: n.value]

    def king_variant(self) -> int:
        """The pairwise kernel variant ck_king* runs on these planes (3 = FP4 tensor path, 2 = int8 beyond 2^23 sites)."""
        v = C.c_int()
        check(self._lib.ck_planes_king_variant(self._h, C.byref(v)))
        return int(v.value)

    def king(self, kin_threshold: float, max_results: int = 10 << 20, sort: bool = True,
             tiles: tuple[int, int] | None = None, out=None):
        """ComputeKingKernel + overflow check + sort (cuking.cu:713-765).  Returns the retained KingResult records
        (numpy structured array, or a slice count when `out` is a torch.cuda uint8/int32 buffer)."""
        n = C.c_uint32(0)
        if out is None:
            out = np.empty(max_results, dtype=RESULT_DTYPE)
        addr, on_device = _ptr(out)
        if tiles is None:
            rc = self._lib.ck_king(self._h, C.c_float(kin_threshold), max_results, addr, int(on_device), C.byref(n),
                                   int(sort))
        else:
            rc = self._lib.ck_king_tiles(self._h, tiles[0], tiles[1], C.c_float(kin_threshold), max_results, addr,
                                         int(on_device), C.byref(n), int(sort))
        self.last_count = int(n.value)
        check(rc)
        return out[: n.value] if isinstance(out, np.ndarray) else int(n.value)

    def king_view(self, view: Submatrix | None, kin_threshold: float, max_results: int = 10 << 20,
                  part: tuple[int, int] = (0, 1), sort: bool = True, out=None):
        """One part of a shard given as a VIEW into these planes (ck_king_view): `view` is any sub-matrix inside the
        planes' sample range (None = the planes' own), `part = (index, count)` the snake-dealt band partition across
        the GPUs of a box.  Negative thresholds with room for every pair take the dense (sort-free) output path."""
        n = C.c_uint32(0)
        if out is None:
            out = np.empty(max_results, dtype=RESULT_DTYPE)
        addr, on_device = _ptr(out)
        rc = self._lib.ck_king_view(self._h, C.byref(view) if view is not None else None, part[0], part[1],
                                    C.c_float(kin_threshold), max_results, addr, int(on_device), C.byref(n), int(sort))
        self.last_count = int(n.value)
        check(rc)
        return out[: n.value] if isinstance(out, np.ndarray) else int(n.value)

    def king_view_sink(self, view: Submatrix | None, kin_threshold: float, sink, max_results: int = 10 << 20,
                       part: tuple[int, int] = (0, 1), chunk_records: int = 0) -> int:
        """The same with the sorted records handed to `sink(records: np.ndarray)` chunk by chunk (ck_king_view_sink);
        the array is only valid during the call.  Returns the number of records delivered."""
        def trampoline(_user, ptr, count):
            try:
                buf = (C.c_char * (count * RESULT_DTYPE.itemsize)).from_address(ptr)
                sink(np.frombuffer(buf, dtype=RESULT_DTYPE, count=count))
                return 0
            except Exception:  # never unwind through the C frame
                import traceback

                traceback.print_exc()
                return 1

        cb = capi.RESULT_SINK(trampoline)
        n = C.c_uint64(0)
        check(self._lib.ck_king_view_sink(self._h, C.byref(view) if view is not None else None, part[0], part[1],
                                          C.c_float(kin_threshold), max_results, chunk_records, cb, None, C.byref(n)))
        return int(n.value)

    def counts(self, sample_i, sample_j) -> tuple[np.ndarray, np.ndarray]:
        """Raw six counters + kin for explicit pairs (parity hook, ck_king_counts)."""
        si = np.ascontiguousarray(sample_i, dtype=np.uint32)
        sj = np.ascontiguousarray(sample_j, dtype=np.uint32)
        counts = np.empty(si.size, dtype=COUNTS_DTYPE)
        kin = np.empty(si.size, dtype=np.float32)
        check(self._lib.ck_king_counts(self._h, si.ctypes.data, sj.ctypes.data, si.size, counts.ctypes.data,
                                       kin.ctypes.data))
        return counts, kin

    def close(self) -> None:
        if self._h:
            self._lib.ck_planes_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
