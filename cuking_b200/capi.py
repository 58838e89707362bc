"""ctypes binding of libcuking_b200.so — the C ABI declared in include/cuking_b200.h.

This is plumbing only: device memory, kernels and result compaction all live behind the C ABI.  There is no CPU
fallback: if the shared library is missing or no CUDA device is present every device call raises CukingError.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcuking_b200.so")

CK_OK, CK_ERR_INVALID_ARGUMENT, CK_ERR_CUDA, CK_ERR_RESULT_OVERFLOW = 0, 1, 2, 3
CK_ERR_INVALID_GENOTYPE, CK_ERR_OUT_OF_RANGE, CK_ERR_OUT_OF_MEMORY = 4, 5, 6
_STATUS_NAMES = {1: "InvalidArgument", 2: "Cuda", 3: "ResourceExhausted", 4: "FailedPrecondition(InvalidGenotype)",
                 5: "OutOfRange", 6: "OutOfMemory"}


class CukingError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{_STATUS_NAMES.get(code, code)}: {message}")
        self.code = code
        self.message = message


class Submatrix(C.Structure):
    """struct Submatrix, /root/reference/cuking.cu:129-179."""

    _fields_ = [("i_begin", C.c_uint32), ("i_end", C.c_uint32), ("j_begin", C.c_uint32), ("j_end", C.c_uint32)]

    def __repr__(self):
        return f"Submatrix(rows=[{self.i_begin},{self.i_end}), cols=[{self.j_begin},{self.j_end}))"


class Counts(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("het_i", "het_j", "both_het", "opposing_hom", "concordant_hom", "shared_sites")]


class WorkItem(C.Structure):
    """ck_work_item: one (shard, part) assigned to a GPU by ck_plan_work."""

    _fields_ = [("shard_index", C.c_uint32), ("part_index", C.c_uint32), ("num_parts", C.c_uint32), ("gpu", C.c_uint32),
                ("pairs", C.c_uint64)]

    def __repr__(self):
        return f"WorkItem(shard={self.shard_index}, part={self.part_index}/{self.num_parts}, gpu={self.gpu}, pairs={self.pairs})"


class Run(C.Structure):
    """ck_run: one run of a Parquet RLE / bit-packed hybrid stream, or a PLAIN page (ck_pack_encoded)."""

    _fields_ = [("first_value", C.c_uint32), ("kind", C.c_uint32), ("bit_width", C.c_uint32), ("payload", C.c_uint32)]


RUN_DTYPE = np.dtype([("first_value", "<u4"), ("kind", "<u4"), ("bit_width", "<u4"), ("payload", "<u4")])
CK_RUN_RLE, CK_RUN_BITPACKED, CK_RUN_PLAIN = 0, 1, 2


class EncodedColumn(C.Structure):
    """ck_encoded_column: one column of one window of rows as page payloads + run table + dictionary."""

    _fields_ = [("bytes", C.c_void_p), ("num_bytes", C.c_uint64), ("runs", C.c_void_p), ("num_runs", C.c_uint32),
                ("dict", C.c_void_p), ("dict_len", C.c_uint32), ("value_width", C.c_uint32), ("skip", C.c_uint32)]


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("missing_rate", C.c_double)]


class Timings(C.Structure):
    _fields_ = [("pack_ms", C.c_float), ("finalize_ms", C.c_float), ("import_ms", C.c_float), ("king_ms", C.c_float),
                ("sort_ms", C.c_float), ("h2d_ms", C.c_float), ("d2h_ms", C.c_float), ("king_launches", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# struct KingResult, cuking.cu:182-186
RESULT_DTYPE = np.dtype(
    [("sample_i", "<u4"), ("sample_j", "<u4"), ("kin", "<f4"), ("ibs0", "<u4"), ("ibs1", "<u4"), ("ibs2", "<u4")]
)
COUNTS_DTYPE = np.dtype([(n, "<u4") for n, _ in Counts._fields_])

# every symbol include/cuking_b200.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    "ck_abi_version", "ck_last_error", "ck_device_count", "ck_submatrix_init", "ck_num_shards",
    "ck_submatrix_num_rows", "ck_submatrix_num_cols", "ck_submatrix_num_samples", "ck_submatrix_contains",
    "ck_submatrix_sample_offset", "ck_words_per_sample", "ck_ctx_create", "ck_ctx_set_stream",
    "ck_ctx_set_king_variant", "ck_ctx_synchronize", "ck_ctx_get_timings", "ck_measure_int_peaks", "ck_ctx_destroy", "ck_planes_create",
    "ck_planes_reset", "ck_planes_destroy", "ck_planes_finalize", "ck_planes_num_sites", "ck_planes_device_bytes",
    "ck_pack_triples", "ck_host_alloc", "ck_host_free", "ck_planes_import_bitset", "ck_planes_export_bitset", "ck_planes_synthesize", "ck_king",
    "ck_king_num_tiles", "ck_planes_king_variant", "ck_king_tiles", "ck_king_counts", "ck_king_host_bitset", "ck_king_host_bitset_part", "ck_king_stream_granularity", "ck_king_stream_begin",
    "ck_king_stream_rows", "ck_king_stream_end", "ck_synth_genotypes_host",
    "ck_synth_triples_device", "ck_ctx_fp4_selftest", "ck_planes_and_reduce", "ck_king_view", "ck_king_view_sink", "ck_plan_work", "ck_measure_fp4_peak", "ck_measure_fp4_peak_sustained", "ck_pack_triples_narrow",
    "ck_rle_scan", "ck_pack_encoded", "ck_ctx_screen_stats",
]

# typedef int (*ck_result_sink)(void *user, const ck_result *records, size_t count)
RESULT_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t)

_lib = None


def load() -> C.CDLL:
    """Loads libcuking_b200.so (built in-tree by __graft_entry__.build / cuking_b200/csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CukingError(CK_ERR_CUDA, f"{LIB_PATH} is missing - run `python -c 'import __graft_entry__ as g; "
                                       f"g.build()'` or `make -C cuking_b200/csrc`; there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    u32, u64, vp, f32, i32 = C.c_uint32, C.c_uint64, C.c_void_p, C.c_float, C.c_int
    SMp = C.POINTER(Submatrix)
    sig = {
        "ck_abi_version": ([], i32), "ck_last_error": ([], C.c_char_p), "ck_device_count": ([C.POINTER(i32)], i32),
        "ck_submatrix_init": ([u32, u32, u32, SMp], i32), "ck_num_shards": ([u32], u32),
        "ck_submatrix_num_rows": ([SMp], u32), "ck_submatrix_num_cols": ([SMp], u32),
        "ck_submatrix_num_samples": ([SMp], u32), "ck_submatrix_contains": ([SMp, u32], u32),
        "ck_submatrix_sample_offset": ([SMp, u32], u32), "ck_words_per_sample": ([u32], u32),
        "ck_ctx_create": ([i32, C.POINTER(vp)], i32), "ck_ctx_set_stream": ([vp, vp], i32),
        "ck_ctx_set_king_variant": ([vp, i32], i32), "ck_ctx_synchronize": ([vp], i32),
        "ck_ctx_get_timings": ([vp, C.POINTER(Timings)], i32), "ck_ctx_destroy": ([vp], i32),
        "ck_measure_int_peaks": ([vp, C.POINTER(C.c_double), C.POINTER(C.c_double)], i32),
        "ck_planes_create": ([vp, SMp, u32, C.POINTER(vp)], i32), "ck_planes_reset": ([vp], i32),
        "ck_planes_destroy": ([vp], i32), "ck_planes_finalize": ([vp], i32),
        "ck_planes_num_sites": ([vp, C.POINTER(u32)], i32), "ck_planes_device_bytes": ([vp, C.POINTER(u64)], i32),
        "ck_pack_triples": ([vp, vp, vp, vp, C.c_size_t, i32], i32),
        "ck_pack_triples_narrow": ([vp, vp, vp, vp, C.c_size_t, i32], i32),
        "ck_rle_scan": ([vp, C.c_size_t, u32, u32, u32, u32, vp, u32, C.POINTER(u32)], i32),
        "ck_pack_encoded": ([vp, C.POINTER(EncodedColumn), u32], i32),
        "ck_host_alloc": ([C.c_size_t, C.POINTER(vp)], i32), "ck_host_free": ([vp], i32),
        "ck_planes_import_bitset": ([vp, vp, i32], i32), "ck_planes_export_bitset": ([vp, vp, i32], i32),
        "ck_planes_synthesize": ([vp, C.POINTER(SynthParams)], i32),
        "ck_king": ([vp, f32, u32, vp, i32, C.POINTER(u32), i32], i32),
        "ck_king_num_tiles": ([vp, C.POINTER(u64)], i32),
        "ck_planes_king_variant": ([vp, C.POINTER(i32)], i32),
        "ck_king_tiles": ([vp, u64, u64, f32, u32, vp, i32, C.POINTER(u32), i32], i32),
        "ck_king_counts": ([vp, vp, vp, C.c_size_t, vp, vp], i32),
        "ck_king_host_bitset": ([vp, u32, u32, u32, u32, vp, f32, u32, vp, C.POINTER(u32)], i32),
        "ck_king_host_bitset_part": ([vp, u32, u32, u32, u32, vp, f32, u32, vp, C.POINTER(u32), u32, u32], i32),
        "ck_king_stream_granularity": ([], u32), "ck_king_stream_begin": ([vp, f32, u32, u32, u32], i32),
        "ck_king_stream_rows": ([vp, vp, i32, u32, u32], i32), "ck_king_stream_end": ([vp, vp, C.POINTER(u32)], i32),
        "ck_measure_fp4_peak": ([vp, C.POINTER(C.c_double)], i32),
        "ck_measure_fp4_peak_sustained": ([vp, C.c_double, C.POINTER(C.c_double)], i32),
        "ck_plan_work": ([u32, u32, u32, u32, u32, C.POINTER(WorkItem), u32, C.POINTER(u32)], i32),
        "ck_ctx_fp4_selftest": ([vp, C.POINTER(i32)], i32),
        "ck_ctx_screen_stats": ([vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(i32)], i32),
        "ck_planes_and_reduce": ([C.POINTER(vp), u32], i32),
        "ck_king_view": ([vp, SMp, u32, u32, f32, u32, vp, i32, C.POINTER(u32), i32], i32),
        "ck_king_view_sink": ([vp, SMp, u32, u32, f32, u32, C.c_size_t, RESULT_SINK, vp, C.POINTER(u64)], i32),
        "ck_synth_genotypes_host": ([C.POINTER(SynthParams), u32, u32, u32, u32, vp], i32),
        "ck_synth_triples_device": ([vp, C.POINTER(SynthParams), u32, u32, u32, u32, C.POINTER(vp), C.POINTER(vp),
                                     C.POINTER(vp), C.POINTER(C.c_size_t)], i32),
    }
    for name, (argtypes, restype) in sig.items():
        fn = getattr(L, name)
        fn.argtypes, fn.restype = argtypes, restype
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != CK_OK:
        raise CukingError(rc, (load().ck_last_error() or b"").decode())
