"""ctypes front-end of the CPU oracle (oracle/king_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this module.
The product package (cuking_b200/) never does; it fails loudly when its CUDA library is missing.

The C file restates /root/reference/cuking.cu:129-179 (Submatrix), :496-523 + :675-703 (bit-set layout and
pack), :191-314 (ComputeKingKernel) and :761-765 (result sort); see its header for the parity-pinning note.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")


class Submatrix(C.Structure):
    """cuking.cu:129-179."""

    _fields_ = [("i_begin", C.c_uint32), ("i_end", C.c_uint32), ("j_begin", C.c_uint32), ("j_end", C.c_uint32)]


class Counts(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("het_i", "het_j", "both_het", "opposing_hom", "concordant_hom", "shared_sites")]


RESULT_DTYPE = np.dtype(
    [("sample_i", "<u4"), ("sample_j", "<u4"), ("kin", "<f4"), ("ibs0", "<u4"), ("ibs1", "<u4"), ("ibs2", "<u4")]
)  # cuking.cu:182-186, 24 bytes
assert RESULT_DTYPE.itemsize == 24


def _cpu_tag() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return hashlib.sha1(line.encode()).hexdigest()[:10]
    except OSError:
        pass
    return "unknown"


def _make(target: str, out_name: str, arch: str | None = None) -> str:
    out = os.path.join(_BUILD, out_name)
    src = os.path.join(_HERE, "king_oracle.c")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(_BUILD, exist_ok=True)
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        cmd = [cc, "-O3", f"-march={arch}", "-fopenmp", "-ffp-contract=off", "-fPIC", "-std=c11", "-shared",
               "-o", out, src]
        subprocess.run(cmd, check=True, capture_output=True)
    return out


def build(native: bool = False) -> str:
    """Compile the oracle (portable x86-64-v3 build, or a -march=native build keyed by this host's CPU flags)."""
    if native:
        return _make("native", f"libking_oracle_native_{_cpu_tag()}.so", "native")
    return _make("all", "libking_oracle.so", "x86-64-v3")


_LIBS: dict[bool, C.CDLL] = {}


def lib(native: bool = False) -> C.CDLL:
    if native in _LIBS:
        return _LIBS[native]
    L = C.CDLL(build(native))
    u32, u64p, f32 = C.c_uint32, C.POINTER(C.c_uint64), C.c_float
    SMp = C.POINTER(Submatrix)
    L.ko_submatrix_init.argtypes = [SMp, u32, u32, u32]
    L.ko_submatrix_init.restype = None
    for name in ("ko_num_rows", "ko_num_cols", "ko_num_samples"):
        getattr(L, name).argtypes = [SMp]
        getattr(L, name).restype = u32
    for name in ("ko_contains", "ko_sample_offset"):
        getattr(L, name).argtypes = [SMp, u32]
        getattr(L, name).restype = u32
    L.ko_padded_sites.argtypes = [u32]
    L.ko_padded_sites.restype = u32
    L.ko_words_per_sample.argtypes = [u32]
    L.ko_words_per_sample.restype = u32
    L.ko_bitset_init.argtypes = [C.c_void_p, C.c_size_t]
    L.ko_bitset_init.restype = None
    L.ko_pack.argtypes = [C.c_void_p, u32, SMp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.ko_pack.restype = C.c_int64
    L.ko_pair_counts.argtypes = [C.c_void_p, u32, u32, u32, C.POINTER(Counts), C.POINTER(f32)]
    L.ko_pair_counts.restype = None
    L.ko_king.argtypes = [SMp, u32, C.c_void_p, f32, u32, C.c_void_p, C.POINTER(u32), C.POINTER(u32)]
    L.ko_king.restype = None
    L.ko_sort.argtypes = [C.c_void_p, C.c_size_t]
    L.ko_sort.restype = None
    L.ko_bench_rect.argtypes = [C.c_void_p, u32, u32, u32, u32, u32, f32]
    L.ko_bench_rect.restype = C.c_uint64
    L.ko_num_threads.argtypes = []
    L.ko_num_threads.restype = C.c_int
    L.ko_set_num_threads.argtypes = [C.c_int]
    L.ko_set_num_threads.restype = None
    L.ko_synth_genotypes.argtypes = [C.c_uint64, C.c_double, u32, u32, u32, u32, C.c_void_p]
    L.ko_synth_genotypes.restype = None
    L.ko_synth_bitset.argtypes = [C.c_uint64, C.c_double, u32, u32, u32, C.c_void_p]
    L.ko_synth_bitset.restype = None
    _LIBS[native] = L
    return L


# ---- numpy-level helpers -------------------------------------------------------------------------------------


def submatrix(num_samples: int, split_factor: int = 1, shard_index: int = 0) -> Submatrix:
    sm = Submatrix()
    lib().ko_submatrix_init(C.byref(sm), num_samples, split_factor, shard_index)
    return sm


def words_per_sample(num_sites: int) -> int:
    L = lib()
    return int(L.ko_words_per_sample(L.ko_padded_sites(num_sites)))


def new_bitset(sm: Submatrix, num_sites: int) -> np.ndarray:
    """All-ones (= all missing) bit set in the reference layout, cuking.cu:513-523."""
    wps = words_per_sample(num_sites)
    bs = np.empty(wps * int(lib().ko_num_samples(C.byref(sm))), dtype=np.uint64)
    lib().ko_bitset_init(bs.ctypes.data, bs.size)
    return bs


def pack(bit_set: np.ndarray, num_sites: int, sm: Submatrix, row_idx, col_idx, n_alt) -> int:
    row_idx = np.ascontiguousarray(row_idx, dtype=np.int64)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.int64)
    n_alt = np.ascontiguousarray(n_alt, dtype=np.int32)
    assert row_idx.size == col_idx.size == n_alt.size
    return int(lib().ko_pack(bit_set.ctypes.data, words_per_sample(num_sites), C.byref(sm), row_idx.ctypes.data,
                             col_idx.ctypes.data, n_alt.ctypes.data, row_idx.size))


def pack_dense(genotypes: np.ndarray, sm: Submatrix | None = None) -> tuple[np.ndarray, Submatrix]:
    """genotypes[sample, site] in {0,1,2} or -1 (= missing, i.e. absent from the triples)."""
    n, s = genotypes.shape
    sm = sm or submatrix(n)
    bs = new_bitset(sm, s)
    col, row = np.nonzero(genotypes >= 0)
    bad = pack(bs, s, sm, row, col, genotypes[col, row])
    assert bad == -1
    return bs, sm


def pair_counts(bit_set: np.ndarray, num_sites: int, slot_i: int, slot_j: int) -> tuple[dict, float]:
    c, kin = Counts(), C.c_float()
    lib().ko_pair_counts(bit_set.ctypes.data, words_per_sample(num_sites), slot_i, slot_j, C.byref(c), C.byref(kin))
    return {n: int(getattr(c, n)) for n, _ in Counts._fields_}, float(kin.value)


def king(bit_set: np.ndarray, num_sites: int, sm: Submatrix, kin_threshold: float, max_results: int = 10 << 20,
         sort: bool = True, native: bool = False) -> tuple[np.ndarray, int, bool]:
    """Returns (results[:min(count, max_results)], count, overflow) — cuking.cu:191-314 (+ :761-765 if sort)."""
    L = lib(native)
    res = np.zeros(max_results, dtype=RESULT_DTYPE)  # zero-initialised like cuking.cu:719
    idx, ovf = C.c_uint32(0), C.c_uint32(0)
    L.ko_king(C.byref(sm), words_per_sample(num_sites), bit_set.ctypes.data, C.c_float(kin_threshold), max_results,
              res.ctypes.data, C.byref(idx), C.byref(ovf))
    n = min(int(idx.value), max_results)
    res = res[:n].copy()
    if sort:
        L.ko_sort(res.ctypes.data, n)
    return res, int(idx.value), bool(ovf.value)


def synth_genotypes(seed: int, missing_rate: float, sample_begin: int, sample_end: int, site_begin: int, site_end: int,
                    native: bool = False) -> np.ndarray:
    """Dense int8 genotypes (-1 = missing) of the bench's synthetic cohort, [samples, sites]."""
    out = np.empty((sample_end - sample_begin, site_end - site_begin), dtype=np.int8)
    lib(native).ko_synth_genotypes(seed, missing_rate, sample_begin, sample_end, site_begin, site_end, out.ctypes.data)
    return out


def synth_bitset(seed: int, missing_rate: float, sample_begin: int, sample_end: int, num_sites: int,
                 native: bool = False) -> np.ndarray:
    """Reference-layout bit set of samples [sample_begin, sample_end) of the synthetic cohort, slots 0 .. n-1."""
    bs = np.empty(words_per_sample(num_sites) * (sample_end - sample_begin), dtype=np.uint64)
    lib(native).ko_synth_bitset(seed, missing_rate, sample_begin, sample_end, num_sites, bs.ctypes.data)
    return bs
