"""ctypes front-end of oracle/_ref/libcuking_ref.so — the reference's own ComputeKingKernel, sliced from
/root/reference/cuking.cu:100-314 and compiled for sm_100a by oracle/build_ref.sh.  TEST / BASELINE ONLY.
Needs a GPU to run; the .so is prebuilt in the CPU container and travels to the GPU box."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .king_oracle import RESULT_DTYPE

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libcuking_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(_PATH)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(_PATH)
        u32, f32 = C.c_uint32, C.c_float
        L.ref_king.argtypes = [u32, u32, u32, u32, C.c_void_p, f32, u32, C.c_void_p, C.POINTER(u32), C.POINTER(u32),
                               C.POINTER(f32), C.c_int]
        L.ref_king.restype = C.c_int
        L.ref_king_device_resident.argtypes = [u32, u32, u32, u32, C.c_void_p, f32, u32, C.c_void_p, C.c_void_p,
                                               C.POINTER(f32)]
        L.ref_king_device_resident.restype = C.c_int
        _lib = L
    return _lib


def king(bit_set: np.ndarray, num_samples: int, split_factor: int, shard_index: int, words_per_sample: int,
         kin_threshold: float, max_results: int, managed: bool = False):
    """Runs the reference kernel; returns (sorted results, count, overflow, kernel_ms)."""
    res = np.zeros(max_results, dtype=RESULT_DTYPE)
    cnt, ovf, ms = C.c_uint32(0), C.c_uint32(0), C.c_float(0)
    rc = lib().ref_king(num_samples, split_factor, shard_index, words_per_sample, bit_set.ctypes.data,
                        C.c_float(kin_threshold), max_results, res.ctypes.data, C.byref(cnt), C.byref(ovf),
                        C.byref(ms), 1 if managed else 0)
    if rc != 0:
        raise RuntimeError("reference kernel harness failed")
    return res[: min(cnt.value, max_results)], int(cnt.value), bool(ovf.value), float(ms.value)
