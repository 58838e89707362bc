"""cuking_b200 — B200-native pairwise KING relatedness (the hot path of populationgenomics/cuKING).

Host-side mirror of the reference's seam (/root/reference/cuking.cu): `Submatrix` (:129-179), the packed bit set
(`Planes`, :507-523, :675-703) and the kernel launch (`Planes.king`, :191-314, :734-765), over the C ABI in
include/cuking_b200.h.  The compute is hand-written CUDA for sm_100a in cuking_b200/csrc; Python only moves
pointers.  No CPU fallback exists.
"""
from .capi import CukingError, Submatrix, RESULT_DTYPE, COUNTS_DTYPE, SynthParams  # noqa: F401
from .engine import Context, Planes, submatrix, num_shards, words_per_sample, synth_genotypes_host, and_reduce, plan_work, rle_scan  # noqa: F401

__all__ = ["CukingError", "Submatrix", "RESULT_DTYPE", "COUNTS_DTYPE", "SynthParams", "Context", "Planes",
           "submatrix", "num_shards", "words_per_sample", "synth_genotypes_host", "and_reduce", "plan_work", "rle_scan"]
