#!/usr/bin/env python3
"""Device-side rate of ck_pack_encoded: windows of dictionary-encoded triples in Hail's order, built here
with a plain numpy writer (bit-packed runs of 63 groups for col_idx / n_alt_alleles, RLE runs for row_idx - what parquet-cpp
and parquet-mr emit for such columns), decoded and packed by decode_pack_kernel.

  python tools/decode_bench.py [--samples 100000] [--rows 1048576] [--windows 8] [--reps 5]

Prints one JSON line: rows per second of the call as the library times it (upload of the window + kernel, CUDA events) and
the bytes per row that crossed PCIe.  For an ncu capture of the kernel alone run it with --reps 1 --windows 1 under ncu.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cuking_b200 as ck  # noqa: E402
from cuking_b200.io import encoded_column as column  # noqa: E402


ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=100_000)
ap.add_argument("--rows", type=int, default=1 << 20)
ap.add_argument("--windows", type=int, default=8)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()

rng = np.random.default_rng(42)
sites = args.windows * (args.rows // args.samples + 2)
windows = []
bytes_per_row = 0.0
row0 = 0
for w in range(args.windows):  # site-major, sample-minor rows (Hail order), 1 % of the genotypes missing
    idx = np.arange(row0, row0 + int(args.rows * 1.0101))
    idx = idx[rng.random(len(idx)) >= 0.01][: args.rows]
    row0 = int(idx[-1]) + 1
    site, sample = idx // args.samples, idx % args.samples
    alt = rng.choice(np.array([0, 1, 2]), size=len(idx), p=[0.55, 0.35, 0.10])
    cols = [column(site, 8, True), column(sample, 8, False), column(alt, 4, False)]
    bytes_per_row += sum(c["bytes"].nbytes + c["runs"].nbytes + c["dict"].nbytes for c in cols) / len(idx) / args.windows
    windows.append((cols, len(idx)))

with ck.Context(0) as ctx, ctx.planes(ck.submatrix(args.samples), sites + 1) as pl:
    best = None
    for rep in range(args.reps):
        total_ms = 0.0
        for cols, n in windows:
            pl.pack_encoded(cols, n)
            total_ms += ctx.timings()["pack_ms"]
        best = total_ms if best is None else min(best, total_ms)
    rows = sum(n for _, n in windows)
    print(json.dumps({"tool": "decode_bench", "kernel": "decode_pack_kernel", "samples": args.samples, "rows_per_window": args.rows,
                      "windows": args.windows, "ms_per_window": best / args.windows, "rows_per_s": rows / (best * 1e-3),
                      "bytes_per_row_over_pcie": bytes_per_row, "upload_gbs": rows * bytes_per_row / (best * 1e-3) / 1e9,
                      "timed": "cudaEvents around [upload of the window (pageable numpy arrays here) + decode_pack_kernel], best of %d" % args.reps}))
