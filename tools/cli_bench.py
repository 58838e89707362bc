#!/usr/bin/env python3
"""End-to-end timing of bin/cuking through real Parquet files (decode -> pack -> pairwise -> write).

  python tools/cli_bench.py [--samples 2000] [--sites 100000] [--files 32] [--threads 16]

Writes a synthetic cohort (SURVEY.md §8d generator) as zstd Parquet part files in Hail's layout, runs the binary, checks
the output row count against the library called directly, and prints one JSON line with the phase times the binary logs.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import cuking_b200 as ck  # noqa: E402
from cuking_b200 import io as ckio  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=2000)
ap.add_argument("--sites", type=int, default=100_000)
ap.add_argument("--files", type=int, default=32)
ap.add_argument("--threads", type=int, default=16)
ap.add_argument("--threshold", type=float, default=0.0884)
ap.add_argument("--num-gpus", default="1", help="comma list: the binary is run once per entry on the same input")
ap.add_argument("--split-factor", type=int, default=1, help="> 1: --all_shards with this split factor")
ap.add_argument("--repeat", type=int, default=1, help="runs per configuration (page cache warm after the first)")
ap.add_argument("--modes", default="device", help="comma list of device (pages decoded on the GPU, the default of the binary), "
                "narrow (CUKING_HOST_DECODE=1: libparquet decode + 9-byte triples), wide (that + CUKING_WIDE_TRIPLES=1)")
ap.add_argument("--window-rows", type=int, default=0, help="CUKING_DECODE_WINDOW_ROWS for the device mode")
args = ap.parse_args()

with tempfile.TemporaryDirectory() as tmp:
    t0 = time.perf_counter()
    from oracle import king_oracle as ko  # the OpenMP restatement of the same generator (tests pin the two against each other)

    g = ko.synth_genotypes(42, 0.01, 0, args.samples, 0, args.sites)
    info = ckio.write_input_dir(os.path.join(tmp, "in"), g, num_files=args.files)
    gen_s = time.perf_counter() - t0
    in_bytes = sum(os.path.getsize(os.path.join(tmp, "in", f)) for f in os.listdir(os.path.join(tmp, "in"))
                   if f.endswith(".parquet"))
    with ck.Context(0) as ctx, ctx.planes(ck.submatrix(args.samples), args.sites) as pl:
        pl.synthesize(42, 0.01)
        want = len(pl.king(args.threshold, 10 << 20))
    runs = [(int(x), m, r) for x in args.num_gpus.split(",") for r in range(args.repeat) for m in args.modes.split(",")]  # modes interleaved
    for gpus, mode, rep in runs:
        out = f"{tmp}/out{gpus}{mode}{rep}"
        cmd = [os.path.join(ROOT, "bin", "cuking"), f"--input_uri={tmp}/in", f"--output_uri={out}",
               f"--kin_threshold={args.threshold}", f"--num_reader_threads={args.threads}", f"--num_gpus={gpus}"]
        if args.split_factor > 1:
            cmd += [f"--split_factor={args.split_factor}", "--all_shards", "--write_success_file"]
        t0 = time.perf_counter()
        env = dict(os.environ)
        env["CUKING_INGEST_STATS"] = "1"
        if mode in ("narrow", "wide"):
            env["CUKING_HOST_DECODE"] = "1"
        if mode == "wide":
            env["CUKING_WIDE_TRIPLES"] = "1"
        if mode == "device" and args.window_rows:
            env["CUKING_DECODE_WINDOW_ROWS"] = str(args.window_rows)
        p = subprocess.run(cmd, capture_output=True, text=True, env=env)
        wall = time.perf_counter() - t0
        if p.returncode != 0:
            sys.exit(p.stderr)
        phases = dict(re.findall(r"^(Reading metadata|Initializing CUDA|Allocating memory for bit set|Listing input files|Processing Parquet tables|"
                                 r"Exchanging bit sets[^.]*|Computed [^(]*)(?:\.\.\.)?.*\(([^;)]+)\)$", p.stdout, re.M))
        kernels = re.findall(r"^Running KING CUDA kernel for (.*?)\.\.\. \(([^;]+); kernel ([0-9.e+]+) ms on GPU (\d+)\)", p.stdout, re.M)
        rows = ckio.read_output_dir(out, allow_row_groups=True).num_rows
        assert rows == want, (rows, want)
        decode_s = None
        m = re.search(r"Processing Parquet tables\.\.\..* entries \(([0-9.]+)(ms|s)\)", p.stdout)
        if m:
            decode_s = float(m.group(1)) * (1e-3 if m.group(2) == "ms" else 1.0)
        print(json.dumps({"tool": "cli_bench", "samples": args.samples, "sites": args.sites, "triples": info["num_triples"],
                          "parquet_bytes": in_bytes, "files": args.files, "reader_threads": args.threads, "num_gpus": gpus, "triples_mode": mode, "rep": rep,
                          "split_factor": args.split_factor, "generate_input_s": round(gen_s, 2), "cuking_wall_s": round(wall, 3),
                          "phases": phases, "kernels": [{"what": k[0], "wall": k[1], "kernel_ms": float(k[2]), "gpu": int(k[3])} for k in kernels],
                          "triples_per_s_end_to_end": info["num_triples"] / wall,
                          "triples_per_s_decode_and_pack": (info["num_triples"] / decode_s) if decode_s else None,
                          "retained_pairs": rows,
                          "ingest_stats": (re.search(r"^Ingest thread-seconds.*$", p.stdout, re.M) or [None])[0]}), flush=True)
