"""On-disk formats of the cuKING hot path without Hail.

write_input_dir  stands in for /root/reference/mt_to_cuking_inputs.py:28-47 — a directory of Parquet files with the
                 sparse (row_idx INT64, col_idx INT64, n_alt_alleles INT32) entries in Hail's order plus
                 metadata.json = {"num_sites": .., "samples": [..]}.
read_output_dir  stands in for /root/reference/cuking_outputs_to_ht.py:12-15 — reads every part-*.parquet of an
                 output directory, checks the schema the reference writes (cuking.cu:770-791) and returns one table
                 keyed by (i, j).
"""
from __future__ import annotations

import json
import os

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq

INPUT_SCHEMA = pa.schema([("row_idx", pa.int64()), ("col_idx", pa.int64()), ("n_alt_alleles", pa.int32())])
OUTPUT_FIELDS = [("i", pa.string()), ("j", pa.string()), ("kin", pa.float32()), ("ibs0", pa.int32()),
                 ("ibs1", pa.int32()), ("ibs2", pa.int32())]


def write_input_dir(path: str, genotypes: np.ndarray, sample_ids: list[str] | None = None, num_files: int = 16,
                    compression: str = "zstd", nullable: bool = True, row_group_size: int | None = None,
                    decoys: bool = True) -> dict:
    """genotypes[sample, site] in {0,1,2} or -1 (missing -> no entry).  Files split by site range like Spark
    partitions of a row-keyed MatrixTable; `nullable=True` writes OPTIONAL columns as Spark does."""
    os.makedirs(path, exist_ok=True)
    n_samples, n_sites = genotypes.shape
    sample_ids = sample_ids or [f"S{idx:07d}" for idx in range(n_samples)]
    assert len(sample_ids) == n_samples
    schema = INPUT_SCHEMA if nullable else pa.schema([pa.field(f.name, f.type, nullable=False) for f in INPUT_SCHEMA])
    bounds = np.linspace(0, n_sites, num_files + 1).astype(int)
    total = 0
    for f in range(num_files):
        lo, hi = int(bounds[f]), int(bounds[f + 1])
        site, sample = np.nonzero(genotypes[:, lo:hi].T >= 0)  # site-major, sample-minor
        table = pa.table({"row_idx": pa.array(site.astype(np.int64) + lo), "col_idx": pa.array(sample.astype(np.int64)),
                          "n_alt_alleles": pa.array(genotypes[sample, site + lo].astype(np.int32))}, schema=schema)
        pq.write_table(table, os.path.join(path, f"part-{f:05d}-synthetic.c000.zstd.parquet"), compression=compression,
                       row_group_size=row_group_size)
        total += len(site)
    with open(os.path.join(path, "metadata.json"), "w") as fh:
        json.dump({"num_sites": int(n_sites), "samples": sample_ids}, fh)  # mt_to_cuking_inputs.py:43-47
    if decoys:  # things Spark leaves behind that the reader must skip (cuking.cu:530-540)
        open(os.path.join(path, "_SUCCESS"), "w").close()
        os.makedirs(os.path.join(path, "_temporary", "0"), exist_ok=True)
        open(os.path.join(path, "_temporary", "0", "ignored.parquet"), "w").close()
    return {"num_triples": total, "sample_ids": sample_ids}


def read_output_dir(path: str, allow_row_groups: bool = False) -> pa.Table:
    """All part files of an output directory as one table; validates names, schema, codec and row-group layout."""
    parts = sorted(f for f in os.listdir(path) if f.endswith(".parquet"))
    if not parts:
        raise FileNotFoundError(f"no part files in {path}")
    tables = []
    for name in parts:
        if not (name.startswith("part-") and name.endswith(".snappy.parquet") and len(name) == len("part-00000.snappy.parquet")):
            raise ValueError(f"unexpected output file name {name} (cuking.cu:868-870)")
        pf = pq.ParquetFile(os.path.join(path, name))
        sch = pf.schema_arrow
        if [(f.name, f.type) for f in sch] != OUTPUT_FIELDS:
            raise ValueError(f"{name}: schema {sch} differs from cuking.cu:770-791")
        md = pf.metadata
        for c in range(md.num_columns):
            col = pf.schema.column(c)
            if col.max_definition_level != 0:
                raise ValueError(f"{name}: column {col.name} is not REQUIRED")
        if md.num_row_groups > 1 and not allow_row_groups:
            raise ValueError(f"{name}: {md.num_row_groups} row groups (the reference writes one, cuking.cu:804-805)")
        for rg in range(md.num_row_groups):
            for c in range(md.num_columns):
                if md.row_group(rg).column(c).compression != "SNAPPY":
                    raise ValueError(f"{name}: column {c} is not Snappy-compressed (cuking.cu:797-798)")
        tables.append(pf.read())
    return pa.concat_tables(tables)


# ---- encoded columns for ck_pack_encoded (synthetic inputs of bench.py / tools/decode_bench.py) ---------------------
def _varint(v: int) -> bytes:
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def bitpacked_stream(codes: np.ndarray, bit_width: int) -> bytes:
    """Parquet RLE / bit-packed hybrid stream made of bit-packed runs only: at most 63 groups of 8 values per run, LSB
    first, the last group zero-padded (what parquet-cpp / parquet-mr write for columns without long repeats)."""
    n = len(codes)
    v = np.concatenate([codes.astype(np.uint64), np.zeros((-n) % 8, dtype=np.uint64)])
    bits = ((v[:, None] >> np.arange(bit_width, dtype=np.uint64)) & np.uint64(1)).astype(np.uint8)
    packed = np.packbits(bits.reshape(-1), bitorder="little").tobytes()
    out = bytearray()
    groups = len(v) // 8
    for g0 in range(0, groups, 63):
        g = min(63, groups - g0)
        out += _varint((g << 1) | 1)
        out += packed[g0 * bit_width:(g0 + g) * bit_width]
    return bytes(out)


def rle_stream(codes: np.ndarray, bit_width: int) -> bytes:
    """The same encoding with one RLE run per stretch of equal values (row_idx of a site-major table)."""
    out = bytearray()
    edges = np.flatnonzero(np.diff(codes)) + 1
    starts = np.concatenate([[0], edges])
    ends = np.concatenate([edges, [len(codes)]])
    vb = (bit_width + 7) // 8
    for a, b in zip(starts, ends):
        out += _varint(int(b - a) << 1)
        out += int(codes[a]).to_bytes(vb, "little")
    return bytes(out)


def encoded_column(values: np.ndarray, width: int, rle: bool) -> dict:
    """One dictionary-encoded column of a window in the form Planes.pack_encoded takes (sorted dictionary, one stream)."""
    from .capi import RUN_DTYPE
    from .engine import rle_scan

    uniq, codes = np.unique(values, return_inverse=True)
    bw = max(1, int(len(uniq) - 1).bit_length())
    data = rle_stream(codes, bw) if rle else bitpacked_stream(codes, bw)
    runs = np.concatenate([rle_scan(data, bw, len(values)), np.array([(len(values), 0, 0, 0)], dtype=RUN_DTYPE)])
    return {"bytes": np.frombuffer(data, dtype=np.uint8), "runs": runs, "dict": uniq.astype(np.int64 if width == 8 else np.int32),
            "value_width": width, "skip": 0}
