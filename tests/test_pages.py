"""Parquet pages decoded on the device: ck_rle_scan / ck_pack_encoded and the host reader that
feeds them (cuking_b200/host/parquet_io.cc ReadEncoded), against oracle/parquet_pages.py - which is itself pinned here
against pyarrow's reader on files pyarrow writes - and against the oracle's pack (cuking.cu:675-703)."""
from __future__ import annotations

import os
import subprocess

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import cuking_b200 as ck
from cuking_b200 import capi
from cuking_b200 import io as ckio
from oracle import king_oracle as ko
from oracle import parquet_pages as pp
from tests.helpers import oracle_bitset, random_genotypes, triples_of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK = os.path.join(ROOT, "bin", "encoded_check")

# writer settings a cuKING input may arrive with (the producer is Spark / parquet-mr with zstd, mt_to_cuking_inputs.py:28-31)
WRITERS = {
    "zstd_v1": dict(compression="zstd"),
    "snappy_v2_pages": dict(compression="snappy", data_page_version="2.0"),
    "uncompressed_small_pages": dict(compression="none", data_page_size=4096),
    "gzip_plain": dict(compression="gzip", use_dictionary=False),
    "dictionary_fallback": dict(compression="zstd", dictionary_pagesize_limit=2048, data_page_size=8192),
    "many_row_groups": dict(compression="zstd", row_group_size=3000),
}


def triple_table(g: np.ndarray, nullable: bool = True) -> pa.Table:
    site, sample, alt = triples_of(g)
    schema = pa.schema([pa.field("row_idx", pa.int64(), nullable), pa.field("col_idx", pa.int64(), nullable),
                        pa.field("n_alt_alleles", pa.int32(), nullable)])
    return pa.table({"row_idx": site, "col_idx": sample, "n_alt_alleles": alt}, schema=schema)


def write_variants(tmp_path, g: np.ndarray) -> dict[str, str]:
    out = {}
    for i, (name, kw) in enumerate(WRITERS.items()):
        path = str(tmp_path / f"{name}.parquet")
        pq.write_table(triple_table(g, nullable=(i % 2 == 0)), path, **kw)
        out[name] = path
    return out


def window_column(pages: dict, row_begin: int, row_end: int) -> dict:
    """One column of Planes.pack_encoded for the rows [row_begin, row_end) of a column chunk: every page that overlaps
    the range, described as a whole (like EncodedWindow in the C++ host); run tables from the library's ck_rle_scan."""
    payload = bytearray()
    runs = []
    first = None
    v = table_values = 0
    width = pages["value_width"]
    for p in pages["pages"]:
        a, b = v, v + p["num_values"]
        v = b
        if b <= row_begin or a >= row_end:
            continue
        if first is None:
            first = a
        payload.extend(b"\0" * (-len(payload) % 8))
        off = len(payload)
        if p["encoding"] in (pp.ENC_RLE_DICTIONARY, pp.ENC_PLAIN_DICTIONARY):
            stream = p["values"][1:]
            payload += stream
            runs.append(ck.rle_scan(stream, p["values"][0], p["num_values"], table_values, off))
        else:
            assert p["encoding"] == pp.ENC_PLAIN
            payload += p["values"][: p["num_values"] * width]
            runs.append(np.array([(table_values, capi.CK_RUN_PLAIN, 0, off)], dtype=capi.RUN_DTYPE))
        table_values += p["num_values"]
    runs.append(np.array([(table_values, 0, 0, 0)], dtype=capi.RUN_DTYPE))  # sentinel
    return {"bytes": np.frombuffer(bytes(payload), dtype=np.uint8), "runs": np.concatenate(runs), "dict": pages["dictionary"],
            "value_width": width, "skip": row_begin - first}


# ---- CPU: the oracle's page layer against pyarrow, the library's run scan against the oracle ------------------------
def test_hybrid_codec_round_trip_and_scan_matches_oracle():
    rng = np.random.default_rng(11)
    for bw in (0, 1, 2, 3, 7, 8, 12, 13, 17, 24, 31, 32):
        for n in (1, 7, 8, 9, 63, 64, 1000, 5003):
            v = rng.integers(0, 1 << bw, n, dtype=np.uint64) if bw else np.zeros(n, dtype=np.uint64)
            if n > 50:
                v[10:40] = v[10]  # a repeat long enough for an RLE run
                v[n // 2:n // 2 + 9] = v[n // 2]
            for odd in (None, rng):
                data = pp.encode_hybrid(v, bw, odd)
                assert np.array_equal(pp.decode_hybrid(data, bw, n), v)
                got = ck.rle_scan(data, bw, n, first_value=5, payload_base=100)
                want = pp.scan_hybrid(data, bw, n, first_value=5, payload_base=100)
                assert [tuple(int(x) for x in r) for r in got] == want, (bw, n)
                # the table, interpreted the way the kernel does, gives the values back (identity dictionary)
                table = np.concatenate([ck.rle_scan(data, bw, n), np.array([(n, 0, 0, 0)], dtype=capi.RUN_DTYPE)])
                ident = np.arange(min(1 << bw, 1 << 16), dtype=np.int64) if bw <= 16 else None
                if ident is not None:
                    assert np.array_equal(pp.decode_runs(data, table, ident, 8), v.astype(np.int64))


def test_rle_scan_rejects_malformed_streams():
    data = pp.encode_hybrid(np.arange(100, dtype=np.uint64) % 7, 3)
    with pytest.raises(ck.CukingError, match="ends after|beyond"):
        ck.rle_scan(data[: len(data) // 2], 3, 100)
    with pytest.raises(ck.CukingError, match="bit width"):
        ck.rle_scan(data, 33, 100)
    lib = capi.load()
    runs = np.zeros(1, dtype=capi.RUN_DTYPE)
    n = capi.C.c_uint32(0)
    many = pp.encode_hybrid(np.repeat(np.arange(20, dtype=np.uint64), 10), 5)  # 20 RLE runs
    buf = np.frombuffer(many, dtype=np.uint8)
    assert lib.ck_rle_scan(buf.ctypes.data, buf.size, 5, 200, 0, 0, runs.ctypes.data, 1, capi.C.byref(n)) == capi.CK_ERR_OUT_OF_RANGE


def test_oracle_page_reader_pinned_against_pyarrow(tmp_path):
    rng = np.random.default_rng(12)
    g = random_genotypes(rng, 40, 300)
    for name, path in write_variants(tmp_path, g).items():
        pf = pq.ParquetFile(path)
        encodings = set()
        for rg in range(pf.metadata.num_row_groups):
            want = pf.read_row_group(rg)
            for c in range(3):
                pages = pp.read_pages(path, rg, c)
                encodings |= {p["encoding"] for p in pages["pages"]}
                assert np.array_equal(pp.decode_column(pages), want.column(c).to_numpy()), (name, rg, c)
        if name in ("gzip_plain", "dictionary_fallback"):
            assert pp.ENC_PLAIN in encodings, name  # the PLAIN branch really is exercised
        if name == "snappy_v2_pages":
            assert all(p["version"] == 2 for p in pp.read_pages(path, 0, 1)["pages"])


def test_host_reader_windows_equal_libparquet_decode(tmp_path):
    """bin/encoded_check (tools/encoded_check.cc): ReadEncoded's windows, interpreted by a plain loop, against ReadTriples."""
    assert os.path.exists(CHECK), "run __graft_entry__.build() first"
    rng = np.random.default_rng(13)
    g = random_genotypes(rng, 60, 500)
    files = write_variants(tmp_path, g)
    rows = int((g >= 0).sum())
    # whole row groups; windows that straddle pages; windows smaller than a page; windows staged in lent slices - large
    # ones, and ones so small that the reader has to halve its windows (and, below a few thousand rows, use its own arena)
    for window in ("2097152", "4099", "257", "2097152:4194304", "20000:30000", "4099:9000"):
        out = subprocess.run([CHECK, window, *files.values()], capture_output=True, text=True, check=True).stdout.splitlines()
        assert len(out) == len(files)
        for line in out:
            assert line.startswith(f"OK {rows} rows"), (window, line)
        if window == "2097152:4194304":
            assert all("(0 in slices)" not in line for line in out), out
    # nulls are refused with the host path's message (cuking.cu:617-623); an encoding the kernel does not take is reported
    t = triple_table(g).to_pydict()
    t["col_idx"][17] = None
    nulls = str(tmp_path / "nulls.parquet")
    pq.write_table(pa.table(t, schema=triple_table(g).schema), nulls)
    delta = str(tmp_path / "delta.parquet")
    pq.write_table(triple_table(g), delta, use_dictionary=False, column_encoding={"row_idx": "DELTA_BINARY_PACKED", "col_idx": "PLAIN",
                                                                                   "n_alt_alleles": "PLAIN"})
    out = subprocess.run([CHECK, "1000", nulls, delta], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0].startswith("HOST_ERROR Null values in") and out[1].startswith("SAME_ERROR Null values in"), out
    assert out[2].startswith("UNSUPPORTED"), out


# ---- GPU: decode + pack against the oracle's pack -----------------------------------------------------------------
@pytest.mark.gpu
def test_pack_encoded_equals_oracle_pack_on_pyarrow_files(tmp_path):
    rng = np.random.default_rng(14)
    n, s = 70, 900
    g = random_genotypes(rng, n, s)
    want = oracle_bitset(g, ko.submatrix(n))
    with ck.Context(0) as ctx:
        for name, path in write_variants(tmp_path, g).items():
            pf = pq.ParquetFile(path)
            for window in (1 << 20, 1237):
                with ctx.planes(ck.submatrix(n), s) as pl:
                    for rg in range(pf.metadata.num_row_groups):
                        cols = [pp.read_pages(path, rg, c) for c in range(3)]
                        rows = pf.metadata.row_group(rg).num_rows
                        for a in range(0, rows, window):
                            b = min(a + window, rows)
                            pl.pack_encoded([window_column(c, a, b) for c in cols], b - a)
                    assert np.array_equal(pl.export_bitset(), want), (name, window)


@pytest.mark.gpu
def test_pack_encoded_shard_filter_and_odd_streams():
    """Hand-built windows: every bit width, writer oddities (split / empty runs, padded varints), a PLAIN column beside
    dictionary columns, skips, and the Contains() filter of an off-diagonal shard (cuking.cu:677-679)."""
    rng = np.random.default_rng(15)
    n, s = 50, 4000
    g = random_genotypes(rng, n, s)
    site, sample, alt = triples_of(g)
    perm = rng.permutation(len(site))  # any order: the pack is an AND
    site, sample, alt = site[perm], sample[perm], alt[perm]

    def dict_column(values, width, lead):
        uniq, idx = np.unique(values, return_inverse=True)
        order = rng.permutation(len(uniq))  # dictionaries are in first-seen order in real files, i.e. not sorted
        inv = np.empty_like(order)
        inv[order] = np.arange(len(order))
        pad = 1 + int(rng.integers(0, 5))
        d = np.concatenate([uniq[order], rng.integers(0, 1 << 20, pad)]).astype(np.int64 if width == 8 else np.int32)  # unused tail entries
        codes = np.concatenate([rng.integers(0, len(uniq), lead), inv[idx]]).astype(np.uint64)  # `lead` values before the window
        bw = max(int(len(d) - 1).bit_length(), int(rng.integers(0, 3)))
        data = pp.encode_hybrid(codes, bw, rng)
        runs = np.concatenate([ck.rle_scan(data, bw, len(codes)), np.array([(len(codes), 0, 0, 0)], dtype=capi.RUN_DTYPE)])
        return {"bytes": np.frombuffer(data, dtype=np.uint8), "runs": runs, "dict": d, "value_width": width, "skip": lead}

    def plain_column(values, width, lead):
        dt = np.int64 if width == 8 else np.int32
        data = np.concatenate([np.zeros(lead, dtype=dt), values.astype(dt)]).tobytes()
        runs = np.array([(0, capi.CK_RUN_PLAIN, 0, 0), (lead + len(values), 0, 0, 0)], dtype=capi.RUN_DTYPE)
        return {"bytes": np.frombuffer(data, dtype=np.uint8), "runs": runs, "dict": None, "value_width": width, "skip": lead}

    with ck.Context(0) as ctx:
        for k, shard in ((1, 0), (3, 1), (3, 5)):
            sm = ck.submatrix(n, k, shard)
            want = oracle_bitset(g, ko.submatrix(n, k, shard))
            for layout in range(3):
                with ctx.planes(sm, s) as pl:
                    for a in range(0, len(site), 50_000):
                        b = min(a + 50_000, len(site))
                        mk = [dict_column, plain_column] if layout == 1 else [plain_column, dict_column] if layout == 2 else [dict_column, dict_column]
                        cols = [mk[0](site[a:b], 8, int(rng.integers(0, 100))), mk[1](sample[a:b], 8, int(rng.integers(0, 100))),
                                dict_column(alt[a:b], 4, int(rng.integers(0, 100)))]
                        pl.pack_encoded(cols, b - a)
                    assert np.array_equal(pl.export_bitset(), want), (k, shard, layout)


@pytest.mark.gpu
def test_pack_encoded_errors():
    ident = np.arange(8, dtype=np.int64)

    def col(values, d, width=8, bw=3):
        data = pp.encode_hybrid(np.asarray(values, dtype=np.uint64), bw)
        runs = np.concatenate([ck.rle_scan(data, bw, len(values)), np.array([(len(values), 0, 0, 0)], dtype=capi.RUN_DTYPE)])
        return {"bytes": np.frombuffer(data, dtype=np.uint8), "runs": runs, "dict": d, "value_width": width, "skip": 0}

    alt_dict = np.array([0, 1, 2, 7], dtype=np.int32)
    with ck.Context(0) as ctx, ctx.planes(ck.submatrix(8), 8) as pl:
        rows = [1, 2, 3, 4, 5, 6]
        # n_alt_alleles = 7 at row 4 (cuking.cu:698-701): the value and the row are reported
        with pytest.raises(ck.CukingError, match=r"Invalid value for n_alt_alleles \(7\) encountered at triple 4") as e:
            pl.pack_encoded([col(rows, ident), col(rows, ident), col([0, 1, 2, 0, 3, 1], alt_dict, 4)], 6)
        assert e.value.code == capi.CK_ERR_INVALID_GENOTYPE
        # site outside [0, num_sites)
        with pytest.raises(ck.CukingError, match="row_idx out of range") as e:
            pl.pack_encoded([col(rows, ident + 5), col(rows, ident), col([0] * 6, alt_dict, 4)], 6)
        assert e.value.code == capi.CK_ERR_OUT_OF_RANGE
        # a bit-packed index beyond the dictionary: only the device sees it
        with pytest.raises(ck.CukingError, match="dictionary index outside the dictionary at row 5"):
            pl.pack_encoded([col(rows, ident[:6]), col(rows, ident), col([0] * 6, alt_dict, 4)], 6)
        # tables that contradict their buffers are refused before anything runs
        c = col(rows, ident)
        c["runs"] = c["runs"].copy()
        c["runs"]["payload"][0] = 1 << 20
        with pytest.raises(ck.CukingError, match="beyond the payload"):
            pl.pack_encoded([c, col(rows, ident), col([0] * 6, alt_dict, 4)], 6)
        with pytest.raises(ck.CukingError, match="skip \\+ num_rows exceeds"):
            pl.pack_encoded([col(rows, ident), col(rows, ident), col([0] * 6, alt_dict, 4)], 7)
        rle = {"bytes": np.zeros(0, dtype=np.uint8), "runs": np.array([(0, capi.CK_RUN_RLE, 0, 9), (6, 0, 0, 0)], dtype=capi.RUN_DTYPE),
               "dict": ident, "value_width": 8, "skip": 0}
        with pytest.raises(ck.CukingError, match="outside the dictionary of 8 values"):
            pl.pack_encoded([rle, col(rows, ident), col([0] * 6, alt_dict, 4)], 6)
        # and after all those failures the planes still take a clean window (bit width 0: a one-entry dictionary)
        one = {"bytes": np.zeros(0, dtype=np.uint8), "runs": np.array([(0, capi.CK_RUN_RLE, 0, 0), (6, 0, 0, 0)], dtype=capi.RUN_DTYPE),
               "dict": np.array([3], dtype=np.int64), "value_width": 8, "skip": 0}
        pl.reset()
        pl.pack_encoded([one, col(rows, ident), col([0, 1, 2, 0, 1, 2], alt_dict, 4)], 6)
        bs = ko.new_bitset(ko.submatrix(8), 8)
        assert ko.pack(bs, 8, ko.submatrix(8), np.full(6, 3, dtype=np.int64), np.array(rows, dtype=np.int64), np.array([0, 1, 2, 0, 1, 2], dtype=np.int32)) == -1
        assert np.array_equal(pl.export_bitset(), bs)


@pytest.mark.gpu
def test_cli_device_decode_equals_host_decode(tmp_path):
    """bin/cuking on an input directory that mixes every writer variant: the default (pages decoded on the GPU), the
    host-decode path and the oracle agree; a file the device decoder does not take falls back by itself."""
    rng = np.random.default_rng(16)
    n, s = 64, 1200
    g = random_genotypes(rng, n, s)
    d = tmp_path / "in"
    ckio.write_input_dir(str(d), g, num_files=1)
    os.remove(next(p for p in d.iterdir() if p.name.endswith(".parquet")))
    bounds = np.linspace(0, s, len(WRITERS) + 2).astype(int)
    for f, (name, kw) in enumerate(list(WRITERS.items()) + [("delta", dict(use_dictionary=False, column_encoding={
            "row_idx": "DELTA_BINARY_PACKED", "col_idx": "PLAIN", "n_alt_alleles": "PLAIN"}))]):
        lo, hi = int(bounds[f]), int(bounds[f + 1])
        part = np.full_like(g, -1)
        part[:, lo:hi] = g[:, lo:hi]
        pq.write_table(triple_table(part), str(d / f"part-{f:05d}-{name}.parquet"), **kw)
    sm = ko.submatrix(n)
    want, count, _ = ko.king(oracle_bitset(g, sm), s, sm, 0.05, 1 << 20)
    outs = {}
    for mode, env in (("device", {}), ("host", {"CUKING_HOST_DECODE": "1"}), ("small_windows", {"CUKING_DECODE_WINDOW_ROWS": "777"})):
        out = tmp_path / f"out_{mode}"
        p = subprocess.run([os.path.join(ROOT, "bin", "cuking"), f"--input_uri={d}", f"--output_uri={out}", "--kin_threshold=0.05"],
                           capture_output=True, text=True, env={**os.environ, **env})
        assert p.returncode == 0, p.stderr
        assert ("1 file(s) decoded on the host" in p.stdout) == (mode != "host"), p.stdout
        outs[mode] = pq.read_table(out / "part-00000.snappy.parquet")
        assert outs[mode].num_rows == count > 0
        assert np.array_equal(outs[mode].column("kin").to_numpy().view(np.uint32), want["kin"].view(np.uint32))
    assert outs["device"].equals(outs["host"]) and outs["device"].equals(outs["small_windows"])


def test_host_reader_degenerate_files(tmp_path):
    """No rows at all, one row, row groups of one row: the window reader must agree with libparquet on these too."""
    assert os.path.exists(CHECK), "run __graft_entry__.build() first"
    g = np.full((3, 4), -1, dtype=np.int8)
    empty = str(tmp_path / "empty.parquet")
    pq.write_table(triple_table(g), empty)
    g[1, 2] = 2
    one = str(tmp_path / "one.parquet")
    pq.write_table(triple_table(g), one)
    g[:] = 1
    tiny_groups = str(tmp_path / "tiny_groups.parquet")
    pq.write_table(triple_table(g), tiny_groups, row_group_size=1)
    out = subprocess.run([CHECK, "1000:4096", empty, one, tiny_groups], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0].startswith("OK 0 rows") and out[1].startswith("OK 1 rows") and out[2].startswith("OK 12 rows, 12 windows"), out
