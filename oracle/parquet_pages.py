"""CPU restatement of the Parquet page layer the ingest path decodes.  TEST INFRASTRUCTURE ONLY (see king_oracle.py).

What the reference does with its input columns is `ReadBatch` of three Arrow/Parquet column readers
(/root/reference/cuking.cu:603-672; Arrow 8.0.0, /root/reference/Dockerfile:116 - a third-party dependency that is not
vendored in /root/reference).  The algorithm restated here is therefore the published format, Apache Parquet
`Encodings.md` / `parquet.thrift` (format 2.x):

  * page headers: Thrift compact protocol, struct PageHeader (parquet.thrift) - `read_pages`;
  * data page v1 layout: [u32 length + hybrid definition levels, if the column is OPTIONAL][values];
    data page v2: [definition levels, length in the header, never compressed][values];
  * values: PLAIN (little-endian fixed width) or RLE_DICTIONARY / PLAIN_DICTIONARY = [u8 bit width][hybrid indices];
  * the RLE / bit-packing hybrid: runs of <varint header>; header & 1 ? (header >> 1) groups of 8 bit-packed values, LSB
    first : (header >> 1) copies of one value stored in ceil(bit_width / 8) bytes - `decode_hybrid` / `encode_hybrid`.

Pinned by tests/test_pages.py against pyarrow's own reader on files pyarrow writes (several codecs, page versions,
dictionary fallback, OPTIONAL and REQUIRED columns): the columns decoded here equal `pq.read_table`'s.
"""
from __future__ import annotations

import numpy as np

RUN_RLE, RUN_BITPACKED, RUN_PLAIN = 0, 1, 2  # enum ck_run_kind, include/cuking_b200.h
ENC_PLAIN, ENC_PLAIN_DICTIONARY, ENC_RLE, ENC_BIT_PACKED, ENC_RLE_DICTIONARY = 0, 2, 3, 4, 8  # parquet.thrift Encoding
PAGE_DATA, PAGE_INDEX, PAGE_DICTIONARY, PAGE_DATA_V2 = 0, 1, 2, 3  # parquet.thrift PageType


# ---- RLE / bit-packing hybrid ------------------------------------------------------------------------------------
def _varint(data: bytes, pos: int) -> tuple[int, int]:
    shift = value = 0
    while True:
        b = data[pos]
        pos += 1
        value |= (b & 0x7F) << shift
        if not b & 0x80:
            return value, pos
        shift += 7


def _put_varint(out: bytearray, v: int, pad: int = 0) -> None:
    """`pad` > 0 writes a non-minimal (longer) varint of the same value: legal, and readers must cope."""
    groups = []
    while True:
        groups.append(v & 0x7F)
        v >>= 7
        if v == 0:
            break
    groups += [0] * pad
    for g in groups[:-1]:
        out.append(g | 0x80)
    out.append(groups[-1])


def unpack_bits(data: bytes, bit_width: int, count: int) -> np.ndarray:
    """`count` values of `bit_width` bits, LSB first (Encodings.md "bit-packed ... from the LSB of each byte")."""
    if bit_width == 0:
        return np.zeros(count, dtype=np.uint64)
    bits = np.unpackbits(np.frombuffer(data, dtype=np.uint8, count=(count * bit_width + 7) // 8), bitorder="little")
    bits = bits[: count * bit_width].reshape(count, bit_width).astype(np.uint64)
    return (bits << np.arange(bit_width, dtype=np.uint64)).sum(axis=1, dtype=np.uint64)


def pack_bits(values: np.ndarray, bit_width: int) -> bytes:
    if bit_width == 0:
        return b""
    v = np.asarray(values, dtype=np.uint64)
    bits = ((v[:, None] >> np.arange(bit_width, dtype=np.uint64)) & np.uint64(1)).astype(np.uint8)
    return np.packbits(bits.reshape(-1), bitorder="little").tobytes()


def decode_hybrid(data: bytes, bit_width: int, num_values: int) -> np.ndarray:
    """All `num_values` values of one hybrid stream (uint64)."""
    out = np.empty(num_values, dtype=np.uint64)
    pos = done = 0
    vbytes = (bit_width + 7) // 8
    while done < num_values:
        header, pos = _varint(data, pos)
        if header & 1:
            groups = header >> 1
            take = min(groups * 8, num_values - done)
            out[done:done + take] = unpack_bits(data[pos:pos + groups * bit_width], bit_width, take)
            pos += groups * bit_width
        else:
            take = min(header >> 1, num_values - done)
            out[done:done + take] = int.from_bytes(data[pos:pos + vbytes], "little")
            pos += vbytes
        done += take
    return out


def encode_hybrid(values: np.ndarray, bit_width: int, rng: np.random.Generator | None = None) -> bytes:
    """A writer in the style of parquet-cpp / parquet-mr (RLE for >= 8 repeats, bit-packed groups of 8 otherwise, at most
    63 groups per bit-packed run).  With `rng`, legal oddities are mixed in: split runs, short RLE runs, empty runs of both
    kinds, short bit-packed runs, padded varints - structure a reader must not depend on."""
    v = np.asarray(values, dtype=np.uint64)
    n = len(v)
    out = bytearray()
    vbytes = (bit_width + 7) // 8

    def run_len(i: int) -> int:
        r = 1
        while i + r < n and v[i + r] == v[i]:
            r += 1
        return r

    i = 0
    while i < n:
        odd = rng is not None and rng.random() < 0.2
        run = run_len(i)
        if run >= 8 or (odd and rng.random() < 0.5):
            if odd and run > 1:
                run = int(rng.integers(1, run + 1))  # split the run
            if odd and rng.random() < 0.3:
                _put_varint(out, 0)  # an empty RLE run
                out += int(v[i]).to_bytes(vbytes, "little")
            if odd and rng.random() < 0.3:
                _put_varint(out, 1)  # an empty bit-packed run
            _put_varint(out, run << 1, pad=int(rng.integers(0, 3)) if odd and run < (1 << 20) else 0)
            out += int(v[i]).to_bytes(vbytes, "little")
            i += run
            continue
        max_groups = int(rng.integers(1, 8)) if odd else 63
        groups, j = 0, i
        while j < n and groups < max_groups:
            if groups and run_len(j) >= 8:
                break  # a long repeat starts on this group boundary: leave it to an RLE run
            groups += 1
            j += 8
        chunk = v[i:min(j, n)]
        padded = np.zeros(groups * 8, dtype=np.uint64)  # the last group of a stream is padded with zeros
        padded[: len(chunk)] = chunk
        _put_varint(out, (groups << 1) | 1)
        out += pack_bits(padded, bit_width)
        i = min(j, n)
    return bytes(out)


def scan_hybrid(data: bytes, bit_width: int, num_values: int, first_value: int = 0, payload_base: int = 0) -> list[tuple]:
    """The run table ck_rle_scan must produce: (first_value, kind, bit_width, payload) per non-empty run."""
    runs = []
    pos = done = 0
    vbytes = (bit_width + 7) // 8
    while done < num_values:
        header, pos = _varint(data, pos)
        if header & 1:
            take = min((header >> 1) * 8, num_values - done)
            if take:
                runs.append((first_value + done, RUN_BITPACKED, bit_width, payload_base + pos))
            pos += (header >> 1) * bit_width
        else:
            take = min(header >> 1, num_values - done)
            if take:
                runs.append((first_value + done, RUN_RLE, 0, int.from_bytes(data[pos:pos + vbytes], "little")))
            pos += vbytes
        done += take
    return runs


def decode_runs(payload: bytes, runs: np.ndarray, dictionary: np.ndarray | None, value_width: int) -> np.ndarray:
    """Interprets a ck_run table (sentinel included) the way the device kernel does (int64 values)."""
    n = int(runs["first_value"][-1])
    out = np.empty(n, dtype=np.int64)
    dt = np.dtype("<i8") if value_width == 8 else np.dtype("<i4")
    for r in range(len(runs) - 1):
        a, b = int(runs["first_value"][r]), int(runs["first_value"][r + 1])
        kind, bw, pay = int(runs["kind"][r]), int(runs["bit_width"][r]), int(runs["payload"][r])
        if kind == RUN_PLAIN:
            out[a:b] = np.frombuffer(payload, dtype=dt, count=b - a, offset=pay)
        elif kind == RUN_RLE:
            out[a:b] = dictionary[pay]
        else:
            out[a:b] = dictionary[unpack_bits(payload[pay:pay + ((b - a) * bw + 7) // 8], bw, b - a).astype(np.int64)]
    return out


# ---- Thrift compact protocol (just enough for PageHeader) -------------------------------------------------------
def _zigzag(v: int) -> int:
    return (v >> 1) ^ -(v & 1)


def _read_struct(data: bytes, pos: int) -> tuple[dict, int]:
    fields: dict[int, object] = {}
    last = 0
    while True:
        head = data[pos]
        pos += 1
        if head == 0:
            return fields, pos
        delta, ftype = head >> 4, head & 0x0F
        if delta:
            fid = last + delta
        else:
            raw, pos = _varint(data, pos)
            fid = _zigzag(raw)
        last = fid
        fields[fid], pos = _read_value(data, pos, ftype)


def _read_value(data: bytes, pos: int, ftype: int):
    if ftype in (1, 2):  # BOOLEAN_TRUE / BOOLEAN_FALSE live in the type nibble
        return ftype == 1, pos
    if ftype == 3:  # byte
        return data[pos], pos + 1
    if ftype in (4, 5, 6):  # i16 / i32 / i64: zigzag varints
        raw, pos = _varint(data, pos)
        return _zigzag(raw), pos
    if ftype == 7:  # double
        return None, pos + 8
    if ftype == 8:  # binary
        n, pos = _varint(data, pos)
        return data[pos:pos + n], pos + n
    if ftype in (9, 10):  # list / set
        head = data[pos]
        pos += 1
        n, etype = head >> 4, head & 0x0F
        if n == 15:
            n, pos = _varint(data, pos)
        items = []
        for _ in range(n):
            item, pos = _read_value(data, pos, etype)
            items.append(item)
        return items, pos
    if ftype == 12:  # struct
        return _read_struct(data, pos)
    raise ValueError(f"unsupported thrift compact type {ftype}")


# ---- pages of one column chunk -----------------------------------------------------------------------------------
def read_pages(path: str, row_group: int, column: int) -> dict:
    """Walks the pages of one column chunk of a Parquet file: returns {"dictionary": ndarray | None, "pages": [...]} where
    every data page is a dict with `num_values`, `encoding`, `values` (the values section, decompressed), `levels`
    (hybrid definition-level stream or None), `version`.  File metadata (offsets, codec, schema) comes from pyarrow's footer
    reader; everything below the footer is parsed here."""
    import pyarrow as pa
    import pyarrow.parquet as pq

    pf = pq.ParquetFile(path)
    col = pf.metadata.row_group(row_group).column(column)
    descr = pf.schema.column(column)
    width = {"INT64": 8, "INT32": 4}[col.physical_type]
    dtype = np.dtype("<i8") if width == 8 else np.dtype("<i4")
    start = col.dictionary_page_offset if col.has_dictionary_page and col.dictionary_page_offset else col.data_page_offset
    with open(path, "rb") as f:
        f.seek(start)
        chunk = f.read(col.total_compressed_size)
    codec = None if col.compression == "UNCOMPRESSED" else pa.Codec(col.compression.lower())
    out = {"dictionary": None, "pages": [], "value_width": width, "optional": descr.max_definition_level > 0}
    pos = seen = 0
    while seen < col.num_values:
        header, pos = _read_struct(chunk, pos)
        ptype, usize, csize = header[1], header[2], header[3]
        body = chunk[pos:pos + csize]
        pos += csize
        if ptype == PAGE_DICTIONARY:
            raw = body if codec is None else codec.decompress(body, decompressed_size=usize).to_pybytes()
            dph = header[7]
            assert dph[2] in (ENC_PLAIN, ENC_PLAIN_DICTIONARY)
            out["dictionary"] = np.frombuffer(raw, dtype=dtype, count=dph[1]).copy()
        elif ptype == PAGE_DATA:
            raw = body if codec is None else codec.decompress(body, decompressed_size=usize).to_pybytes()
            dph = header[5]
            levels = None
            if descr.max_definition_level > 0:
                assert dph[3] == ENC_RLE
                n = int.from_bytes(raw[:4], "little")
                levels, raw = raw[4:4 + n], raw[4 + n:]
            out["pages"].append({"num_values": dph[1], "encoding": dph[2], "values": raw, "levels": levels, "version": 1})
            seen += dph[1]
        elif ptype == PAGE_DATA_V2:
            dph = header[8]
            dl, rl = dph[5], dph[6]
            levels = body[rl:rl + dl] if descr.max_definition_level > 0 else None
            vals = body[rl + dl:]
            if codec is not None and dph.get(7, True):
                vals = codec.decompress(vals, decompressed_size=usize - rl - dl).to_pybytes()
            out["pages"].append({"num_values": dph[1], "encoding": dph[4], "values": vals, "levels": levels, "version": 2})
            seen += dph[1]
        else:
            continue  # index pages carry no values
    return out


def decode_column(pages: dict) -> np.ndarray:
    """All values of a column chunk (no nulls allowed: every definition level must be 1, cuking.cu:617-623)."""
    parts = []
    for p in pages["pages"]:
        n = p["num_values"]
        if p["levels"] is not None and not np.all(decode_hybrid(p["levels"], 1, n) == 1):
            raise ValueError("Null values")
        if p["encoding"] == ENC_PLAIN:
            dt = np.dtype("<i8") if pages["value_width"] == 8 else np.dtype("<i4")
            parts.append(np.frombuffer(p["values"], dtype=dt, count=n).astype(np.int64))
        elif p["encoding"] in (ENC_RLE_DICTIONARY, ENC_PLAIN_DICTIONARY):
            idx = decode_hybrid(p["values"][1:], p["values"][0], n)
            parts.append(pages["dictionary"][idx.astype(np.int64)].astype(np.int64))
        else:
            raise ValueError(f"unsupported encoding {p['encoding']}")
    return np.concatenate(parts) if parts else np.empty(0, dtype=np.int64)
