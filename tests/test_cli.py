"""The `cuking` binary (cuking_b200/host): flag surface, error behaviour (CPU) and end-to-end parity through real
Parquet files (GPU).  Reference behaviour: /root/reference/cuking.cu:27-52 (flags), :437-462 (validation),
:889-892 (error format), :529-545 (listing), :585-590 (column checks), :868-870 (output name)."""
import json
import os
import subprocess

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import cuking_b200 as ck
from cuking_b200 import io as ckio
from oracle import king_oracle as ko
from tests.helpers import random_genotypes, oracle_bitset

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin", "cuking")


def run(*args):
    return subprocess.run([BIN, *args], capture_output=True, text=True, timeout=600)


def test_binary_exists():
    assert os.path.exists(BIN), "bin/cuking missing - run __graft_entry__.build()"


@pytest.mark.parametrize("args,msg", [
    ([], "INVALID_ARGUMENT: No input URI specified"),
    (["--input_uri=/tmp/x"], "INVALID_ARGUMENT: No output URI specified"),
    (["--input_uri=/tmp/x", "--output_uri=/tmp/y", "--num_reader_threads=0"], "INVALID_ARGUMENT: Invalid number of reader threads"),
    (["--input_uri=/tmp/x", "--output_uri=/tmp/y", "--split_factor=0"], "INVALID_ARGUMENT: Invalid split factor"),
    (["--input_uri=/tmp/x", "--output_uri=/tmp/y", "--split_factor=4", "--shard_index=10"], "INVALID_ARGUMENT: Invalid shard index"),
    (["--input-uri", "/tmp/x", "--output-uri", "/tmp/y", "--split-factor", "2", "--shard-index", "3"], "INVALID_ARGUMENT: Invalid shard index"),
    (["--input_uri=gs://bucket/in", "--output_uri=/tmp/y"], "INVALID_ARGUMENT: Unsupported URI: gs://bucket/in"),
])
def test_flag_validation(args, msg):
    p = run(*args)
    assert p.returncode == 1
    assert p.stderr.startswith("\nError: "), p.stderr  # cuking.cu:890
    assert msg in p.stderr


def test_unknown_flag_and_bad_value():
    p = run("--bogus=1")
    assert p.returncode == 1 and "Unknown command line flag 'bogus'" in p.stderr
    p = run("--max_results=abc")
    assert p.returncode == 1 and "Illegal value 'abc' specified for flag 'max_results'" in p.stderr
    p = run("--kin_threshold")
    assert p.returncode == 1 and "Missing the value for the flag 'kin_threshold'" in p.stderr


def test_metadata_errors(tmp_path):
    p = run(f"--input_uri={tmp_path}", f"--output_uri={tmp_path}/out")
    assert p.returncode == 1 and "FAILED_PRECONDITION: Failed to read metadata" in p.stderr
    (tmp_path / "metadata.json").write_text('{"num_sites": 10, "samples": ["a", ')
    p = run(f"--input_uri={tmp_path}", f"--output_uri={tmp_path}/out")
    assert p.returncode == 1 and "Failed to parse metadata JSON" in p.stderr


def test_io_shims_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    g = random_genotypes(rng, 12, 90)
    info = ckio.write_input_dir(str(tmp_path / "in"), g, num_files=4)
    assert info["num_triples"] == int(np.sum(g >= 0))
    files = sorted(f for f in os.listdir(tmp_path / "in") if f.endswith(".parquet"))
    assert len(files) == 4
    t = pa.concat_tables([pq.read_table(tmp_path / "in" / f) for f in files])
    assert t.schema.names == ["row_idx", "col_idx", "n_alt_alleles"]
    back = np.full_like(g, -1)
    back[t["col_idx"].to_numpy(), t["row_idx"].to_numpy()] = t["n_alt_alleles"].to_numpy()
    assert np.array_equal(back, g)
    md = json.load(open(tmp_path / "in" / "metadata.json"))
    assert md["num_sites"] == 90 and len(md["samples"]) == 12


# ---- GPU: end to end ---------------------------------------------------------------------------------------------


def expected_records(g, k, shard, thr, cap=1 << 20):
    sm = ko.submatrix(g.shape[0], k, shard)
    res, count, ovf = ko.king(oracle_bitset(g, sm), g.shape[1], sm, thr, cap)
    assert not ovf
    return res


def check_output(table, want, ids):
    assert table.num_rows == len(want)
    assert table["i"].to_pylist() == [ids[x] for x in want["sample_i"]]
    assert table["j"].to_pylist() == [ids[x] for x in want["sample_j"]]
    assert np.array_equal(table["kin"].to_numpy().view(np.uint32), want["kin"].view(np.uint32))
    for f in ("ibs0", "ibs1", "ibs2"):
        assert np.array_equal(table[f].to_numpy().astype(np.uint32), want[f])


@pytest.mark.gpu
def test_cfg1_end_to_end_through_parquet(tmp_path):
    # BASELINE.json configs[0]: 1,000 samples x 10,000 sites, threshold 0.05, real Parquet in and out
    g = ck.synth_genotypes_host(42, 0.02, 0, 1000, 0, 10_000)
    ids = [f"NA{idx:05d}" for idx in range(1000)]
    ckio.write_input_dir(str(tmp_path / "in"), g, ids, num_files=16)
    p = run(f"--input_uri={tmp_path}/in", f"--output_uri=file://{tmp_path}/out", "--kin_threshold=0.05")
    assert p.returncode == 0, p.stderr
    assert "Found 16 input files." in p.stdout and "Running KING CUDA kernel for 1000 x 1000 matrix" in p.stdout
    assert os.listdir(tmp_path / "out") == ["part-00000.snappy.parquet"]
    want = expected_records(g, 1, 0, 0.05)
    assert len(want) > 100
    check_output(ckio.read_output_dir(str(tmp_path / "out")), want, ids)


@pytest.mark.gpu
def test_shards_and_hyphen_flags(tmp_path):
    rng = np.random.default_rng(4)
    g = random_genotypes(rng, 203, 1500)
    ids = [f"s{idx}" for idx in range(203)]
    ckio.write_input_dir(str(tmp_path / "in"), g, ids, num_files=5, nullable=False, row_group_size=1000)
    k = 3
    for shard in range(6):
        p = run("--input-uri", f"{tmp_path}/in", "--output-uri", f"{tmp_path}/out", "--kin-threshold", "0.04",
                "--split-factor", str(k), "--shard-index", str(shard), "--num_reader_threads=3")
        assert p.returncode == 0, p.stderr
        want = expected_records(g, k, shard, 0.04)
        check_output(pq.read_table(tmp_path / "out" / f"part-{shard:05d}.snappy.parquet"), want, ids)
    full = expected_records(g, 1, 0, 0.04)
    merged = ckio.read_output_dir(str(tmp_path / "out"))
    assert merged.num_rows == len(full)
    # --all_shards decodes once and writes the same six part files
    p = run(f"--input_uri={tmp_path}/in", f"--output_uri={tmp_path}/out_all", "--kin_threshold=0.04", "--split_factor=3",
            "--all_shards")
    assert p.returncode == 0, p.stderr
    for shard in range(6):
        a = pq.read_table(tmp_path / "out" / f"part-{shard:05d}.snappy.parquet")
        b = pq.read_table(tmp_path / "out_all" / f"part-{shard:05d}.snappy.parquet")
        assert a.equals(b)


def _gpu_count():
    import ctypes as C

    from cuking_b200 import capi

    n = C.c_int(0)
    try:
        capi.load().ck_device_count(C.byref(n))
    except Exception:
        return 0
    return n.value


@pytest.mark.gpu
def test_all_shards_success_marker_dense_output_and_row_groups(tmp_path):
    # the orchestrator's contract (cloud_batch_submit.py:73, :103-127): one part file per shard, then _SUCCESS;
    # --kin_threshold -1 takes the dense output path, streamed into the Parquet writer chunk by chunk
    rng = np.random.default_rng(8)
    n = 1300
    g = random_genotypes(rng, n, 400)
    ids = [f"s{idx}" for idx in range(n)]
    ckio.write_input_dir(str(tmp_path / "in"), g, ids, num_files=4)
    p = run(f"--input_uri={tmp_path}/in", f"--output_uri={tmp_path}/out", "--kin_threshold=-1", "--split_factor=2",
            "--all_shards", "--write_success_file", "--row_group_rows=100000")
    assert p.returncode == 0, p.stderr
    assert sorted(os.listdir(tmp_path / "out")) == ["_SUCCESS"] + [f"part-{q:05d}.snappy.parquet" for q in range(3)]
    assert os.path.getsize(tmp_path / "out" / "_SUCCESS") == 0
    for shard in range(3):
        f = pq.ParquetFile(tmp_path / "out" / f"part-{shard:05d}.snappy.parquet")
        want = expected_records(g, 2, shard, -1.0)
        assert f.metadata.num_row_groups == -(-len(want) // 100000)
        check_output(f.read(), want, ids)
    # without the flag no marker is written (reference default: --write-success-file is opt-in)
    p = run(f"--input_uri={tmp_path}/in", f"--output_uri={tmp_path}/out2", "--kin_threshold=0.1", "--split_factor=2", "--all_shards")
    assert p.returncode == 0 and not os.path.exists(tmp_path / "out2" / "_SUCCESS")


@pytest.mark.gpu
@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("gpus", [2, 8])
def test_multi_gpu_cli_equals_single_gpu(tmp_path, gpus):
    # --num_gpus N: triples dealt to the GPUs and packed once, planes AND-reduced over NVLink, shards (or the parts of a
    # lone shard) scheduled across the GPUs; every part file must equal the single-GPU one byte for byte in content
    if _gpu_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    rng = np.random.default_rng(9)
    n = 2700
    g = random_genotypes(rng, n, 600)
    ids = [f"s{idx}" for idx in range(n)]
    ckio.write_input_dir(str(tmp_path / "in"), g, ids, num_files=7, row_group_size=50_000)
    common = [f"--input_uri={tmp_path}/in", "--kin_threshold=0.04", "--num_reader_threads=4"]
    for extra, shards in ([], [0]), (["--split_factor=3", "--shard_index=4"], [4]), (["--split_factor=3", "--all_shards"], range(6)):
        p1 = run(*common, *extra, f"--output_uri={tmp_path}/one")
        assert p1.returncode == 0, p1.stderr
        pn = run(*common, *extra, f"--output_uri={tmp_path}/many", f"--num_gpus={gpus}")
        assert pn.returncode == 0, pn.stderr
        assert "Exchanging bit sets" in pn.stdout
        for shard in shards:
            a = pq.read_table(tmp_path / "one" / f"part-{shard:05d}.snappy.parquet")
            b = pq.read_table(tmp_path / "many" / f"part-{shard:05d}.snappy.parquet")
            assert a.num_rows > 0 and a.equals(b), (extra, shard)
        k = 3 if extra else 1
        for shard in shards:
            check_output(pq.read_table(tmp_path / "many" / f"part-{shard:05d}.snappy.parquet"), expected_records(g, k, shard, 0.04), ids)
    # dense output split into parts and merged on the host
    pn = run(f"--input_uri={tmp_path}/in", "--kin_threshold=-1", f"--output_uri={tmp_path}/dense", f"--num_gpus={gpus}")
    assert pn.returncode == 0, pn.stderr
    check_output(pq.read_table(tmp_path / "dense" / "part-00000.snappy.parquet"), expected_records(g, 1, 0, -1.0, cap=1 << 22), ids)


@pytest.mark.gpu
def test_input_errors(tmp_path):
    rng = np.random.default_rng(5)
    g = random_genotypes(rng, 20, 100)
    d = tmp_path / "in"
    ckio.write_input_dir(str(d), g, num_files=2)
    # invalid genotype value (cuking.cu:698-701)
    bad = pa.table({"row_idx": pa.array([1], pa.int64()), "col_idx": pa.array([2], pa.int64()),
                    "n_alt_alleles": pa.array([3], pa.int32())})
    pq.write_table(bad, d / "zz-bad.parquet")
    p = run(f"--input_uri={d}", f"--output_uri={tmp_path}/out")
    assert p.returncode == 1 and "FAILED_PRECONDITION: Invalid value for n_alt_alleles (3)" in p.stderr
    os.remove(d / "zz-bad.parquet")
    # wrong physical type (cuking.cu:608-612)
    bad = pa.table({"row_idx": pa.array([1], pa.int32()), "col_idx": pa.array([2], pa.int64()),
                    "n_alt_alleles": pa.array([1], pa.int32())})
    pq.write_table(bad, d / "zz-bad.parquet")
    p = run(f"--input_uri={d}", f"--output_uri={tmp_path}/out")
    assert p.returncode == 1 and "Expected INT64 type, found INT32" in p.stderr
    os.remove(d / "zz-bad.parquet")
    # wrong column count (cuking.cu:585-590)
    pq.write_table(pa.table({"a": pa.array([1], pa.int64())}), d / "zz-bad.parquet")
    p = run(f"--input_uri={d}", f"--output_uri={tmp_path}/out")
    assert p.returncode == 1 and "Expected 3 columns, found 1" in p.stderr
    os.remove(d / "zz-bad.parquet")
    # result overflow (cuking.cu:747-751): nothing is written
    p = run(f"--input_uri={d}", f"--output_uri={tmp_path}/out", "--kin_threshold=-1", "--max_results=5")
    assert p.returncode == 1
    assert "RESOURCE_EXHAUSTED: Could not store all results: try increasing the --max_results parameter." in p.stderr
    assert not os.path.exists(tmp_path / "out" / "part-00000.snappy.parquet")
    # no parquet files at all (cuking.cu:542-544)
    e = tmp_path / "empty"
    e.mkdir()
    (e / "metadata.json").write_text(json.dumps({"num_sites": 5, "samples": ["a", "b"]}))
    p = run(f"--input_uri={e}", f"--output_uri={tmp_path}/out")
    assert p.returncode == 1 and "FAILED_PRECONDITION: No input files found" in p.stderr
