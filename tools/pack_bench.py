"""Times the GPU pack kernel alone (the `pack` object of the bench line): tools/pack_bench.py [samples]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import cuking_b200 as ck

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
peak = json.load(open(os.path.join(bench.ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
with ck.Context(0) as ctx:
    print(json.dumps(bench.run_pack_bench(ctx, n, 0.01, peak)))
