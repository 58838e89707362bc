#!/usr/bin/env python3
"""bench.py — sample-pair·sites / second of the pairwise KING hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU, NCCL only for barrier/max-reduce)

A *step* is one full pass of the hot path over the workload: every i<j sample pair evaluated at every site,
kinship, threshold, compaction, device sort and copy-out of the retained pairs.  Workload at N = 1 is BASELINE.json
configs[1]: 100,000 samples x 100,000 sites (synthetic cohort of SURVEY.md §8d, seed 42, 1 % missing, threshold
0.0884).  For N > 1 the cohort grows to 100,000*sqrt(N) samples so that every GPU keeps the N = 1 amount of work
(weak scaling); every rank holds all planes and takes a contiguous slice of the tile grid — no data-path collective.

`value`   : planes already resident in HBM when the timed region starts (device-timed, max over ranks).
`e2e`     : the same pass through the host-buffer C-ABI call ck_king_host_bitset: the reference-layout bit set starts
            in pinned HOST memory, H2D (overlapped with the kernel, last sample chunk first) + layout transpose + code
            derivation + kernel + sort + D2H all inside the timed region.
`roofline`: the pairwise kernel against the tensor throughput it is bound by: 10 fp4 ops (5 exact E2M1 MACs) per
            pair·site against the dense kind::mxf4 rate measured on this pool's B200 (tools/umma_mxf4_probe.cu); the
            SURVEY.md §8d view (0.1875 POPC.32 lane-ops per pair·site against the POPC issue rate measured live on
            this GPU) is reported beside it as `popc_equivalent`, and is the roofline of --variant 0/1.
`cpu_baseline` / --impl reference: the oracle's OpenMP restatement of the reference loop on this box's host cores, on
            a bounded sample of the same workload (the reference has no CPU implementation; kind = "port").
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "sample-pair·sites/sec"
UNIT = "pair·sites/s"
SEED = 42
ALGO_POPC_PER_UNIT = 12.0 / 64.0  # reference formulation: 6 __popcll per 64 sites = 12 POPC.32 (cuking.cu:232-239)

WORKLOADS = {
    # name: (samples at N=1, sites, missing, threshold)
    "cfg2": (100_000, 100_000, 0.01, 0.0884),   # BASELINE.json configs[1]
    "cfg1": (1_000, 10_000, 0.02, 0.05),        # configs[0] (parity-test case; selectable for quick runs)
    "mid": (20_000, 100_000, 0.01, 0.0884),     # quick smoke of the bench itself
    "prof": (8_192, 100_000, 0.01, 0.0884),     # short kernel for ncu captures (profiles/)
    "cfg5": (50_000, 1_000_000, 0.01, -1.0),    # BASELINE.json configs[4]: dense output, every finite-kin pair is emitted
    "cfg4": (1_000_000, 100_000, 0.05, 0.0442), # BASELINE.json configs[3]: fixed cohort at every N (strong scaling): its
                                                # ms_per_step is the metric's "wall-time for 1M x 100k"
}
FIXED_SIZE = {"cfg4", "cfg5"}  # BASELINE.json quotes these on a fixed cohort spread over the 8 GPUs


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="cfg2")
    ap.add_argument("--variant", type=int, default=-1, help="pairwise kernel variant (-1 = library default)")
    ap.add_argument("--e2e-steps", type=int, default=-1, help="steps of the host-buffer leg (-1 = same as --steps, 0 = skip)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--clock-sample-ms", type=float, default=200.0, help="NVML sampling period during the timed region (0 = off)")
    ap.add_argument("--e2e-mode", choices=["allgather", "host"], default="allgather",
                    help="N > 1 host-buffer leg: planes replicated by NCCL all-gather (default) or N full host uploads")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-ref-gpu", action="store_true")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def units(n_samples: int, n_sites: int) -> float:
    return n_samples * (n_samples - 1) / 2.0 * n_sites


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU every 200 ms through NVML (the nvidia-smi fields of
    B200_PROFILING.md) while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.2):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self.period_s = period_s
        self._stop_evt = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None or self.period_s <= 0:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period_s)

    def stop(self) -> dict:
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---- CPU baseline (oracle port, bounded sample) ------------------------------------------------------------------


def cpu_sample_bitset(n_sample: int, n_sites: int, missing: float):
    """Reference-layout bit set of the first n_sample samples of the workload cohort.  Built by the product's GPU
    generator when a GPU is present (identical data to the timed workload), else by the host generator."""
    from oracle import king_oracle as ko
    import cuking_b200 as ck

    try:
        with ck.Context(int(os.environ.get("LOCAL_RANK", "0"))) as ctx, ctx.planes(ck.submatrix(n_sample), n_sites) as pl:
            pl.synthesize(SEED, missing)
            return pl.export_bitset()
    except ck.CukingError:
        g = ck.synth_genotypes_host(SEED, missing, 0, n_sample, 0, n_sites)
        bs, _ = ko.pack_dense(g)
        return bs


def run_cpu_baseline(n_sites: int, missing: float, thr: float, target_s: float, steps: int = 1, warmup: int = 0):
    """Times the oracle (OpenMP, all host threads) on an r x r off-diagonal rectangle of the workload sized for
    ~target_s seconds; returns (pair·sites/s, cores, sample description, per-step seconds)."""
    from oracle import king_oracle as ko

    L = ko.lib(native=True)
    cores = int(L.ko_num_threads())
    wps = ko.words_per_sample(n_sites)
    cal = 256
    bs = cpu_sample_bitset(2 * cal, n_sites, missing)
    t0 = time.perf_counter()
    L.ko_bench_rect(bs.ctypes.data, wps, 0, cal, cal, cal, thr)
    t_cal = max(time.perf_counter() - t0, 1e-4)
    rate = cal * cal * n_sites / t_cal
    r = int(min(8192, max(cal, (target_s * rate / n_sites) ** 0.5)))
    r = max(64, (r // 64) * 64)
    if r != cal:
        bs = cpu_sample_bitset(2 * r, n_sites, missing)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        L.ko_bench_rect(bs.ctypes.data, wps, 0, r, r, r, thr)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    value = r * r * n_sites / (sum(times) / len(times))
    sample = f"{r}x{r}-sample rectangle (rows 0..{r - 1} x cols {r}..{2 * r - 1}) x {n_sites} sites of the workload cohort"
    return value, cores, sample, times


def run_reference_impl(args, rank: int):
    """--impl reference: the reference's path on the host CPU.  The reference has no CPU implementation (it is one
    CUDA kernel), so this is the oracle's OpenMP port of cuking.cu:216-307 with all host threads."""
    if rank != 0:
        return
    n, s, missing, thr = WORKLOADS[args.workload]
    per_step = max(2.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
    value, cores, sample, times = run_cpu_baseline(s, missing, thr, per_step, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 popcount + fp32 kinship", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {n} samples x {s} sites, missing {missing}, kin_threshold {thr}",
                   "note": "CPU port of the reference loop (the reference itself is GPU-only); each step is a bounded sample"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- reference GPU kernel on the same B200 (bounded sample) --------------------------------------------------------


def run_reference_gpu_kernel(ctx, n_sites: int, missing: float, thr: float):
    """The reference's own ComputeKingKernel (oracle/_ref, cuking.cu:191-314 for sm_100a) on an off-diagonal shard of
    the workload cohort, device-resident and as shipped (managed memory, host first touch)."""
    from oracle import ref_kernel
    import cuking_b200 as ck

    if not ref_kernel.available():
        return {"unavailable": "oracle/_ref/libcuking_ref.so not built (needs /root/reference at build time)"}
    out = {}
    wps = ck.words_per_sample(n_sites)
    for mode, r in (("device_resident", 4096), ("managed_as_shipped", 2048)):
        n = 2 * r
        sm = ck.submatrix(n, 2, 1)  # rows [0, r) x cols [r, 2r): r*r pairs
        with ctx.planes(sm, n_sites) as pl:
            pl.synthesize(SEED, missing)
            bs = pl.export_bitset()
        best = None
        for _ in range(2):
            _, count, ovf, ms = ref_kernel.king(bs, n, 2, 1, wps, thr, 1 << 20, managed=(mode == "managed_as_shipped"))
            best = ms if best is None else min(best, ms)
        out[mode] = {"value": r * r * n_sites / (best * 1e-3), "unit": UNIT, "kernel_ms": best,
                     "sample": f"{r}x{r} off-diagonal shard x {n_sites} sites", "retained_pairs": count}
    return out


# ---- pack stage (HBM-bound): triples -> bit planes on the GPU --------------------------------------------------------


def run_pack_bench(ctx, n_samples: int, missing: float, hbm_peak_gbs):
    """Times the GPU pack kernel (cuking.cu:675-703 moved to the device) on a slab of the workload's triples generated on
    the device in Hail's order: all samples x 2048 sites.  Algorithmic traffic: 20 B read per triple (int64, int64, int32)."""
    import cuking_b200 as ck

    sites = 2048
    r, c, a, n = ctx.synth_triples_device(SEED, missing, 0, n_samples, 0, sites)
    best = None
    with ctx.planes(ck.submatrix(n_samples), sites) as pl:
        for _ in range(4):
            pl.pack_device_ptrs(r, c, a, n)
            ms = ctx.timings()["pack_ms"]
            best = ms if best is None else min(best, ms)
    gbs = n * 20.0 / (best * 1e-3) / 1e9
    return {"kernel": "pack_kernel", "triples": int(n), "ms": best, "triples_per_s": n / (best * 1e-3), "achieved_gbs": gbs,
            "peak_gbs": hbm_peak_gbs, "frac": (gbs / hbm_peak_gbs) if hbm_peak_gbs else None, "bound": "hbm",
            "algorithmic_per_unit": "20 B read per triple (row_idx int64, col_idx int64, n_alt_alleles int32); plane writes are 2 bits per genotype",
            "sample": f"{n_samples} samples x {sites} sites of the workload cohort, Hail order"}


# ---- our arm ------------------------------------------------------------------------------------------------------


def main():
    args = parse_args()
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        run_reference_impl(args, rank)
        return

    import torch
    import cuking_b200 as ck

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback (use --impl reference for the CPU arm)")
    n_gpus = args.gpus
    distributed = world > 1
    if distributed:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        # NCCL kernels of the plane replication run beside the pairwise kernel, which fills every SM: give them priority
        try:
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=opts)
        except (AttributeError, TypeError):  # older torch: default-priority NCCL streams still work, only slower to get SMs
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        assert world == n_gpus, f"--gpus {n_gpus} but WORLD_SIZE={world}"
    elif n_gpus != 1:
        raise SystemExit("for --gpus N > 1 launch with torchrun (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    n1, n_sites, missing, thr = WORKLOADS[args.workload]
    fixed = args.workload in FIXED_SIZE
    n_samples = int(round(n1 * (n_gpus ** 0.5) / 64.0)) * 64 if (n_gpus > 1 and not fixed) else n1  # weak scaling: pairs ~ N
    max_results = 10 << 20  # the reference's default --max_results (cuking.cu:40)
    if thr < 0:  # dense-output stress: room for every pair of this rank's slice
        max_results = int(n_samples * (n_samples - 1) // 2 // max(1, n_gpus) * 1.02) + 1024

    # a dedicated stream: the library launches on it and the timing events are recorded on it
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx = ck.Context(local_rank, stream=stream.cuda_stream)
    if args.variant >= 0:
        ctx.set_king_variant(args.variant)
    peaks = ctx.measure_int_peaks()

    # ---- inputs resident in HBM (outside the timed region) ----
    sm = ck.submatrix(n_samples)
    planes = ctx.planes(sm, n_sites)
    t0 = time.perf_counter()
    planes.synthesize(SEED, missing)
    planes.finalize()
    synth_s = time.perf_counter() - t0
    tiles = planes.num_tiles()
    t_begin, t_end = tiles * rank // world, tiles * (rank + 1) // world
    results = np.empty(max_results, dtype=ck.RESULT_DTYPE)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step():
        return planes.king(thr, max_results, sort=True, tiles=(t_begin, t_end), out=results)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank, args.clock_sample_ms / 1e3)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    king_ms, launches, retained = [], 0, 0
    ev0.record(stream)
    for _ in range(args.steps):
        res = step()
        t = ctx.timings()
        king_ms.append(t["king_ms"])
        launches += int(t["king_launches"])
        retained = len(res)
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    if distributed:
        tmax = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tmax.item())
        tot = torch.tensor([float(retained)], device=dev, dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        retained = int(tot.item())
    total_units = units(n_samples, n_sites)
    ms_per_step = elapsed_ms / args.steps
    value = total_units / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (this rank's slice, device events around the launch on its stream) ----
    my_units = total_units * (t_end - t_begin) / max(1, tiles)
    kernel_ms = float(np.mean(king_ms))
    achieved = my_units * ALGO_POPC_PER_UNIT / (kernel_ms * 1e-3)
    hbm_peak = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak = json.load(f).get("hbm_gbs")
    except OSError:
        pass
    # algorithmic HBM bytes per launch: each tile streams its row samples and its column samples once
    # (tcgen05 variant: 128 x 80 tiles of 4-bit genotype codes; LOP3+POPC variants: 64 x 64 tiles of 3 bit planes)
    words = -(-(-(-n_sites // 32)) // 16) * 16
    variant = args.variant if args.variant >= 0 else (3 if n_sites <= (1 << 23) else 2)  # the library's own choice
    umma = variant in (2, 3)
    tile_bytes = (128 + 80) * words * 16 if umma else 2 * 64 * words * 12
    popc_view = {
        "achieved": achieved / 1e9, "peak": peaks["popc_lane_ops_per_s"] / 1e9, "unit": "G POPC.32 lane-ops/s",
        "frac": achieved / peaks["popc_lane_ops_per_s"],
        "algorithmic_per_unit": "0.1875 POPC.32 lane-ops per pair-site (reference formulation: 6 popcounts per site-bit, cuking.cu:232-239)",
        "peak_source": "POPC.32 issue rate measured live on this GPU by ck_measure_int_peaks (16 lanes/clk/SM)",
        "lop3_peak": peaks["lop3_lane_ops_per_s"] / 1e9,
    }
    hbm_view = {"algorithmic_gbs": (t_end - t_begin) * tile_bytes / (kernel_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                "note": "tile operand streaming (each tile reads its row and column samples once); mostly L2 hits"}
    bf16 = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            bf16 = json.load(f)
    except OSError:
        pass
    # DRAM bytes of one king_fp4_kernel launch on the single-GPU cfg2 shape: dram__bytes_read.sum + dram__bytes_write.sum
    # of the ncu --set full capture in profiles/r01_king_fp4_cfg2_ncu.txt (425.06 GB + 0.16 GB; 7 % of the DRAM peak - the
    # codes are re-read from L2, hit rate 77 %; compulsory traffic is 5 GB).  Other shapes were not captured: null.
    fp4_traffic = 425.21e9 if (args.workload == "cfg2" and n_gpus == 1) else None
    if variant == 3:
        # 5 exact E2M1 MACs = 10 ops per pair-site (xx, yy, yh, hy, hh) on the FP4 tensor path (kind::mxf4, unit block
        # scales, fp32 accumulation).  Peak = the dense FP4 rate measured on this pool's B200 by tools/umma_mxf4_probe.cu
        # (15,595 MAC/clk/SM, 8481 TOP/s at the burst clock; nominal 9000) - MEASURED_PEAKS.json only holds bf16
        # (FP4 is nominally 4x bf16: 4 x 1657 burst = 6626, 4 x 1390 sustained = 5562 TOP/s).
        tops = my_units * 10.0 / (kernel_ms * 1e-3) / 1e12
        roofline = {
            "bound": "tensor", "kernel": "king_fp4_kernel", "achieved": tops, "peak": 8481.0, "unit": "TOP/s (fp4 e2m1, dense)",
            "frac": tops / 8481.0, "traffic": fp4_traffic, "kernel_ms": kernel_ms, "units_per_launch": my_units,
            "algorithmic_per_unit": "10 fp4 ops (5 MACs: xx, yy, yh, hy, hh) per pair-site, fp32 accumulation (exact: counts <= 2^23)",
            "peak_source": "measured: tools/umma_mxf4_probe.cu on this pool's B200 (profiles/r01_mxf4_probe.txt), "
                           "kind::mxf4 M=128 N=208, 15595 MAC/clk/SM = 8481 TOP/s at the burst clock (nominal dense fp4: 9000)",
            "vs_4x_measured_bf16_burst": (tops / (4 * bf16["bf16_tflops"])) if bf16 else None,
            "vs_4x_measured_bf16_sustained": (tops / (4 * bf16["bf16_tflops_sustained"])) if bf16 else None,
            "int8_equivalent": {"peak": 4075.0, "frac": tops / 4075.0,
                                "note": "the same 10 ops per pair-site against the measured int8 tensor peak (the round's first tensor kernel, variant 2)"},
            "popc_equivalent": popc_view, "hbm": hbm_view,
        }
    elif umma:
        # 5 exact int8 MACs = 10 ops per pair-site (xx, yy, yh, hy, hh); peak = dense int8 tcgen05 rate measured on this
        # pool's B200 by tools/umma_i8_probe.cu (8190 MAC/clk/SM, 4075 TOP/s burst at N=256) - MEASURED_PEAKS.json only
        # holds bf16 (int8 is nominally 2x bf16: 2 x 1390 sustained = 2781, 2 x 1657 burst = 3313 TOP/s)
        tops = my_units * 10.0 / (kernel_ms * 1e-3) / 1e12
        roofline = {
            "bound": "tensor", "kernel": "king_umma_kernel", "achieved": tops, "peak": 4075.0, "unit": "TOP/s (int8, dense)",
            "frac": tops / 4075.0, "traffic": None, "kernel_ms": kernel_ms, "units_per_launch": my_units,
            "algorithmic_per_unit": "10 int8 ops (5 MACs: xx, yy, yh, hy, hh) per pair-site, s32 accumulation",
            "peak_source": "measured: tools/umma_i8_probe.cu on this pool's B200 (profiles/r01_umma_probe.txt), "
                           "kind::i8 M=128 N=256, 8190 MAC/clk/SM = 4075 TOP/s at the burst clock",
            "vs_2x_measured_bf16_sustained": (tops / (2 * bf16["bf16_tflops_sustained"])) if bf16 else None,
            "popc_equivalent": popc_view, "hbm": hbm_view,
        }
    else:
        roofline = dict(popc_view, bound="popc", kernel="king_tile_kernel", traffic=None, kernel_ms=kernel_ms,
                        units_per_launch=my_units, hbm=hbm_view)

    # ---- e2e: host buffers through the reference-facing C-ABI call (rank-local slice is the whole shard at N=1) ----
    e2e, exchange = None, None
    e2e_steps = args.steps if args.e2e_steps < 0 else args.e2e_steps
    if e2e_steps > 0:
        bits_np = planes.export_bitset()  # reference layout (cuking.cu:507-523), built once outside the timed region
        host_bits = torch.from_numpy(bits_np).pin_memory()
        del bits_np
        h2d = host_bits.numel() * 8
        # N = 1: ck_king_host_bitset (upload overlapped with the kernel).  N > 1: the same schedule with the planes
        # replicated over NVLink (every rank uploads 1/N of each chunk through its own PCIe link, NCCL all-gather,
        # ck_king_stream_rows) - or, with --e2e-mode host, N independent full uploads through ck_king_host_bitset_part.
        allgather = world > 1 and args.e2e_mode == "allgather"
        if allgather:
            from cuking_b200.distributed import king_host_bitset_allgather
            side = torch.cuda.Stream(dev, priority=-1)

            def e2e_step():
                with ctx.planes(sm, n_sites) as pl:
                    return king_host_bitset_allgather(pl, host_bits, ck.words_per_sample(n_sites), thr, max_results,
                                                      out=results, side_stream=side)
        else:
            def e2e_step():
                return ctx.king_host_bitset(n_samples, 1, 0, n_sites, host_bits, thr, max_results, out=results, part=(rank, world))
        e2e_step()  # warm-up (allocations)
        barrier()
        ev0.record(stream)
        for _ in range(e2e_steps):
            r = e2e_step()
        ev1.record(stream)
        barrier()
        e_ms = ev0.elapsed_time(ev1)
        e2e_retained = len(r)
        if distributed:
            tmax = torch.tensor([e_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            e_ms = float(tmax.item())
            tot = torch.tensor([float(e2e_retained)], device=dev, dtype=torch.float64)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            e2e_retained = int(tot.item())
        if e2e_retained != retained:  # the host-buffer leg must keep exactly the pairs the resident leg kept
            raise SystemExit(f"e2e leg retained {e2e_retained} pairs, resident leg {retained}")
        e2e = {"value": total_units / (e_ms / e2e_steps * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d // world) if allgather else int(h2d),  # per rank
               "d2h_bytes_per_step": int(len(r) * 24 + 8), "steps": e2e_steps, "ms_per_step": e_ms / e2e_steps,
               "api": ("ck_king_host_bitset" if world == 1 else
                       f"ck_king_stream_* fed by 1/{world} uploads + NCCL all-gather (cuking_b200.distributed.king_host_bitset_allgather)"
                       if allgather else f"ck_king_host_bitset_part (part r of {world} on GPU r)")
                      + " (pinned host bit set in the reference layout -> sorted KingResult[] on the host)"}
        # north_star: replicate the planes with an NCCL broadcast over NVLink "only if it beats per-GPU host loading".
        # Measured here on the e2e call's own bytes: (a) every rank copies the whole pinned bit set to its GPU at the
        # same time (what ck_king_host_bitset_part does, hidden behind its kernel), (b) rank 0 alone copies it,
        # (c) rank 0 broadcasts the device copy to the other GPUs.  Device-timed, max over ranks.
        if distributed:
            hb = host_bits.view(torch.int64)
            dbuf = torch.empty_like(hb, device=dev)

            def timed(fn, reps=2):
                fn()  # warm-up (NCCL channel setup, page-locking checks)
                barrier()
                ev0.record(stream)
                for _ in range(reps):
                    fn()
                ev1.record(stream)
                barrier()
                t = torch.tensor([ev0.elapsed_time(ev1) / reps], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return float(t.item())

            all_ms = timed(lambda: dbuf.copy_(hb, non_blocking=True))
            one_ms = timed(lambda: dbuf.copy_(hb, non_blocking=True) if rank == 0 else None)
            bcast_ms = timed(lambda: dist.broadcast(dbuf, src=0))
            gb = h2d / 1e9
            exchange = {
                "bytes": int(h2d), "per_gpu_host_load_ms": all_ms, "per_gpu_host_load_gbs_each": gb / (all_ms * 1e-3),
                "single_gpu_host_load_ms": one_ms, "nccl_broadcast_ms": bcast_ms, "nccl_broadcast_gbs": gb / (bcast_ms * 1e-3),
                "load_once_then_broadcast_ms": one_ms + bcast_ms,
                "decision": "at N = 2 concurrent per-GPU host loading matches a single load; at N = 8 it is host-limited (about 23 GB/s "
                            "per GPU) and replication over NVLink wins, so the e2e leg uploads 1/N of every chunk per GPU and "
                            "all-gathers it (king_host_bitset_allgather); both are overlapped with the pairwise kernel",
            }
            del dbuf, hb
        del host_bits

    cpu_baseline, ref_gpu, pack = None, None, None
    if rank == 0 and n_gpus == 1:
        try:
            pack = run_pack_bench(ctx, n_samples, missing, hbm_peak)
        except Exception as exc:
            pack = {"error": repr(exc)}
        if not args.skip_ref_gpu:
            try:
                ref_gpu = run_reference_gpu_kernel(ctx, n_sites, missing, thr)
            except Exception as exc:  # the baseline must never take the bench down
                ref_gpu = {"error": repr(exc)}
        if not args.skip_cpu:
            v, cores, sample, _ = run_cpu_baseline(n_sites, missing, thr, args.cpu_seconds)
            cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if fixed else "weak", "vs_baseline": None,
            "dtype": {3: "fp4 (e2m1 indicators, exact fp32 accumulation)", 2: "int8 (indicators, s32 accumulation)"}.get(variant, "u32 (bit planes, LOP3+POPC)"),
            "data": "synthetic",
            "config": {
                "workload": f"{args.workload}: {n_samples} samples x {n_sites} sites, missing {missing}, "
                            f"kin_threshold {thr}, max_results {max_results}"
                            + ((f" (fixed cohort, tile grid split over {n_gpus} GPUs)" if fixed else
                                f" (weak scaling: {n1}*sqrt({n_gpus}) samples, tile grid split over {n_gpus} GPUs)") if n_gpus > 1 else ""),
                "tiles": tiles, "retained_pairs": retained, "kernel_variant": variant,
                "l2": (f"inputs larger than L2: {(-(-n_samples // 64) * 64 * words * 16) >> 20} MiB of genotype codes streamed per step" if umma else
                       f"inputs larger than L2: {(-(-n_samples // 64) * 64 * words * 12) >> 20} MiB of compute planes streamed per step"),
                "input_synthesis_s": round(synth_s, 3),
            },
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "reference_gpu_kernel": ref_gpu, "pack": pack, "plane_exchange": exchange,
        }
        print(json.dumps(line), flush=True)
    planes.close()
    ctx.close()
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
