#!/usr/bin/env python3
"""Summarises ncu artefacts brought back in gpurun_out/ into small text files under profiles/.

  tools/ncu_summary.py launches gpurun_out/r01_launches.csv            > profiles/r01_launches_summary.txt
  tools/ncu_summary.py full     gpurun_out/r01_king_csa.ncu-rep [...]   > profiles/r01_king_kernel_ncu.txt
"""
import collections
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]


def launches(path):
    rows = [r for r in csv.DictReader(l for l in open(path) if l.startswith('"'))]
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(r["Kernel Name"], [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e6
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot:.3f} ms total (ncu --metrics gpu__time_duration.sum --clock-control none)")
    print(f"# {'n':>4} {'total ms':>11} {'share':>7} {'grid':>16} {'block':>12}  kernel")
    for k, v in agg.items():
        print(f"  {v[0]:4d} {v[1]:11.3f} {100 * v[1] / tot:6.2f}% {v[2]:>16} {v[3]:>12}  {k}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, unit = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"# {path}\n# kernel: {name}")
        for i, h in enumerate(hdr):
            if h in KEEP:
                print(f"  {h:70s} {vals[i]:>18} {unit[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h:
                try:
                    stalls.append((float(vals[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(s for s, _ in stalls) or 1.0
        print("  warp-state samples (pc sampling):")
        for s, n in sorted(stalls, reverse=True)[:10]:
            print(f"    {n:28s} {100 * s / tot:6.2f}%")
        print()


if __name__ == "__main__":
    mode, paths = sys.argv[1], sys.argv[2:]
    for p in paths:
        (launches if mode == "launches" else full)(p)
