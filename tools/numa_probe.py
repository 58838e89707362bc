#!/usr/bin/env python3
"""Host <-> device copy bandwidth of N concurrent ranks with page-locked buffers allocated (a) wherever the unbound
process happens to run and (b) after binding the process to the CPUs NVML reports as local to its GPU (the pinned pages
are then first-touched on that NUMA node).  torchrun --nproc-per-node N tools/numa_probe.py"""
import json
import os
import subprocess
import time

import torch
import torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def measure(tag):
    n = 2 << 30
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    out = {}
    for name, fn in (("h2d", lambda: d.copy_(host, non_blocking=True)), ("d2h", lambda: host.copy_(d, non_blocking=True))):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        barrier()
        out[name] = n / dt / 1e9
    t = torch.tensor([out["h2d"], out["d2h"]], device=dev, dtype=torch.float64)
    if world > 1:
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
    else:
        allt = [t]
    if rank == 0:
        print(json.dumps({"probe": "numa", "mode": tag, "ranks": world,
                          "h2d_gbs_per_rank": [round(float(x[0]), 1) for x in allt],
                          "d2h_gbs_per_rank": [round(float(x[1]), 1) for x in allt],
                          "h2d_total": round(sum(float(x[0]) for x in allt), 1), "d2h_total": round(sum(float(x[1]) for x in allt), 1)}), flush=True)
    del host, d


if rank == 0:
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"]):
        try:
            print(subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout[-3000:], flush=True)
        except Exception as exc:
            print(cmd, exc)
measure("unbound")
try:
    import pynvml

    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    before = len(os.sched_getaffinity(0))
    pynvml.nvmlDeviceSetCpuAffinity(h)
    after = sorted(os.sched_getaffinity(0))
    print(json.dumps({"rank": rank, "cpus_before": before, "cpus_after": len(after), "first": after[:2], "last": after[-2:]}), flush=True)
except Exception as exc:
    print("affinity failed", repr(exc), flush=True)
measure("bound to the GPU's CPUs (NVML)")
if world > 1:
    dist.destroy_process_group()
