#!/usr/bin/env python
"""Concurrent device-to-host copy rate of N GPUs of one box into pinned host buffers (plain cudaMemcpyAsync, no
kernels): what bounds the end-to-end leg of bench.py at N > 1, where every rank pulls its rows of A (8 GB in all)
through the host's PCIe / memory system at once.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/tools/d2h_concurrency.py [--numa]

--numa binds every rank (and therefore the pages of its pinned buffer, first touch) to the CPUs `nvidia-smi topo -m`
lists as local to its GPU.  Prints one JSON line: per-rank and aggregate GB/s, alone and all together."""
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist


def local_cpus(gpu):
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout.splitlines()
        hdr = next(l for l in out if "CPU Affinity" in l)
        col = hdr.split("\t").index("CPU Affinity") if "\t" in hdr else None
        row = next(l for l in out if l.startswith(f"GPU{gpu}\t") or l.startswith(f"GPU{gpu} "))
        cell = row.split("\t")[col].strip() if col is not None else row.split()[-3]
        cpus = set()
        for part in cell.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return cpus
    except Exception:
        return None


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = "--numa" in sys.argv
    bound = None
    if numa:
        bound = local_cpus(local)
        if bound:
            os.sched_setaffinity(0, bound)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << 27                                  # 1 GiB of doubles per rank
    dev = torch.ones(n, dtype=torch.float64, device="cuda")
    host = torch.empty(n, dtype=torch.float64, pin_memory=True)
    host.fill_(0.0)                              # first touch here
    s = torch.cuda.Stream()

    def copy_rate(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s):
            for _ in range(reps):
                host.copy_(dev, non_blocking=True)
        s.synchronize()
        return reps * n * 8 / (time.perf_counter() - t0) / 1e9

    copy_rate(1)
    alone = []
    for r in range(world):                       # one rank at a time
        dist.barrier()
        v = copy_rate(4) if r == rank else 0.0
        t = torch.tensor([v], device="cuda"); dist.all_reduce(t); alone.append(float(t[0]))
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mine = copy_rate(8)                           # all together
    dist.barrier()
    wall = time.perf_counter() - t0
    t = torch.tensor([mine], device="cuda"); g = [torch.zeros_like(t) for _ in range(world)]; dist.all_gather(g, t)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "numa_bound": numa, "cpus_rank0": sorted(bound)[:4] + ["..."] if bound else None,
                          "alone_GBps": [round(x, 1) for x in alone],
                          "together_GBps_per_rank": [round(float(x[0]), 1) for x in g],
                          "together_aggregate_GBps": round(world * 8 * n * 8 / wall / 1e9, 1),
                          "bytes_per_rank": n * 8 * 8}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
