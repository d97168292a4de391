#!/usr/bin/env python
"""Pinned host -> device copy bandwidth per GPU when 1, 2, 4, 8 ranks upload at the same time (the e2e limiter).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 \
        benchmarks/h2d_scaling.py --json gpurun_out/h2d_scaling.json

Every rank owns one GPU and a pinned buffer of --mb megabytes (655 MB = one bench step's upload of 64 images).
For k in (1, 2, 4, 8): ranks < k copy `--reps` times concurrently (plain cudaMemcpyAsync via torch, CUDA events), the
others idle.  Repeated with the pinned buffer allocated (a) wherever the process happens to run and (b) after binding
the process to the GPU's NVML-reported CPU affinity (parallel.bind_to_gpu_numa), and for D2H.  Also prints the box's
topology (nvidia-smi topo -m, NUMA nodes, cores)."""
import argparse
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from faster_rcnn_b200 import parallel      # noqa: E402


def copy_gbs(dst, src, reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    return src.numel() * src.element_size() * reps / (a.elapsed_time(b) / 1e3) / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=655)
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    rank, world, local = parallel.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n = args.mb * 1000 * 1000 // 4
    d = torch.empty(n, dtype=torch.float32, device=dev)
    res = {"world": world, "mb": args.mb, "runs": []}
    groups = [k for k in (1, 2, 4, 8) if k <= world]
    for binding in ("unbound", "gpu_numa"):
        cores = None
        if binding == "gpu_numa":
            cores = parallel.bind_to_gpu_numa(local)
        h = torch.empty(n, dtype=torch.float32).pin_memory()
        h.fill_(1.0)                                           # first touch under the current affinity
        for direction in ("h2d", "d2h"):
            for k in groups:
                if world > 1:
                    dist.barrier()
                gbs = 0.0
                if rank < k:
                    gbs = copy_gbs(d, h, args.reps) if direction == "h2d" else copy_gbs(h, d, args.reps)
                t = torch.tensor([gbs], device=dev, dtype=torch.float64)
                allv = [torch.zeros_like(t) for _ in range(world)]
                if world > 1:
                    dist.all_gather(allv, t)
                else:
                    allv = [t]
                per = [round(float(v.item()), 2) for v in allv[:k]]
                if rank == 0:
                    row = {"binding": binding, "direction": direction, "active_ranks": k, "per_gpu_GBps": per,
                           "sum_GBps": round(sum(per), 1), "min_GBps": min(per),
                           "bound_cores": len(cores) if cores else None}
                    res["runs"].append(row)
                    print(json.dumps(row), flush=True)
        del h
    if rank == 0:
        def sh(cmd):
            try:
                return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout
            except Exception as e:      # noqa: BLE001
                return repr(e)
        res["topo"] = sh("nvidia-smi topo -m")
        res["numa"] = sh("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'")
        res["mem"] = sh("grep -E 'MemTotal|Hugepagesize|HugePages_Total' /proc/meminfo")
        res["pcie"] = sh("nvidia-smi --query-gpu=index,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv")
        print(res["topo"])
        print(res["numa"])
        print(res["pcie"])
        if args.json:
            os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
            json.dump(res, open(args.json, "w"), indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
