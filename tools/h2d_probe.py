#!/usr/bin/env python
"""Host->device copy ceiling of one box, per GPU and in aggregate (developer probe).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/h2d_probe.py > gpurun_out/h2d_probe.json

Every rank owns one GPU.  For each *placement* of the pinned source buffer (default, CPU affinity bound to the
GPU's NUMA node before allocating, explicit MPOL_BIND, write-combined) and each *active set* of ranks, the active
ranks copy `--mb` MB `--reps` times with two copies in flight, timed with CUDA events between two barriers.
Rank 0 prints one JSON document: the box topology and a GB/s table.  Nothing of the product runs here.
"""
from __future__ import annotations

import argparse
import ctypes
import glob
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

from semantic_depth_b200 import hostmem


def topo_report():
    out = {}
    for name, cmd in (("topo", ["nvidia-smi", "topo", "-m"]),
                      ("gpus", ["nvidia-smi", "--query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current",
                                "--format=csv,noheader"])):
        try:
            out[name] = subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout
        except Exception as e:
            out[name] = f"failed: {e}"
    out["nodes"] = {}
    for n in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        try:
            mem = open(os.path.join(n, "meminfo")).read().split("\n")[0:2]
            out["nodes"][os.path.basename(n)] = {"cpulist": open(os.path.join(n, "cpulist")).read().strip(), "mem": mem}
        except Exception as e:
            out["nodes"][os.path.basename(n)] = str(e)
    out["affinity"] = sorted(os.sched_getaffinity(0))
    try:
        out["status"] = [l.strip() for l in open("/proc/self/status") if l.startswith(("Cpus_allowed_list", "Mems_allowed_list"))]
    except Exception:
        pass
    out["cpu_count"] = os.cpu_count()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=210)
    ap.add_argument("--reps", type=int, default=24)
    ap.add_argument("--modes", default="default,affinity,mbind,wc,register")
    a = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG", "NONE")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = a.mb << 20
    dst = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    info = hostmem.gpu_locality(local)
    sets = [[0]]
    if world >= 2:
        sets += [[0, 1]]
    if world >= 4:
        sets += [[0, 2], [0, 1, 2, 3]]
    if world >= 8:
        sets += [[0, 4], [4, 5, 6, 7], [0, 2, 4, 6], list(range(8))]
    full_affinity = sorted(os.sched_getaffinity(0))
    results = []
    for mode in a.modes.split(","):
        os.sched_setaffinity(0, full_affinity)
        hostmem.set_mempolicy_default()
        placed = {"mode": mode}
        try:
            if mode == "default":
                src = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
            elif mode == "affinity":
                placed["bound"] = hostmem.bind_to_gpu(local, mempolicy=False)
                src = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
            elif mode == "mbind":
                placed["bound"] = hostmem.bind_to_gpu(local, mempolicy=True)
                src = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
            elif mode == "wc":
                placed["bound"] = hostmem.bind_to_gpu(local, mempolicy=True)
                src = [hostmem.pinned_empty(nbytes, write_combined=True) for _ in range(2)]
            elif mode == "register":
                placed["bound"] = hostmem.bind_to_gpu(local, mempolicy=True)
                src = [hostmem.registered_empty(nbytes) for _ in range(2)]
            else:
                continue
            for s in src:
                s.fill_(1)
            placed["pages_on_node"] = hostmem.node_histogram(src[0].data_ptr(), nbytes)
        except Exception as e:
            placed["error"] = repr(e)
            src = None
        ok = torch.tensor([1.0 if src is not None else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0.0:
            results.append({"rank": rank, **placed, "skipped": True})
            continue
        for act in sets:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            gbs = None
            if rank in act:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                for w in range(2):                                      # warm
                    with torch.cuda.stream(streams[w]):
                        dst[w].copy_(src[w], non_blocking=True)
                torch.cuda.synchronize()
                e0.record()
                for s in streams:
                    s.wait_event(e0)
                for i in range(a.reps):
                    with torch.cuda.stream(streams[i % 2]):
                        dst[i % 2].copy_(src[i % 2], non_blocking=True)
                for s in streams:
                    torch.cuda.current_stream().wait_stream(s)
                e1.record()
                torch.cuda.synchronize()
                gbs = a.reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
            if world > 1:
                dist.barrier()
            results.append({"rank": rank, "mode": mode, "active": act, "gb_per_s": gbs})
        results.append({"rank": rank, **placed})
        del src
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, {"rank": rank, "locality": info, "results": results})
    else:
        gathered = [{"rank": rank, "locality": info, "results": results}]
    if rank == 0:
        table = {}
        for g in gathered:
            for r in g["results"]:
                if r.get("gb_per_s") is not None:
                    key = f"{r['mode']}|{','.join(map(str, r['active']))}"
                    table.setdefault(key, {})[g["rank"]] = round(r["gb_per_s"], 2)
        summary = {k: {"per_gpu": v, "aggregate": round(sum(v.values()), 1), "min": min(v.values())} for k, v in table.items()}
        place = [{k: v for k, v in r.items()} for g in gathered for r in g["results"] if "active" not in r]
        print(json.dumps({"topology": topo_report(), "locality": [g["locality"] for g in gathered], "summary": summary,
                          "placement": place}, indent=1))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
