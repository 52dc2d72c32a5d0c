"""Placement of the pinned host buffers that feed one GPU (host plumbing of the end-to-end path).

The reference reads each frame from disk into host memory and hands it to the networks
(/root/reference/semantic_depth.py:867-901); here the host-facing call copies a batch of raw network outputs
(210 MB at 5 x 1024x2048) to the GPU every step.  On a two-socket 8-GPU box that copy only runs at PCIe speed
when the pinned pages live on the NUMA node the GPU hangs off: ``bind_to_gpu`` pins the calling process to the
CPUs local to the GPU (sysfs ``local_cpulist`` of the PCI function) and, optionally, sets the memory policy to
that node *before* anything is allocated and pinned, so that ``cudaHostAlloc`` / first touch place the pages there.

Linux only; every function degrades to a no-op (and says so in its return value) when sysfs or the
syscalls are not available -- placement is an optimisation, never a correctness requirement.
"""
from __future__ import annotations

import ctypes
import mmap
import os
import platform

import torch

_SYS = {"x86_64": {"mbind": 237, "set_mempolicy": 238, "get_mempolicy": 239, "move_pages": 279},
        "aarch64": {"mbind": 235, "set_mempolicy": 237, "get_mempolicy": 236, "move_pages": 239}}.get(platform.machine(), {})
MPOL_DEFAULT, MPOL_PREFERRED, MPOL_BIND = 0, 1, 2
_libc = ctypes.CDLL(None, use_errno=True)


def _parse_cpulist(s: str) -> list[int]:
    out = []
    for part in s.strip().split(","):
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-")
            out.extend(range(int(lo), int(hi) + 1))
        else:
            out.append(int(part))
    return out


def gpu_locality(index: int) -> dict:
    """NUMA node and local CPUs of CUDA device ``index`` (as torch numbers it), from sysfs."""
    info = {"gpu": index, "numa_node": -1, "local_cpus": [], "bdf": None}
    try:
        p = torch.cuda.get_device_properties(index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        info["bdf"] = bdf
        base = f"/sys/bus/pci/devices/{bdf}"
        info["numa_node"] = int(open(f"{base}/numa_node").read().strip())
        info["local_cpus"] = _parse_cpulist(open(f"{base}/local_cpulist").read())
    except Exception as e:            # no sysfs (container) or no such attribute
        info["error"] = repr(e)
    if info["numa_node"] >= 0 and not info["local_cpus"]:
        try:
            info["local_cpus"] = _parse_cpulist(open(f"/sys/devices/system/node/node{info['numa_node']}/cpulist").read())
        except Exception:
            pass
    return info


def set_mempolicy(mode: int, node: int | None) -> bool:
    if "set_mempolicy" not in _SYS:
        return False
    if node is None or node < 0:
        r = _libc.syscall(_SYS["set_mempolicy"], MPOL_DEFAULT, None, 0)
        return r == 0
    nwords = node // 64 + 1
    mask = (ctypes.c_ulong * nwords)()
    mask[node // 64] = 1 << (node % 64)
    r = _libc.syscall(_SYS["set_mempolicy"], mode, ctypes.byref(mask), ctypes.c_ulong(64 * nwords + 1))
    return r == 0


def set_mempolicy_default() -> bool:
    return set_mempolicy(MPOL_DEFAULT, None)


def bind_to_gpu(index: int, mempolicy: bool = True, rank_in_node: int | None = None, ranks_per_node: int | None = None) -> dict:
    """Restrict the calling process to the CPUs next to GPU ``index`` and prefer that node's memory.

    Call BEFORE allocating / pinning the host buffers.  If several ranks share a node the local CPUs are
    split evenly between them when ``rank_in_node`` / ``ranks_per_node`` are given.  Returns what was done."""
    info = gpu_locality(index)
    allowed = sorted(os.sched_getaffinity(0))
    local = [c for c in info["local_cpus"] if c in allowed]
    done = {"gpu": index, "numa_node": info["numa_node"], "cpus": None, "mempolicy": False}
    if local:
        if rank_in_node is not None and ranks_per_node and len(local) >= ranks_per_node:
            per = len(local) // ranks_per_node
            local = local[rank_in_node * per:(rank_in_node + 1) * per]
        try:
            os.sched_setaffinity(0, local)
            done["cpus"] = f"{local[0]}-{local[-1]} ({len(local)})"
        except OSError as e:
            done["cpus_error"] = repr(e)
    if mempolicy and info["numa_node"] >= 0:
        done["mempolicy"] = set_mempolicy(MPOL_PREFERRED, info["numa_node"])
    return done


def node_histogram(ptr: int, nbytes: int, samples: int = 256) -> dict:
    """On which NUMA nodes the pages of [ptr, ptr+nbytes) live (move_pages query on a sample of pages)."""
    if "move_pages" not in _SYS or nbytes <= 0:
        return {}
    page = mmap.PAGESIZE
    npages = max(1, nbytes // page)
    n = min(samples, npages)
    pages = (ctypes.c_void_p * n)(*[(ptr // page * page) + (i * npages // n) * page for i in range(n)])
    status = (ctypes.c_int * n)()
    r = _libc.syscall(_SYS["move_pages"], 0, ctypes.c_ulong(n), pages, None, status, 0)
    if r != 0:
        return {"error": ctypes.get_errno()}
    hist: dict = {}
    for s in status:
        hist[int(s)] = hist.get(int(s), 0) + 1
    return hist


# ---------------------------------------------------------------------------------------------------
# pinned allocations with explicit flags (torch's pin_memory() is cudaHostAlloc(default) under the hood)
# ---------------------------------------------------------------------------------------------------
def _cudart():
    import glob
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            return ctypes.CDLL(name)
        except OSError:
            pass
    for cand in glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*")):
        return ctypes.CDLL(cand)
    raise OSError("libcudart not found")


class _HostBlock:
    def __init__(self, ptr, nbytes, free):
        self.ptr, self.nbytes, self._free = ptr, nbytes, free

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass


def _as_tensor(ptr: int, nbytes: int, owner) -> torch.Tensor:
    buf = (ctypes.c_uint8 * nbytes).from_address(ptr)
    buf._sd_owner = owner        # torch.frombuffer keeps `buf` alive for as long as any view of the storage exists
    return torch.frombuffer(buf, dtype=torch.uint8)


def pinned_empty(nbytes: int, write_combined: bool = False) -> torch.Tensor:
    """uint8 tensor over cudaHostAlloc memory (portable; optionally write-combined: fast for the device to read over
    PCIe, slow for the CPU to read back -- for upload-only staging)."""
    rt = _cudart()
    p = ctypes.c_void_p()
    flags = 0x01 | (0x04 if write_combined else 0)     # cudaHostAllocPortable | cudaHostAllocWriteCombined
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    if rc != 0:
        raise RuntimeError(f"cudaHostAlloc failed: {rc}")
    blk = _HostBlock(p.value, nbytes, lambda: rt.cudaFreeHost(ctypes.c_void_p(p.value)))
    return _as_tensor(p.value, nbytes, blk)


def registered_empty(nbytes: int) -> torch.Tensor:
    """uint8 tensor over an anonymous mapping, first-touched by the caller (so it follows the calling thread's CPU
    affinity / memory policy) and then page-locked with cudaHostRegister."""
    rt = _cudart()
    m = mmap.mmap(-1, nbytes, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    buf = (ctypes.c_uint8 * nbytes).from_buffer(m)
    ptr = ctypes.addressof(buf)
    t = torch.frombuffer(buf, dtype=torch.uint8)
    t.zero_()                                          # first touch
    rc = rt.cudaHostRegister(ctypes.c_void_p(ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(0x01))
    if rc != 0:
        raise RuntimeError(f"cudaHostRegister failed: {rc}")
    buf._sd_owner = (_HostBlock(ptr, nbytes, lambda: rt.cudaHostUnregister(ctypes.c_void_p(ptr))), m)
    return t


# ---------------------------------------------------------------------------------------------------
# which GPUs of the box to use when the job has fewer ranks than the box has GPUs
# ---------------------------------------------------------------------------------------------------
def candidate_device_maps(world: int, visible: int) -> dict:
    """Rank -> device maps worth trying when ``world < visible``: the first GPUs, the last GPUs, every
    (visible/world)-th GPU.  On a two-socket box whose host memory sits behind one socket the GPUs of the other
    socket share one inter-socket link for their host reads (measured on this pool's 8 x B200 boxes: GPUs 0-3
    together get 115 GB/s, GPUs 4-7 together 221 GB/s, tools/h2d_probe.py), and a hypervisor can hide that
    topology from sysfs -- so the maps are *measured*, not derived."""
    maps = {"first": list(range(world))}
    if visible > world:
        maps["last"] = list(range(visible - world, visible))
        maps["spread"] = [i * visible // world for i in range(world)]
    return maps


def measure_h2d_gbs(device: int, mbytes: int = 96, reps: int = 6, barrier=None) -> float:
    """Pinned host -> device copy rate of ``device`` (two copies in flight, CUDA events).  ``barrier`` (a callable) is
    invoked right before the timed copies so that all ranks of a job copy at the same time."""
    nbytes = mbytes << 20
    with torch.cuda.device(device):
        src = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
        dst = [torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{device}") for _ in range(2)]
        streams = [torch.cuda.Stream(device=device) for _ in range(2)]
        for i in range(2):
            src[i].fill_(i + 1)
            with torch.cuda.stream(streams[i]):
                dst[i].copy_(src[i], non_blocking=True)
        torch.cuda.synchronize(device)
        if barrier is not None:
            barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams:
            s.wait_event(e0)
        for i in range(reps):
            with torch.cuda.stream(streams[i % 2]):
                dst[i % 2].copy_(src[i % 2], non_blocking=True)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize(device)
        ms = e0.elapsed_time(e1)
        del src, dst
        torch.cuda.empty_cache()
    return reps * nbytes / (ms * 1e-3) / 1e9


def choose_device(local_rank: int, world: int, all_reduce_min=None, barrier=None, min_gain: float = 1.10) -> tuple[int, dict]:
    """Device for this rank.  With as many ranks as GPUs (or one rank) it is ``local_rank``.  Otherwise every
    candidate map is timed with all ranks copying at once and the map with the best *slowest* rank wins (the
    plain first-N map unless another one is at least ``min_gain`` x better).  ``all_reduce_min(x) -> float`` and
    ``barrier()`` are the job's CPU-side collectives (gloo); without them the first-N map is used."""
    visible = torch.cuda.device_count()
    report = {"visible_gpus": visible, "world": world, "map": "first", "devices": list(range(min(world, visible)))}
    if world <= 1 or visible <= world or all_reduce_min is None or os.environ.get("SD_DEVICE_MAP", "") == "first":
        return local_rank % max(visible, 1), report
    maps = candidate_device_maps(world, visible)
    forced = os.environ.get("SD_DEVICE_MAP", "")
    rates = {}
    for name, m in maps.items():
        if forced and name != forced:
            continue
        mine = measure_h2d_gbs(m[local_rank], barrier=barrier)
        rates[name] = float(all_reduce_min(mine))
    best = max(rates, key=lambda k: rates[k])
    if "first" in rates and rates[best] < min_gain * rates["first"]:
        best = "first"
    report.update({"map": best, "devices": maps[best], "slowest_rank_h2d_gb_per_s": {k: round(v, 2) for k, v in rates.items()}})
    return maps[best][local_rank], report
