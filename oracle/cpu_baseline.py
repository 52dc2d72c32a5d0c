"""CPU baseline harness -- TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__).

Times the oracle (the NumPy/SciPy port of the reference's fusion section) on the host cores:
frame-parallel over processes, one frame per worker (frames are independent), BLAS /
OpenMP pinned to one thread per worker, plus ``kd_workers`` threads inside each worker's cKDTree
queries when the host has more cores than workers.  Workers rebuild their frame from its seed, so
nothing large crosses process boundaries and frame synthesis stays outside the timed section.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

_FRAME = {}


def _init(height, width):
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    _FRAME["shape"] = (height, width)


def _prepare(seed):
    from semantic_depth_b200 import scene
    h, w = _FRAME["shape"]
    _FRAME["data"] = scene.make_frame(h, w, seed)
    return seed


def _run(kd_workers):
    from oracle import frame_ref
    logits, disp, intr = _FRAME["data"]
    t = time.perf_counter()
    o = frame_ref.fuse_frame(logits, disp, intr.as_q32(), intr.disparity_mult, workers=kd_workers)
    dt = time.perf_counter() - t
    return dt, o["rw"], o["f2f"]


def _prepare_scores(seed):
    from semantic_depth_b200 import scene
    h, w = _FRAME["shape"]
    _FRAME["scores"] = scene.make_frame_scores(h, w, seed)     # (second_skip scores, up-sampling weights, bias, disp, intr)
    return seed


def _run_scores(kd_workers):
    """Score-map arm: the CPU also evaluates FCN-8s' last transposed convolution (fcn8s/fcn.py:207-213) before the path."""
    from oracle import frame_ref
    sc, wts, bs, disp, intr = _FRAME["scores"]
    t = time.perf_counter()
    logits = frame_ref.upsample_scores(sc, wts, bs)
    o = frame_ref.fuse_frame(logits, disp, intr.as_q32(), intr.disparity_mult, workers=kd_workers)
    dt = time.perf_counter() - t
    return dt, o["rw"], o["f2f"]


class CpuBaseline:
    """Pool of worker processes; ``step()`` runs one frame per worker and returns the wall time."""

    def __init__(self, height: int, width: int, procs: int | None = None, kd_workers: int | None = None):
        cores = os.cpu_count() or 1
        self.procs = max(1, min(cores, 32) if procs is None else procs)
        self.kd_workers = max(1, cores // self.procs) if kd_workers is None else kd_workers
        self.cores_used = min(cores, self.procs * self.kd_workers)
        self.height, self.width = height, width
        ctx = mp.get_context("spawn")
        # one single-process pool per worker so that each frame stays pinned to its process
        self.pools = [ctx.Pool(1, initializer=_init, initargs=(height, width)) for _ in range(self.procs)]
        self.seed = 0

    def prepare(self):
        res = [p.apply_async(_prepare, (self.seed + i,)) for i, p in enumerate(self.pools)]
        for r in res:
            r.get()
        self.seed += self.procs

    def step(self):
        """One frame per worker, in parallel.  Returns (wall seconds, frames, per-frame answers)."""
        self.prepare()
        t = time.perf_counter()
        res = [p.apply_async(_run, (self.kd_workers,)) for p in self.pools]
        out = [r.get() for r in res]
        wall = time.perf_counter() - t
        return wall, self.procs, out

    def step_scores(self):
        """``step`` in the score-map mode (inputs = FCN-8s second_skip scores + disparities)."""
        res = [p.apply_async(_prepare_scores, (self.seed + i,)) for i, p in enumerate(self.pools)]
        for r in res:
            r.get()
        self.seed += self.procs
        t = time.perf_counter()
        res = [p.apply_async(_run_scores, (self.kd_workers,)) for p in self.pools]
        out = [r.get() for r in res]
        return time.perf_counter() - t, self.procs, out

    def close(self):
        for p in self.pools:
            p.terminate()
        self.pools = []

    def describe(self) -> str:
        return (f"{self.procs} frame(s) of {self.height}x{self.width} per step, one per process, "
                f"{self.kd_workers} cKDTree thread(s) each (oracle port of semantic_depth.py:183-324 + pcl.py; "
                f"cKDTree stands in for Open3D)")
