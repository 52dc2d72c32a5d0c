#!/usr/bin/env python
"""Benchmark of the fusion hot path: frames/s at 1024x2048 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the fused path over one batch of 5 synthetic Cityscapes-shaped frames
(BASELINE.json configs[1], "Munich-test-set-shaped batch of 5 frames at 1024x2048").  Ranks are
independent (frame-parallel, weak scaling, no collective on the data path).  One JSON line on rank 0.

  value     frames/s with the inputs already resident in HBM (CUDA events, max over ranks)
  e2e       frames/s through the host-facing API: pinned host inputs, H2D + fused path + D2H of the
            answers inside the timed region
  roofline  the dominant kernel of the step (sd::knn_kernel, the statistical filter's neighbour search;
            share in profiles/): SURVEY 8d's algorithmic bytes of that stage per launch / its mean device
            time measured with CUDA events inside the timed region, vs the measured HBM peak of
            MEASURED_PEAKS.json.  roofline_pixel = the same for the HBM-streaming pixel stage;
            roofline_path = the whole path's SURVEY 8d B_alg figure x frames/s
  cpu_baseline / --impl reference   the oracle port of the reference's CPU path on the host cores
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

# measured with tools/h2d_probe.py on this pool's 8 x B200 boxes (profiles/r2_h2d_probe.json): pinned host -> device copies
H2D_CEILING = {"one_gpu_gb_per_s": 55.5, "gpus_0_to_3_together_gb_per_s": 115.6, "gpus_4_to_7_together_gb_per_s": 221.2,
               "all_8_together_gb_per_s": 238.0, "all_8_slowest_gpu_gb_per_s": 23.6,
               "note": "the box is a VM that exposes ONE NUMA node: GPUs 0-3 share one inter-socket path to the host memory "
                       "(115.6 GB/s together), all 8 GPUs together reach 238 GB/s, placement (affinity, mbind, write-combined, "
                       "cudaHostRegister) changes nothing; with equal work per rank the slowest GPU (23.6 GB/s) sets the e2e rate "
                       "at N=8: 23.6 GB/s / 41.9 MB per frame x 8 = 4.5 k frames/s on logits",
               "source": "profiles/r2_h2d_probe.json"}
METRIC = "fusion_frames_per_sec_1024x2048"
UNIT = "frames/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--frames", type=int, default=5, help="frames per batch (= per step)")
    ap.add_argument("--slots", type=int, default=3, help="batches in flight per GPU")
    ap.add_argument("--batches", type=int, default=6, help="distinct input batches resident in HBM (cycled)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-kernel-timing", action="store_true", help="skip the per-stage timing pass after the timed region")
    ap.add_argument("--stage-reps", type=int, default=6, help="batches averaged by the per-stage timing pass")
    ap.add_argument("--skip-configs", action="store_true", help="skip BASELINE configs[3] / configs[4] (2M-point filter, RANSAC sweep)")
    ap.add_argument("--stream", type=int, default=0, help="BASELINE configs[2]: one stream of this many frames sharded over the ranks (strong scaling)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = max(steps, 100)")
    ap.add_argument("--cpu-procs", type=int, default=0)
    ap.add_argument("--ref-budget-s", type=float, default=240.0)
    ap.add_argument("--no-bind", action="store_true", help="do not bind the rank to its GPU's NUMA node before pinning host buffers")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def nvml_index(cuda_index):
    """NVML enumerates every GPU of the box; CUDA only the ones CUDA_VISIBLE_DEVICES lists."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    ids = [v.strip() for v in vis.split(",") if v.strip()]
    if ids and all(v.isdigit() for v in ids) and cuda_index < len(ids):
        return int(ids[cuda_index])
    return cuda_index


def workload_name(a):
    return f"batch of {a.frames} synthetic Cityscapes-shaped frames at {a.height}x{a.width} (BASELINE.json configs[1])"


# ------------------------------------------------------------------------------------------------
# clocks sampling (NVML, in-process)
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz, self.ok = index, [], set(), False, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.004)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# reference arm: the oracle port of the reference's CPU path on the host cores
# ------------------------------------------------------------------------------------------------
def run_reference(a):
    rank, _, world = dist_env()
    if rank != 0:
        return
    from oracle.cpu_baseline import CpuBaseline
    cb = CpuBaseline(a.height, a.width, procs=a.cpu_procs or None)
    t_first = None
    warm = 0
    # CPU code has nothing to warm but the process pool; one warm-up step also calibrates the budget
    for _ in range(min(a.warmup, 1)):
        t_first, _, _ = cb.step()
        warm += 1
    est = t_first if t_first else 25.0
    steps = max(1, min(a.steps, int((a.ref_budget_s - (t_first or 0.0)) // max(est, 1e-3))))
    total, frames = 0.0, 0
    for _ in range(steps):
        wall, n, _ = cb.step()
        total += wall
        frames += n
    # the same CPU path fed by FCN-8s second_skip scores (the CPU evaluates the up-sampling layer too): the arm that matches
    # the B200 line's e2e_score_map_mode
    sc_value = None
    if a.height % 8 == 0 and a.width % 8 == 0 and not a.skip_e2e:
        try:
            wall_sc, n_sc, _ = cb.step_scores()
            sc_value = n_sc / wall_sc
        except Exception:
            sc_value = None
    cb.close()
    value = frames / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
        "steps_requested": a.steps, "warmup": warm, "ms_per_step": 1e3 * total / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32 (NumPy)", "data": "synthetic",
        "config": {"workload": workload_name(a), "height": a.height, "width": a.width, "frames_per_step": cb.procs,
                   "note": "CPU arm: each step = one full-resolution frame per worker process (bounded sample of the "
                           "5-frame batch workload); executed steps are capped by --ref-budget-s"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cb.cores_used, "kind": "port", "sample": cb.describe()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cores": os.cpu_count(),
        "e2e_score_map_mode": {"value": sc_value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "steps": 1,
                               "note": "one more step with the inputs of the score-map mode: upsample_scores (fcn8s/fcn.py:207-213) + the path"},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def golden_check(h: int, w: int, res, seeds):
    """The frames of rank 0's first batch are the seed-0..4 frames: compare their answers with the committed fixtures that
    the reference's own pcl.py chain produced (tests/golden/make_golden.py).  Reads data files only; nothing of oracle/
    runs here."""
    import numpy as np
    out = []
    for f, seed in enumerate(seeds):
        path = os.path.join(ROOT, "tests", "golden", f"frame_{h}x{w}_seed{seed}.npz")
        if not os.path.exists(path):
            continue
        g = np.load(path)
        counts = res.counts(f)
        bad = [k for k, v in counts.items() if int(g[f"count/{k}"]) != int(v)]
        out.append({"frame": f, "fixture": os.path.relpath(path, ROOT), "stage_counts_equal": not bad, "mismatched_stages": bad,
                    "rw_bit_equal": bool(float(g["rw"]) == float(res.rw[f])), "f2f_abs_err": abs(float(g["f2f"]) - float(res.f2f[f])),
                    "status_equal": bool(int(g["status"]) == int(res.status[f]))})
    if not out:
        return None
    return {"frames_checked": len(out), "all_stage_counts_equal": all(o["stage_counts_equal"] for o in out),
            "all_rw_bit_equal": all(o["rw_bit_equal"] for o in out), "max_f2f_abs_err": max(o["f2f_abs_err"] for o in out),
            "all_status_equal": all(o["status_equal"] for o in out), "per_frame": out}


def b_alg_bytes(counts: dict, hw: int) -> float:
    """20*H*W + 12*(N_R0+N_F0) + 12*sum_stages(N_in+N_out) + 12*N_slab_in."""
    c = counts
    stages = [("road_gather", "road_z"), ("road_z", "road_mad_y"), ("road_mad_y", "road_mad_x"), ("road_mad_x", "road_plane"),
              ("road_plane", "road_sor"), ("road_sor", "road_ror"),
              ("fence_gather", "fence_mad_y"), ("fence_mad_y", "fence_abs_z"),
              ("left_split", "left_mad_x"), ("left_mad_x", "left_plane"), ("right_split", "right_mad_x"),
              ("right_mad_x", "right_plane")]
    total = 20.0 * hw + 12.0 * (c["road_gather"] + c["fence_gather"])
    total += 12.0 * sum(c[i] + c[o] for i, o in stages)
    total += 12.0 * (c["fence_abs_z"] + c["left_split"] + c["right_split"])      # extract_pcls: 1 in, 2 out
    total += 12.0 * c["road_ror"]                                                 # slab scan input
    return total


def secondary_configs(dev):
    """BASELINE.json configs[3] and configs[4] on the bench GPU (rank 0, N = 1): the 2 M-point statistical filter (k = 16)
    and the RANSAC scoring sweep at 16 384 hypotheses.  CUDA-event times of the product's own per-call ops."""
    import numpy as np
    import torch
    from semantic_depth_b200 import scene
    from semantic_depth_b200.pcl_gpu import engine_for

    def timed(fn, reps):
        fn(); torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            fn()
        t1.record(); torch.cuda.synchronize()
        return t0.elapsed_time(t1) / reps

    out = {}
    n = 2_000_000
    pts = torch.from_numpy(scene.make_road_cloud(n, seed=0)).to(dev)
    x, y, z = (pts[:, i].contiguous() for i in range(3))
    eng = engine_for(n, dev)
    ms = timed(lambda: eng.knn_mean_distance(x, y, z, 16, 0.5), 5)
    n_out = int((eng.knn_mean_distance(x, y, z, 16, 0.5)[0] > 0).sum())
    out["config4"] = {"workload": "BASELINE.json configs[3]: 2M-point synthetic road cloud, statistical filter k=16 (grid build + k-NN + "
                                  "cloud statistics, one host sync)", "ms": ms, "value": n / (ms * 1e-3), "unit": "points/s",
                      "algorithmic_bytes": 12.0 * (n + n_out), "gb_per_s": 12.0 * (n + n_out) / (ms * 1e-3) / 1e9}
    m, K = 450_000, 16384
    road = torch.from_numpy(scene.make_road_cloud(m, seed=1)).to(dev)
    rx, ry, rz = (road[:, i].contiguous() for i in range(3))
    trip = torch.from_numpy(np.random.default_rng(1234).integers(0, m, (K, 3)).astype(np.int32)).to(dev)
    ms = timed(lambda: eng.ransac_score(rx, ry, rz, 1, 5.0, trip), 3)
    out["config5"] = {"workload": f"BASELINE.json configs[4]: RANSAC scoring, {K} seeded hypotheses x {m} points (fp64, 5 flop per test)",
                      "ms": ms, "value": K * m / (ms * 1e-3), "unit": "point-hypothesis tests/s",
                      "fp64_tflops": 5.0 * K * m / (ms * 1e-3) / 1e12,
                      "note": "fp64 ALU bound, not HBM: 5 dependent fp64 operations per test (no FMA contraction: bit-exact counts)"}
    return out


def run_b200(a):
    import hashlib
    import numpy as np
    import torch
    import torch.distributed as dist

    from semantic_depth_b200 import hostmem, scene
    from semantic_depth_b200.engine import FusionEngine
    from semantic_depth_b200.params import FusionParams, Intrinsics
    from semantic_depth_b200.stream import FramePipeline, gather_results, pack_answers, shard_frames

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    dev_index, device_map = local_rank, {"map": "first", "world": world}
    if world > 1:
        # keep stdout to the one JSON line: NCCL writes its banner / debug output to stdout by default
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN", ""):
            os.environ["NCCL_DEBUG"] = "NONE"
        dist.init_process_group("cpu:gloo,cuda:nccl")

        def cpu_min(x):
            t = torch.tensor([x], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return float(t.item())

        def cpu_barrier():
            dist.all_reduce(torch.zeros(1))

        # fewer ranks than GPUs on the box: pick the GPUs whose host->device copies do not share one path
        # (measured with all ranks copying at once; tools/h2d_probe.py has the full table of this pool's boxes)
        dev_index, device_map = hostmem.choose_device(local_rank, world, None if a.no_bind else cpu_min, cpu_barrier)
    torch.cuda.set_device(dev_index)
    dev = torch.device("cuda", dev_index)
    # host placement BEFORE any pinned allocation: CPUs and memory of the NUMA node this rank's GPU hangs off
    # (a no-op where the hypervisor exposes a single node, as on this pool's boxes)
    binding = {"bound": False} if a.no_bind else hostmem.bind_to_gpu(dev_index)
    H, W, B, HW = a.height, a.width, a.frames, a.height * a.width
    P = FusionParams()
    intr = Intrinsics.synthetic(W)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(device_ids=[dev_index])

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if a.stream > 0:
        return run_stream(a, rank, world, dev, dev_index, device_map, P, intr, barrier, max_over_ranks)

    # ---- synthetic inputs: `batches` distinct batches, pinned on the host and resident in HBM
    nb = max(a.batches, a.slots)
    h_logits = [torch.empty((B, HW, 3), dtype=torch.float32).pin_memory() for _ in range(nb)]
    h_disp = [torch.empty((B, 2, H, W), dtype=torch.float32).pin_memory() for _ in range(nb)]
    for i in range(nb):
        lg, dp, _ = scene.make_batch(B, H, W, first_seed=rank * 100000 + i * B, intr=intr)
        h_logits[i].copy_(torch.from_numpy(lg))
        h_disp[i].copy_(torch.from_numpy(dp))
    d_logits = [t.to(dev) for t in h_logits]
    d_disp = [t.to(dev) for t in h_disp]
    input_bytes = nb * B * HW * 20

    # one CUDA graph per batch and slot; the only events inside the timed region bracket whole batches
    pipe = FramePipeline(H, W, B, slots=a.slots, device=dev, params=P, use_graphs=not a.no_graph)

    # ---- expected answers (also builds job tables and captures the graphs): one pass over every
    #      (slot, batch) pair that the timed loop will use
    expected = {}
    for b in range(nb):
        for tag, res in pipe.warm_device(d_logits[b], d_disp[b], intr, tag=b):
            expected.setdefault(tag, res)

    def run_device_steps(n, check):
        bad = 0
        for i in range(n):
            fin = pipe.submit_device(d_logits[i % nb], d_disp[i % nb], intr, tag=i % nb)
            if fin and check:
                bad += fin[1].raw.tobytes() != expected[fin[0]].raw.tobytes()
        for tag, res in pipe.drain():
            if check:
                bad += res.raw.tobytes() != expected[tag].raw.tobytes()
        return bad

    main = torch.cuda.current_stream()

    def timed_region(body):
        """barrier + synchronize | start event, every slot stream waits for it | body | main waits for the slots, end
        event | barrier + synchronize.  Returns (device ms, body's return value)."""
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(main)
        for s in pipe.slots:
            s.stream.wait_event(t0)
        out = body()
        for s in pipe.slots:
            main.wait_stream(s.stream)
        t1.record(main)
        barrier()
        return t0.elapsed_time(t1), out

    # ---- warm-up, then the timed region (device-resident inputs)
    run_device_steps(a.warmup, False)
    pipe.total_ms.clear()
    sampler = ClockSampler(nvml_index(dev_index))
    sampler.start()
    wall0 = time.perf_counter()
    elapsed_ms, mismatches = timed_region(lambda: run_device_steps(a.steps, True))
    wall = time.perf_counter() - wall0
    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    total_ms = list(pipe.total_ms)

    # ---- every stage by itself: one batch at a time through the library's own stage timers (eager launches, the fence
    #      chain on the same stream, nothing else on the GPU) -- the reference's tic/toc table, and the kernel times of
    #      the roofline entries
    stage_ms = {}
    if not a.no_kernel_timing:
        eng_t = FusionEngine(H, W, max_frames=B, device=dev)
        eng_t.enable_timing(True)
        acc = []
        for i in range(2 + a.stage_reps):
            eng_t.fuse_frames(d_logits[i % nb], d_disp[i % nb], intr, P)
            if i >= 2:
                acc.append(eng_t.stage_times())
        stage_ms = {k: float(np.mean([t[k] for t in acc])) for k in acc[0]}
        eng_t.close()
        del eng_t

    # ---- end to end: pinned host inputs -> H2D -> fused path -> D2H of the answers, pipelined over the slots
    e2e = None
    if not a.skip_e2e:
        ke = a.e2e_steps or max(a.steps, 100)      # its own step count (reported): a 3-deep pipeline needs more than a few steps to fill
        for i in range(min(3, ke)):
            pipe.submit_host(h_logits[i % nb], h_disp[i % nb], intr, tag=i % nb)
        pipe.drain()

        def body():
            bad = 0
            for i in range(ke):
                fin = pipe.submit_host(h_logits[i % nb], h_disp[i % nb], intr, tag=i % nb)
                if fin:
                    bad += fin[1].raw.tobytes() != expected[fin[0]].raw.tobytes()
            for tag, res in pipe.drain():
                bad += res.raw.tobytes() != expected[tag].raw.tobytes()
            return bad
        e2e_ms, bad = timed_region(body)
        mismatches += bad
        e2e = (ke, e2e_ms)

    # ---- the same end-to-end call in the score-map mode: FCN-8s' head is evaluated inside the
    #      label kernel, so 0.19 B/pixel of scores cross PCIe instead of 12 B/pixel of logits
    e2e_sc = None
    if not a.skip_e2e and H % 8 == 0 and W % 8 == 0:
        ke = a.e2e_steps or max(a.steps, 100)      # its own step count (reported): a 3-deep pipeline needs more than a few steps to fill
        nsb = min(nb, 3)
        h_scores = [torch.empty((B, H // 8, W // 8, 3), dtype=torch.float32).pin_memory() for _ in range(nsb)]
        upw = upb = None
        for i in range(nsb):
            for f in range(B):
                sc, wts, bs, _, _ = scene.make_frame_scores(H, W, seed=rank * 100000 + i * B + f, intr=intr)
                h_scores[i][f].copy_(torch.from_numpy(sc))
            upw, upb = torch.from_numpy(wts).to(dev), torch.from_numpy(bs).to(dev)
        for i in range(max(3, len(pipe.slots))):
            pipe.submit_host_scores(h_scores[i % nsb], upw, upb, h_disp[i % nsb], intr, tag=i % nsb)
        first = {t: r.raw.tobytes() for t, r in pipe.drain()}

        def body_sc():
            bad = 0
            for i in range(ke):
                fin = pipe.submit_host_scores(h_scores[i % nsb], upw, upb, h_disp[i % nsb], intr, tag=i % nsb)
                if fin:
                    bad += fin[1].raw.tobytes() != first.get(fin[0], fin[1].raw.tobytes())
            for tag, res in pipe.drain():
                bad += res.raw.tobytes() != first.get(tag, res.raw.tobytes())
            return bad
        sc_ms, bad = timed_region(body_sc)
        mismatches += bad
        e2e_sc = (ke, sc_ms)

    # ---- reduce over ranks (max time)
    elapsed_ms = max_over_ranks(elapsed_ms)
    if e2e:
        e2e = (e2e[0], max_over_ranks(e2e[1]))
    if e2e_sc:
        e2e_sc = (e2e_sc[0], max_over_ranks(e2e_sc[1]))
    mism = int(max_over_ranks(float(mismatches)))

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        counts0 = expected[0].counts(0)
        # algorithmic bytes per launch: pixel stage = sum over the batch of 20*HW + 12*(N_R0+N_F0);
        # statistical_outlier_removal stage = 12 B per point in + 12 B per surviving point out
        per_batch_pixel_bytes, per_batch_knn_bytes, per_batch_alg = [], [], []
        for b in range(nb):
            r = expected[b]
            cs = [r.counts(f) for f in range(B)]
            per_batch_pixel_bytes.append(sum(20.0 * HW + 12.0 * (c["road_gather"] + c["fence_gather"]) for c in cs))
            per_batch_knn_bytes.append(sum(12.0 * (c["road_plane"] + c["road_sor"]) for c in cs))
            per_batch_alg.append(sum(b_alg_bytes(c, HW) for c in cs))
        pix_bytes = float(np.mean(per_batch_pixel_bytes))
        knn_bytes = float(np.mean(per_batch_knn_bytes))
        knn_queries = float(np.mean([sum(expected[b].counts(f)["road_plane"] for f in range(B)) for b in range(nb)]))
        pix_ms = stage_ms.get("pixel", float("nan"))
        k_ms = stage_ms.get("road_knn", float("nan"))

        def gbs(nbytes, ms):
            return nbytes / (ms * 1e-3) / 1e9 if ms == ms and ms > 0 else None

        # ncu evidence of this round (tools/gpu_round_profiles.sh -> tools/ncu_extract.py): DRAM traffic, warp instructions
        # and lanes per instruction of the dominant kernel, one launch over the same 5-frame batch
        prof, prof_file = {}, None
        for cand in ("r2_ncu_summary.json", "r1_ncu_summary.json"):
            try:
                prof = json.load(open(os.path.join(ROOT, "profiles", cand)))
                prof_file = f"profiles/{cand}"
                break
            except Exception:
                pass

        def traffic_of(*kernels):
            t = [prof.get(k, {}).get("dram_bytes_per_launch") for k in kernels]
            return float(sum(t)) if t and all(v is not None for v in t) else None

        clocks = sampler.summary()
        knn_prof = prof.get("knn_kernel", {})
        issue = None
        if knn_prof.get("smsp__inst_executed.sum") and k_ms == k_ms:
            winst = float(knn_prof["smsp__inst_executed.sum"]) + float(prof.get("knn_heavy_kernel", {}).get("smsp__inst_executed.sum") or 0.0)
            lanes = float(knn_prof.get("smsp__thread_inst_executed_per_inst_executed.ratio") or 32.0)
            sm_hz = float(clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)) * 1e6
            peak_issue = 148 * 4 * sm_hz                      # one warp instruction per scheduler and cycle
            ach = winst / (k_ms * 1e-3)
            issue = {"warp_instructions_per_launch": winst, "warp_instructions_per_query": winst / knn_queries if knn_queries else None,
                     "lanes_active_of_32": lanes, "achieved_warp_inst_per_s": ach, "peak_warp_inst_per_s": peak_issue,
                     "frac_issue": ach / peak_issue, "frac_useful_lanes": ach / peak_issue * lanes / 32.0,
                     "source": f"{prof_file} (ncu --set full of one launch) / kernel_ms of this run"}
        knn_ach, pix_ach = gbs(knn_bytes, k_ms), gbs(pix_bytes, pix_ms)
        frames = a.steps * B * world
        value = frames / (elapsed_ms * 1e-3)
        alg_path = float(np.mean(per_batch_alg)) / B
        serial = sum(stage_ms.values()) if stage_ms else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 clouds / f64 reprojection, plane and k-NN arithmetic", "data": "synthetic",
            "config": {"workload": workload_name(a), "height": H, "width": W, "frames_per_step": B,
                       "batches_in_flight": a.slots, "cuda_graphs": not a.no_graph,
                       "l2_policy": f"inputs larger than L2: {nb} distinct batches ({input_bytes / 1e6:.0f} MB) cycled",
                       "parallelism": f"frame-parallel x{world} (no collective on the data path)",
                       "params": "reference literals (semantic_depth.py:206-309), approach=both, SOR k=10, ROR r=0.5/80"},
            "impl": "b200",
            "gpu_launches": int(pipe.slots[0].engine.kernel_count(P)) * a.steps,
            "result_mismatches_vs_first_pass": mism,
            "answers_frame0": {"rw": float(expected[0].rw[0]), "f2f": float(expected[0].f2f[0]), "counts": counts0},
            "golden_check_batch0": golden_check(H, W, expected[0], list(range(B))),
            "clocks": clocks,
            "device_map": device_map,
            "host_binding": {**binding, "pinned_pages_on_node": hostmem.node_histogram(h_logits[0].data_ptr(), h_logits[0].numel() * 4)},
            "host_wall_ms": wall * 1e3,
            "batch_latency_ms": {"mean": float(np.mean(total_ms)) if total_ms else None,
                                 "note": "first to last kernel of one batch, CUDA events, while other batches overlap"},
            "stage_ms": {**stage_ms, "sum": serial,
                         "note": "each stage of one 5-frame batch by itself (library stage timers, eager launches, single stream, "
                                 "nothing else on the GPU): the reference's tic/toc table, semantic_depth.py:445-454"},
            "roofline": {"kernel": "sd::knn_kernel<11> + its heavy-query pass sd::knn_heavy_kernel<11> (statistical_outlier_removal; "
                                   "dominant kernel of the step, profiles/r2_launches_summary.txt)",
                         "bound": "issue", "achieved": knn_ach, "peak": peak_gbs, "unit": "GB/s",
                         "frac": (knn_ach / peak_gbs) if knn_ach else None, "traffic": traffic_of("knn_kernel", "knn_heavy_kernel"),
                         "traffic_source": prof_file, "algorithmic_bytes_per_launch": knn_bytes, "kernel_ms": k_ms,
                         "peak_source": peak_src, "issue_roofline": issue,
                         "note": "exact k-NN on an L1/L2-resident cloud: bound by instruction issue under divergent per-query trip "
                                 "counts and by cache latency, not by HBM (DESIGN.md 3, 4a); achieved/peak/frac are the algorithmic "
                                 "bytes of SURVEY 8d over the kernel's own duration (stage timer, kernel alone on the GPU) against "
                                 "the measured HBM peak; issue_roofline is the instruction-side figure"},
            "roofline_pixel": {"kernel": "sd::pixel_label_kernel + pixel_scan_kernel + pixel_scatter_kernel", "bound": "hbm",
                               "achieved": pix_ach, "peak": peak_gbs, "unit": "GB/s",
                               "frac": (pix_ach / peak_gbs) if pix_ach else None,
                               "traffic": traffic_of("pixel_label_kernel", "pixel_scatter_kernel"), "traffic_source": prof_file,
                               "algorithmic_bytes_per_launch": pix_bytes, "kernel_ms": pix_ms,
                               "note": "pixel stage (3 kernels) as one segment, alone on the GPU"},
            "roofline_path": {"bound": "hbm", "algorithmic_bytes_per_frame": alg_path,
                              "achieved": alg_path * value / world / 1e9, "peak": peak_gbs, "unit": "GB/s",
                              "frac": alg_path * value / world / 1e9 / peak_gbs,
                              "note": "SURVEY.md 8d B_alg x frames/s per GPU; k-NN / radius search are issue / L2 / fp64 bound, not HBM"},
        }
        if e2e:
            ke, ems = e2e
            line["e2e"] = {"value": ke * B * world / (ems * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * HW * 20,
                           "d2h_bytes_per_step": B * 328, "steps": ke,
                           "h2d_gb_per_s_per_gpu": ke * B * HW * 20 / (ems * 1e-3) / 1e9,
                           "h2d_gb_per_s_aggregate": world * ke * B * HW * 20 / (ems * 1e-3) / 1e9,
                           "h2d_ceiling": H2D_CEILING,
                           "api": "FramePipeline.submit_host (pinned host inputs, copies pipelined over the slots)"}
        if e2e_sc:
            ke, ems = e2e_sc
            sc_bytes = B * ((H // 8) * (W // 8) * 12 + HW * 8)
            line["e2e_score_map_mode"] = {"value": ke * B * world / (ems * 1e-3), "unit": UNIT, "h2d_bytes_per_step": sc_bytes,
                                          "d2h_bytes_per_step": B * 328, "steps": ke,
                                          "h2d_gb_per_s_per_gpu": ke * sc_bytes / (ems * 1e-3) / 1e9,
                                          "h2d_gb_per_s_aggregate": world * ke * sc_bytes / (ems * 1e-3) / 1e9,
                                          "note": "same call with FCN-8s' last transposed convolution fused into the label kernel "
                                                  "(SURVEY 8a row 1u): scores [H/8,W/8,3] + disparities cross PCIe, logits never exist"}
        if world == 1 and not a.skip_configs:
            try:
                line.update(secondary_configs(dev))
            except Exception as e:
                line["config4"] = {"value": None, "error": repr(e)}
        if world == 1 and not a.skip_cpu_baseline:
            try:
                from oracle.cpu_baseline import CpuBaseline
                cb = CpuBaseline(H, W, procs=a.cpu_procs or None)
                wall_s, n, answers = cb.step()
                cb.close()
                line["cpu_baseline"] = {"value": n / wall_s, "unit": UNIT, "cores": cb.cores_used, "kind": "port",
                                        "sample": cb.describe(), "host_cores": os.cpu_count(), "wall_s": wall_s}
                if cb.kd_workers == 1:      # one frame on one core (timed inside its worker, all cores busy)
                    per = sorted(t for t, _, _ in answers)
                    line["cpu_baseline"]["single_core"] = {"value": 1.0 / per[len(per) // 2], "unit": UNIT,
                                                           "seconds_per_frame_median": per[len(per) // 2]}
            except Exception as e:   # the baseline is a reported number, never a reason to lose the bench line
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    pipe.close()
    if world > 1:
        dist.destroy_process_group()


def run_stream(a, rank, world, dev, dev_index, device_map, P, intr, barrier, max_over_ranks):
    """BASELINE.json configs[2]: ONE stream of `--stream` frames (frame i = seed i), sharded over the ranks in contiguous
    chunks (stream.shard_frames; the reference's loop is semantic_depth_cityscapes_sequence.py:689-701), every rank runs its
    chunk through its own pipeline, and the 24-byte answers are gathered in frame order with one all_gather.  Strong
    scaling: the total work is fixed.  The gather is inside the timed region."""
    import hashlib
    import numpy as np
    import torch
    import torch.distributed as dist
    from concurrent.futures import ThreadPoolExecutor

    from semantic_depth_b200 import scene
    from semantic_depth_b200.stream import FramePipeline, gather_results, pack_answers, shard_frames

    H, W, B, HW, F = a.height, a.width, a.frames, a.height * a.width, a.stream
    mine = shard_frames(F, rank, world)
    starts = list(range(mine.start, mine.stop, B))
    # inputs of this rank's chunk, resident in HBM (generated a batch at a time on the host)
    d_logits, d_disp = [], []

    def gen(s0):
        nfr = min(B, mine.stop - s0)
        return scene.make_batch(nfr, H, W, first_seed=s0, intr=intr)[:2]
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for lg, dp in ex.map(gen, starts):
            d_logits.append(torch.from_numpy(lg).to(dev)); d_disp.append(torch.from_numpy(dp).to(dev))
    pipe = FramePipeline(H, W, B, slots=a.slots, device=dev, params=P, use_graphs=not a.no_graph)
    if d_logits:                                            # warm: job tables + one graph per slot and batch shape
        for _ in range(max(a.warmup, 1) * len(pipe.slots)):
            pipe.submit_device_stream(d_logits[0], d_disp[0], intr, tag=0)
        if d_logits[-1].shape[0] != B:
            for _ in range(len(pipe.slots)):
                pipe.submit_device_stream(d_logits[-1], d_disp[-1], intr, tag=0)
        pipe.drain()
    main = torch.cuda.current_stream()
    gather_results(pack_answers(np.zeros(len(mine)), np.zeros(len(mine)), np.zeros(len(mine))).to(dev), F, device=dev)   # warm the collective
    sampler = ClockSampler(nvml_index(dev_index))
    sampler.start()
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(main)
    for s in pipe.slots:
        s.stream.wait_event(t0)
    results = {}
    for k in range(len(starts)):
        fin = pipe.submit_device_stream(d_logits[k], d_disp[k], intr, tag=k)     # every batch lives at its own address
        if fin:
            results[fin[0]] = fin[1]
    for tag, res in pipe.drain():
        results[tag] = res
    rw = np.concatenate([results[k].rw for k in range(len(starts))]) if starts else np.zeros(0)
    f2f = np.concatenate([results[k].f2f for k in range(len(starts))]) if starts else np.zeros(0)
    status = np.concatenate([results[k].status for k in range(len(starts))]) if starts else np.zeros(0)
    answers = gather_results(pack_answers(rw, f2f, status).to(dev), F, device=dev).cpu()   # [F,3] in frame order, on every rank
    for s in pipe.slots:
        main.wait_stream(s.stream)
    t1.record(main)
    barrier()
    elapsed_ms = max_over_ranks(t0.elapsed_time(t1))
    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    if rank == 0:
        ans = answers.numpy()
        digest = hashlib.sha256(np.ascontiguousarray(ans).tobytes()).hexdigest()
        gold = []
        for f in range(min(F, 5)):
            path = os.path.join(ROOT, "tests", "golden", f"frame_{H}x{W}_seed{f}.npz")
            if os.path.exists(path):
                g = np.load(path)
                gold.append(bool(float(g["rw"]) == ans[f, 0] and abs(float(g["f2f"]) - ans[f, 1]) <= 1e-3 and int(g["status"]) == int(ans[f, 2])))
        line = {
            "metric": METRIC, "value": F / (elapsed_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": (F + B - 1) // B,
            "warmup": a.warmup, "ms_per_step": elapsed_ms / max(1, (F + B - 1) // B), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 clouds / f64 reprojection, plane and k-NN arithmetic", "data": "synthetic",
            "config": {"workload": f"one stream of {F} synthetic Cityscapes-shaped frames at {H}x{W}, frame i = seed i "
                                   f"(BASELINE.json configs[2]), contiguous shards over {world} rank(s), answers all_gathered in frame order",
                       "height": H, "width": W, "frames": F, "frames_per_batch": B, "batches_in_flight": a.slots,
                       "l2_policy": f"every frame is read once: {F * HW * 20 / 1e9:.1f} GB of inputs resident in HBM across the ranks",
                       "parallelism": f"frame-parallel x{world}; one end-of-run all_gather of {F} x 24 B (NCCL) inside the timed region"},
            "impl": "b200", "elapsed_ms": elapsed_ms,
            "gpu_launches": int(pipe.slots[0].engine.kernel_count(P)) * len(starts),
            "answers_sha256": digest, "answers_head": ans[:3].tolist(),
            "golden_frames_equal": gold, "frames_with_status": int((ans[:, 2] != 0).sum()),
            "rw_mean": float(np.nanmean(ans[:, 0])), "f2f_mean": float(np.nanmean(ans[:, 1])),
            "clocks": sampler.summary(), "device_map": device_map,
        }
        print(json.dumps(line), flush=True)
    pipe.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
