"""Frame streams: pipelined batches on one GPU and frame-parallel sharding across GPUs.

The reference processes a folder / video one frame at a time in a Python ``for`` loop
(/root/reference/semantic_depth.py:867-901, semantic_depth_cityscapes_sequence.py:689-701).  Frames
are independent, so here

* ``FramePipeline`` keeps several batches in flight on one GPU (one workspace + stream + CUDA graph
  per slot, round robin), hiding the ~45 dependent kernel boundaries of one batch behind the other
  batches; host inputs are copied H2D on the slot's stream so copies overlap compute;
* ``shard_frames`` / ``gather_results`` split a stream of frames over the ranks of a
  ``torch.distributed`` job (one process per GPU).  There is no collective on the data path: the
  only exchange is one end-of-run all_gather of the 24-byte answers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import SdFrameResult
from .engine import FusionEngine, FusionResult, RESULT_DTYPE, camera_struct
from .params import FusionParams, Intrinsics


# ---------------------------------------------------------------------------------------------
# multi-GPU sharding (host logic; covered by world_size-2 gloo tests on CPU)
# ---------------------------------------------------------------------------------------------
def shard_frames(n_frames: int, rank: int, world: int) -> range:
    """Contiguous, balanced chunk of frame indices for ``rank`` (first ``n % world`` ranks get one more)."""
    if world < 1 or not (0 <= rank < world) or n_frames < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def pack_answers(rw, f2f, status) -> torch.Tensor:
    """[n,3] float64: rw, f2f, status (status is exact in fp64)."""
    out = np.stack([np.asarray(rw, np.float64), np.asarray(f2f, np.float64), np.asarray(status, np.float64)], axis=1)
    return torch.from_numpy(np.ascontiguousarray(out))


def gather_results(local: torch.Tensor, n_frames: int, device=None):
    """All-gather the per-rank [n_local,3] answers into the global [n_frames,3] array (frame order).

    Uses the default process group (nccl on GPUs, gloo in the CPU tests).  Ranks hold the chunks of
    ``shard_frames``; chunks are padded to the largest chunk for the collective."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local.clone()
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [len(shard_frames(n_frames, r, world)) for r in range(world)]
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} frames, expected {sizes[rank]}")
    pad = max(sizes) if sizes else 0
    buf = torch.full((pad, 3), float("nan"), dtype=torch.float64, device=device or local.device)
    buf[: local.shape[0]] = local.to(buf.device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0).cpu()


# ---------------------------------------------------------------------------------------------
# pipelined batches on one GPU
# ---------------------------------------------------------------------------------------------
def _cam_key(intr: Intrinsics) -> bytes:
    """The camera is captured BY VALUE in a slot's CUDA graph (kernel arguments), so it is part of the graph's key: a
    batch with other intrinsics -- the reference derives disparity_mult from each frame's width,
    semantic_depth.py:109,145 -- gets its own capture instead of silently replaying the first camera."""
    return bytes(camera_struct(intr))


class _Slot:
    def __init__(self, height, width, batch, device, max_hyp=0):
        self.engine = FusionEngine(height, width, max_frames=batch, max_hypotheses=max_hyp, device=device)
        self.stream = torch.cuda.Stream(device=device)
        self.done = torch.cuda.Event()
        self.ev = {0: torch.cuda.Event(enable_timing=True), 4: torch.cuda.Event(enable_timing=True)}   # first / last kernel of a batch
        self.graphs = {}           # key -> CUDAGraph
        self.stage_logits = None   # device staging for host inputs
        self.stage_disp = None
        self.host_results = torch.zeros(batch * C.sizeof(SdFrameResult), dtype=torch.uint8).pin_memory()
        self.busy = False
        self.tag = None
        self.nbytes = 0
        self.timed_total = False   # ev[0] / ev[4] bracket the batch (set by _launch; the score-map path records no brackets)


class FramePipeline:
    """Round-robin pipeline of ``slots`` batches in flight on one GPU.

    ``submit_*`` enqueues a batch on the next slot (waiting for that slot's previous batch first and
    returning its result, if any); ``drain`` returns the results still in flight.  Results come back
    in submission order as ``(tag, FusionResult)``.
    """

    def __init__(self, height: int, width: int, batch: int, slots: int = 2, device="cuda:0",
                 params: FusionParams | None = None, use_graphs: bool = True):
        self.height, self.width, self.batch = height, width, batch
        self.device = torch.device(device)
        self.params = params or FusionParams()
        self.use_graphs = use_graphs
        self.slots = [_Slot(height, width, batch, self.device) for _ in range(slots)]
        self.total_ms: list[float] = []     # first to last kernel of every retired batch (per-stage times: FusionEngine.stage_times)
        self._next = 0

    # -- internals ---------------------------------------------------------------------------------
    def _retire(self, slot: _Slot):
        if not slot.busy:
            return None
        slot.done.synchronize()
        if slot.timed_total:
            self.total_ms.append(slot.ev[0].elapsed_time(slot.ev[4]))
        raw = np.frombuffer(slot.host_results[: slot.nbytes].numpy().tobytes(), dtype=RESULT_DTYPE).copy()
        slot.busy = False
        return slot.tag, FusionResult(raw)

    def _capture(self, slot: _Slot, logits, disp, intr, mask):
        eng = slot.engine
        eng.set_stage_mask(mask)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=slot.stream):
            eng.enqueue(logits, disp, intr, self.params)
        eng.set_stage_mask(15)
        return g

    def _launch(self, slot: _Slot, logits: torch.Tensor, disp: torch.Tensor, intr: Intrinsics, key):
        """Enqueue one batch on the slot's stream: one CUDA graph per key (captured after one eager call that builds the job
        tables), bracketed by two events."""
        eng = slot.engine
        b = logits.shape[0]
        slot.ev[0].record(slot.stream)
        if self.use_graphs:
            g = slot.graphs.get(key)
            if g is None:
                eng.enqueue(logits, disp, intr, self.params)          # eager once: builds the job tables
                slot.stream.synchronize()
                g = self._capture(slot, logits, disp, intr, 15)
                slot.graphs[key] = g
            g.replay()
        else:
            eng.enqueue(logits, disp, intr, self.params)
        slot.ev[4].record(slot.stream)
        slot.nbytes = b * C.sizeof(SdFrameResult)
        slot.host_results[: slot.nbytes].copy_(eng._results[: slot.nbytes], non_blocking=True)
        slot.done.record(slot.stream)
        slot.busy = True
        slot.timed_total = True

    def _take_slot(self):
        slot = self.slots[self._next]
        self._next = (self._next + 1) % len(self.slots)
        return slot, self._retire(slot)

    # -- public ------------------------------------------------------------------------------------
    def submit_device(self, logits: torch.Tensor, disp: torch.Tensor, intr: Intrinsics, tag=None):
        """Inputs already resident on the GPU (fixed tensors are replayed through a cached CUDA graph)."""
        slot, finished = self._take_slot()
        with torch.cuda.stream(slot.stream):
            slot.tag = tag
            self._launch(slot, logits, disp, intr, key=(logits.data_ptr(), disp.data_ptr(), logits.shape[0], _cam_key(intr)))
        return finished

    def submit_device_stream(self, logits: torch.Tensor, disp: torch.Tensor, intr: Intrinsics, tag=None):
        """Device inputs that live at a DIFFERENT address every batch (a stream of frames produced on the GPU): the pixel
        stage -- the only kernels that see the input pointers -- is launched eagerly, everything behind it replays one
        CUDA graph per slot and batch size.  No per-batch capture, no staging copy."""
        slot, finished = self._take_slot()
        eng = slot.engine
        b = logits.shape[0]
        with torch.cuda.stream(slot.stream):
            slot.tag = tag
            key = ("rest", b, _cam_key(intr))
            g = slot.graphs.get(key) if self.use_graphs else None
            slot.ev[0].record(slot.stream)
            if self.use_graphs and g is None:
                eng.enqueue(logits, disp, intr, self.params)              # eager once: builds the job tables (and is this batch's run)
                slot.stream.synchronize()
                eng.set_stage_mask(14)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=slot.stream):
                    eng.enqueue(logits, disp, intr, self.params)
                eng.set_stage_mask(15)
                slot.graphs[key] = g
            elif g is not None:
                eng.set_stage_mask(1)
                eng.enqueue(logits, disp, intr, self.params)              # pixel stage, eager: three launches
                eng.set_stage_mask(15)
                g.replay()
            else:
                eng.enqueue(logits, disp, intr, self.params)
            slot.ev[4].record(slot.stream)
            slot.nbytes = b * C.sizeof(SdFrameResult)
            slot.host_results[: slot.nbytes].copy_(eng._results[: slot.nbytes], non_blocking=True)
            slot.done.record(slot.stream)
            slot.busy = True
            slot.timed_total = True
        return finished

    def warm_device(self, logits: torch.Tensor, disp: torch.Tensor, intr: Intrinsics, tag=None):
        """Run one batch through EVERY slot (captures each slot's graph for these tensors); returns the results."""
        out = []
        for _ in range(len(self.slots)):
            fin = self.submit_device(logits, disp, intr, tag)
            if fin:
                out.append(fin)
        return out + self.drain()

    def submit_host(self, logits, disp, intr: Intrinsics, tag=None):
        """Host inputs (NumPy arrays or CPU tensors; pinned memory makes the copy asynchronous)."""
        slot, finished = self._take_slot()
        lg = logits if isinstance(logits, torch.Tensor) else torch.from_numpy(logits)
        dp = disp if isinstance(disp, torch.Tensor) else torch.from_numpy(disp)
        b = lg.shape[0]
        with torch.cuda.stream(slot.stream):
            if slot.stage_logits is None:
                slot.stage_logits = torch.empty((self.batch, self.height * self.width, 3), dtype=torch.float32, device=self.device)
                slot.stage_disp = torch.empty((self.batch, 2, self.height, self.width), dtype=torch.float32, device=self.device)
            slot.stage_logits[:b].copy_(lg, non_blocking=True)
            slot.stage_disp[:b].copy_(dp, non_blocking=True)
            slot.tag = tag
            self._launch(slot, slot.stage_logits[:b], slot.stage_disp[:b], intr, key=("host", b, _cam_key(intr)))
        return finished

    def submit_host_scores(self, scores, weights_dev: torch.Tensor, bias_dev: torch.Tensor, disp, intr: Intrinsics, tag=None):
        """Host inputs in the score-map mode: ``scores`` [B,H/8,W/8,3] and ``disp`` [B,2,H,W] come from (pinned) host
        memory, the up-sampling kernel and bias are device-resident model weights.  One CUDA graph per slot."""
        slot, finished = self._take_slot()
        sc = scores if isinstance(scores, torch.Tensor) else torch.from_numpy(scores)
        dp = disp if isinstance(disp, torch.Tensor) else torch.from_numpy(disp)
        b = sc.shape[0]
        eng = slot.engine
        with torch.cuda.stream(slot.stream):
            if getattr(slot, "stage_scores", None) is None:
                slot.stage_scores = torch.empty((self.batch, self.height // 8, self.width // 8, 3), dtype=torch.float32, device=self.device)
            if slot.stage_disp is None:
                slot.stage_disp = torch.empty((self.batch, 2, self.height, self.width), dtype=torch.float32, device=self.device)
            slot.stage_scores[:b].copy_(sc, non_blocking=True)
            slot.stage_disp[:b].copy_(dp, non_blocking=True)
            slot.tag = tag
            key = ("host_scores", b, weights_dev.data_ptr(), bias_dev.data_ptr(), _cam_key(intr))
            args = (slot.stage_scores[:b], weights_dev, bias_dev, slot.stage_disp[:b], intr, self.params)
            g = slot.graphs.get(key)
            if g is None and self.use_graphs:
                eng.enqueue_scores(*args)                     # eager once: builds the job tables
                slot.stream.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=slot.stream):
                    eng.enqueue_scores(*args)
                slot.graphs[key] = g
            if g is not None:
                g.replay()
            else:
                eng.enqueue_scores(*args)
            slot.nbytes = b * C.sizeof(SdFrameResult)
            slot.host_results[: slot.nbytes].copy_(eng._results[: slot.nbytes], non_blocking=True)
            slot.done.record(slot.stream)
            slot.busy = True
            slot.timed_total = False
        return finished

    def drain(self):
        out = []
        for _ in range(len(self.slots)):
            slot = self.slots[self._next]
            self._next = (self._next + 1) % len(self.slots)
            r = self._retire(slot)
            if r is not None:
                out.append(r)
        return out

    def close(self):
        for s in self.slots:
            s.engine.close()
