"""FusionEngine: host-side driver of the CUDA fusion path.

PyTorch is plumbing only here (device memory, streams, CUDA graphs).  All arithmetic happens in
libsd_fusion.so, reached through the C ABI of include/sd_fusion.h.

The engine mirrors the role of ``FrameProcessor.process_frame`` in the reference
(/root/reference/semantic_depth.py:98-460, fusion section 183-324): ``fuse_frames`` takes the two
networks' raw outputs for a batch of frames and returns rw / f2f and every per-stage observable.
The per-call methods (``median_mad``, ``filter`` ...) are what ``semantic_depth_lib.pcl`` is built on.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from ._lib import SdCamera, SdFrameResult, SdParams, SdPredicate, check
from .params import FusionParams, Intrinsics


def _on_device(fn):
    """Run a per-call op with the engine's device current, so that the stream handed to the C ABI belongs to it."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **kw):
        with torch.cuda.device(self.device):
            return fn(self, *a, **kw)
    return wrapped


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def camera_struct(intr: Intrinsics) -> SdCamera:
    q = intr.as_q32()
    return SdCamera(float(q[0]), float(q[1]), float(q[2]), float(q[3]), float(np.float32(intr.disparity_mult)))


def params_struct(p: FusionParams) -> SdParams:
    s = SdParams()
    lo, hi = p.slab_bounds()
    s.prob_thr = p.prob_thr
    s.road_z_to_meter = p.road_z_to_meter
    s.road_mad_y_thr, s.road_mad_x_thr = p.road_mad_y_thr, p.road_mad_x_thr
    s.fence_mad_y_thr, s.fence_abs_z_thr = p.fence_mad_y_thr, p.fence_abs_z_thr
    s.left_mad_x_thr, s.right_mad_x_thr = p.left_mad_x_thr, p.right_mad_x_thr
    s.sor_nb_neighbors = p.sor_nb_neighbors
    s.road_plane_thr, s.fence_plane_thr = p.road_plane_thr, p.fence_plane_thr
    s.sor_std_ratio, s.ror_radius = p.sor_std_ratio, p.ror_radius
    s.slab_lo, s.slab_hi = lo, hi
    s.depth = p.depth
    s.ror_nb_points = p.ror_nb_points
    s.use_sor, s.use_ror = int(p.use_sor), int(p.use_ror)
    s.approach_both = int(p.approach == "both")
    if p.label_mode not in ("softmax", "argmax"):
        raise ValueError("label_mode must be 'softmax' or 'argmax'")
    s.label_mode = int(p.label_mode == "argmax")
    return s


RESULT_DTYPE = np.dtype([
    ("rw", "<f8"), ("f2f", "<f8"), ("xl", "<f8"), ("xr", "<f8"), ("left_pt", "<f8", 3), ("right_pt", "<f8", 3),
    ("road_coeff", "<f8", 4), ("left_coeff", "<f8", 4), ("right_coeff", "<f8", 4),
    ("sor_mean", "<f8"), ("sor_std", "<f8"), ("sor_thr", "<f8"),
    ("median", "<f4", 5), ("mad", "<f4", 5), ("fence_mean_x", "<f4"), ("status", "<u4"),
    ("counts", "<i4", _lib.SD_NUM_COUNTS), ("ransac_best", "<i4", 3),
], align=True)
assert RESULT_DTYPE.itemsize == C.sizeof(SdFrameResult), (RESULT_DTYPE.itemsize, C.sizeof(SdFrameResult))


@dataclass
class FusionResult:
    """Per-batch answers; every field is a NumPy array with one row per frame."""
    raw: np.ndarray                      # structured array, RESULT_DTYPE
    rw: np.ndarray = field(init=False)
    f2f: np.ndarray = field(init=False)
    status: np.ndarray = field(init=False)

    def __post_init__(self):
        self.rw, self.f2f, self.status = self.raw["rw"], self.raw["f2f"], self.raw["status"]

    def counts(self, frame: int = 0) -> dict:
        return {n: int(c) for n, c in zip(_lib.COUNT_NAMES, self.raw["counts"][frame])}

    def __len__(self):
        return self.raw.shape[0]


class FusionEngine:
    """One workspace on one GPU.  Not thread-safe; use one engine per stream of frames."""

    def __init__(self, height: int, width: int, max_frames: int = 1, max_hypotheses: int = 0,
                 device: str | torch.device = "cuda:0"):
        if not torch.cuda.is_available():
            raise _lib.SdError("FusionEngine needs a CUDA device: the fusion path has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.height, self.width, self.max_frames, self.max_hyp = height, width, max_frames, max_hypotheses
        self.hw = height * width
        with torch.cuda.device(self.device):
            nbytes = self.lib.sd_ws_bytes(max_frames, height, width, max_hypotheses)
            if nbytes == 0:
                raise _lib.SdError("invalid workspace shape (width must be a positive multiple of 4)")
            self._mem = torch.empty(nbytes + 512, dtype=torch.uint8, device=self.device)
            base = self._mem.data_ptr()
            aligned = (base + 255) // 256 * 256
            self._ws = C.c_void_p()
            check(self.lib.sd_ws_create(C.byref(self._ws), C.c_void_p(aligned), nbytes, max_frames, height, width,
                                        max_hypotheses, _stream_ptr()), "sd_ws_create")
            self._results = torch.zeros(max_frames * C.sizeof(SdFrameResult), dtype=torch.uint8, device=self.device)
            self._results_host = torch.zeros(max_frames * C.sizeof(SdFrameResult), dtype=torch.uint8).pin_memory()
        self._stage_logits = None
        self._stage_disp = None
        self.capacity = (self.hw + 4095) // 4096 * 4096

    def close(self):
        if getattr(self, "_ws", None):
            self.lib.sd_ws_destroy(self._ws)
            self._ws = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    # fused per-frame path
    # ------------------------------------------------------------------------------------------
    def _check_inputs(self, logits, disp):
        b = logits.shape[0]
        if tuple(logits.shape) != (b, self.hw, 3) or tuple(disp.shape) != (b, 2, self.height, self.width):
            raise ValueError(f"expected logits [B,{self.hw},3] and disp [B,2,{self.height},{self.width}], got "
                             f"{tuple(logits.shape)} and {tuple(disp.shape)}")
        if b < 1 or b > self.max_frames:
            raise ValueError(f"batch {b} outside [1, {self.max_frames}]")
        return b

    def enqueue(self, logits: torch.Tensor, disp: torch.Tensor, intr: Intrinsics, params: FusionParams | None = None,
                hypotheses: dict | None = None) -> int:
        """Enqueue the fused path on the current stream (no host sync; CUDA-graph capturable after one
        eager call with the same arguments).  Inputs are CUDA fp32 tensors; results stay on the device
        until ``fetch``.  Returns the batch size."""
        params = params or FusionParams()
        for t in (logits, disp):
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise TypeError("enqueue needs contiguous CUDA float32 tensors")
        b = self._check_inputs(logits, disp)
        cam, ps = camera_struct(intr), params_struct(params)
        hr = hl = hg = None
        n_hyp = 0
        if hypotheses:
            hr, hl, hg = (hypotheses.get(k) for k in ("road", "left", "right"))
            for h in (hr, hl, hg):
                if h is not None:
                    if not (h.is_cuda and h.dtype == torch.int32 and h.is_contiguous() and h.shape[0] == b and h.shape[2] == 3):
                        raise TypeError("hypotheses must be contiguous CUDA int32 [B,K,3]")
                    n_hyp = h.shape[1]
            self._hyp_keepalive = (hr, hl, hg)
        # through the TORCH_LIBRARY op layer (csrc/sd_torch_ops.cpp): device / dtype / contiguity are checked again in C++,
        # the stream is the one PyTorch uses on the tensors' device
        _lib.load_ops().fuse_frames(logits, disp, _lib.struct_tensor(cam), _lib.struct_tensor(ps), hr, hl, hg,
                                    int(self._ws.value), self._results)
        return b

    def fetch(self, batch: int) -> FusionResult:
        """Device -> host copy of the last enqueued batch's results (synchronises the current stream)."""
        nbytes = batch * C.sizeof(SdFrameResult)
        self._results_host[:nbytes].copy_(self._results[:nbytes], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        raw = np.frombuffer(self._results_host[:nbytes].numpy().tobytes(), dtype=RESULT_DTYPE).copy()
        return FusionResult(raw)

    def fuse_frames(self, logits, disp, intr: Intrinsics, params: FusionParams | None = None,
                    hypotheses: dict | None = None) -> FusionResult:
        """rw / f2f for a batch of frames.

        ``logits`` [B,H*W,3] fp32 (FCN-8s ``logits:0``), ``disp`` [B,2,H,W] fp32 (monodepth
        ``disp_left_est[0]`` for (frame, fliplr(frame))).  CUDA tensors are consumed in place; NumPy
        arrays / CPU tensors go through the host entry point (H2D copy, fused path, D2H of the results).
        """
        params = params or FusionParams()
        if isinstance(logits, torch.Tensor) and logits.is_cuda:
            with torch.cuda.device(self.device):
                b = self.enqueue(logits, disp, intr, params, hypotheses)
                return self.fetch(b)
        if hypotheses:
            raise NotImplementedError("RANSAC hypotheses need device inputs")
        lg = logits.numpy() if isinstance(logits, torch.Tensor) else np.ascontiguousarray(logits, dtype=np.float32)
        dp = disp.numpy() if isinstance(disp, torch.Tensor) else np.ascontiguousarray(disp, dtype=np.float32)
        b = self._check_inputs(lg, dp)
        with torch.cuda.device(self.device):
            if self._stage_logits is None:
                self._stage_logits = torch.empty((self.max_frames, self.hw, 3), dtype=torch.float32, device=self.device)
                self._stage_disp = torch.empty((self.max_frames, 2, self.height, self.width), dtype=torch.float32,
                                               device=self.device)
            cam, ps = camera_struct(intr), params_struct(params)
            check(self.lib.sd_fuse_frames_host(C.c_void_p(lg.ctypes.data), C.c_void_p(dp.ctypes.data), b, self.height,
                                               self.width, C.byref(cam), C.byref(ps), _ptr(self._stage_logits),
                                               _ptr(self._stage_disp), _ptr(self._results),
                                               C.c_void_p(self._results_host.data_ptr()), self._ws, _stream_ptr()),
                  "sd_fuse_frames_host")
        nbytes = b * C.sizeof(SdFrameResult)
        raw = np.frombuffer(self._results_host[:nbytes].numpy().tobytes(), dtype=RESULT_DTYPE).copy()
        return FusionResult(raw)

    def kernel_count(self, params: FusionParams | None = None, with_ransac: bool = False) -> int:
        ps = params_struct(params or FusionParams())
        return int(self.lib.sd_fuse_kernel_count(C.byref(ps), int(with_ransac)))

    def enable_timing(self, enable: bool = True):
        """Record CUDA events around the pixel-stage kernel and the whole fused call (see sd_fusion.h)."""
        check(self.lib.sd_ws_enable_timing(self._ws, int(enable)), "sd_ws_enable_timing")

    def set_stage_mask(self, mask: int = 15):
        """Bit 0 = pixel stage, bit 1 = cloud stages up to the search grid + fence chain, bit 2 = the k-NN
        kernel, bit 3 = radius search .. answers; 15 = the whole path (default)."""
        check(self.lib.sd_ws_set_stage_mask(self._ws, int(mask)), "sd_ws_set_stage_mask")

    def stage_ms(self, which: str = "pixel") -> float:
        ms = C.c_float(0)
        check(self.lib.sd_ws_stage_elapsed_ms(self._ws, {"pixel": 0, "total": 1}[which], C.byref(ms)), "sd_ws_stage_elapsed_ms")
        return float(ms.value)

    def stage_times(self) -> dict:
        """Device time in ms of every stage of the last ``fuse_frames`` call (after ``enable_timing(True)``): the
        reference's tic / toc pairs of process_frame (semantic_depth.py:157-332, dumped at :445-454).  In this
        profiling mode the fence chain runs on the same stream as the road chain, so the stages add up."""
        ms = (C.c_float * len(_lib.STAGE_NAMES))()
        check(self.lib.sd_ws_stage_times(self._ws, ms), "sd_ws_stage_times")
        return {n: float(v) for n, v in zip(_lib.STAGE_NAMES, ms)}

    def final_cloud(self, frame: int, which: str = "road"):
        """(points [N,3] fp32, src [N] int32) of a frame's final road / left / right cloud (device)."""
        idx = {"road": 0, "left": 1, "right": 2}[which]
        px, py, pz, ps, pn = (C.c_void_p() for _ in range(5))
        check(self.lib.sd_ws_cloud(self._ws, frame, idx, C.byref(px), C.byref(py), C.byref(pz), C.byref(ps), C.byref(pn)),
              "sd_ws_cloud")
        n = int(self._view(pn.value, 1, torch.int32).item())
        x, y, z = (self._view(p.value, n, torch.float32) for p in (px, py, pz))
        return torch.stack([x, y, z], dim=1), self._view(ps.value, n, torch.int32).clone()

    def stage_src(self, frame: int, stage: str, n: int) -> torch.Tensor:
        """Source pixel indices of a retained intermediate stage (parity tests)."""
        p = C.c_void_p()
        rc = self.lib.sd_ws_stage_src(self._ws, frame, _lib.COUNT_NAMES.index(stage), C.byref(p))
        if rc == 0:
            return self._view(p.value, n, torch.int32).clone()
        # a filter that the fused path does not materialise (alive bytes over the chain's input): select the rows here
        ps, pa, pr = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(self.lib.sd_ws_stage_alive(self._ws, frame, _lib.COUNT_NAMES.index(stage), C.byref(ps), C.byref(pa), C.byref(pr)),
              "sd_ws_stage_alive")
        rows = int(self._view(pr.value, 1, torch.int32).item())
        src = self._view(ps.value, rows, torch.int32)
        alive = self._view(pa.value, rows, torch.uint8)
        out = src[alive != 0].clone()
        if out.numel() != n:
            raise _lib.SdError(f"stage {stage}: {out.numel()} alive rows, expected {n}")
        return out

    def _view(self, ptr: int, n: int, dtype) -> torch.Tensor:
        """Tensor view of `n` elements at raw device address `ptr` inside the workspace."""
        off = ptr - self._mem.data_ptr()
        item = torch.empty((), dtype=dtype).element_size()
        assert 0 <= off and off + n * item <= self._mem.numel()
        return self._mem[off:off + n * item].view(dtype)

    # ------------------------------------------------------------------------------------------
    # pixel stage alone
    # ------------------------------------------------------------------------------------------
    def pixel_stage(self, logits: torch.Tensor | None, disp: torch.Tensor, intr: Intrinsics, prob_thr: float = 0.5,
                    road_z_to_meter: float = 7.0, want_dense: bool = True, raw_disparity: bool = False,
                    scores: tuple | None = None, argmax: bool = False) -> dict:
        """Pixel stage alone.  ``scores`` = (scores [B,H/8,W/8,3], weights [16,16,3,3], bias [3]) selects the
        score-map mode (FCN-8s head evaluated in the kernel; ``logits`` is ignored and the upsampled logits are
        returned as ``out['logits']``)."""
        if scores is not None:
            sc, upw, upb = (t.contiguous() for t in scores)
            b = self._check_scores(sc, upw, upb, disp)
        else:
            b = self._check_inputs(logits, disp)
        dev, hw = self.device, self.hw
        f32 = dict(dtype=torch.float32, device=dev)
        out = {k: torch.empty((b, hw), **f32) for k in ("road_x", "road_y", "road_z", "fence_x", "fence_y", "fence_z")}
        out["road_src"] = torch.empty((b, hw), dtype=torch.int32, device=dev)
        out["fence_src"] = torch.empty((b, hw), dtype=torch.int32, device=dev)
        counts = torch.zeros((b, 3), dtype=torch.int32, device=dev)
        labels = torch.empty((b, hw), dtype=torch.uint8, device=dev) if want_dense else None
        points = torch.empty((b, hw, 3), **f32) if want_dense else None
        disp_pp = torch.empty((b, hw), **f32) if want_dense else None
        cam = camera_struct(intr)
        flags = (1 if raw_disparity else 0) | (2 if argmax else 0)
        clouds = (_ptr(out["road_x"]), _ptr(out["road_y"]), _ptr(out["road_z"]), _ptr(out["road_src"]),
                  _ptr(out["fence_x"]), _ptr(out["fence_y"]), _ptr(out["fence_z"]), _ptr(out["fence_src"]))
        logits_out = None
        with torch.cuda.device(dev):
            if scores is not None:
                logits_out = torch.empty((b, hw, 3), **f32) if want_dense else None
                check(self.lib.sd_pixel_fuse_scores(_ptr(sc), _ptr(upw), _ptr(upb), _ptr(disp), b, self.height, self.width,
                                                    C.byref(cam), prob_thr, road_z_to_meter, flags, *clouds,
                                                    _ptr(counts), _ptr(labels), _ptr(points), _ptr(disp_pp), _ptr(logits_out),
                                                    self._ws, _stream_ptr()), "sd_pixel_fuse_scores")
            else:
                check(self.lib.sd_pixel_fuse(_ptr(logits), _ptr(disp), None, None, b, self.height, self.width, C.byref(cam),
                                             prob_thr, road_z_to_meter, flags, *clouds,
                                             _ptr(counts), _ptr(labels), _ptr(points), _ptr(disp_pp), self._ws, _stream_ptr()),
                      "sd_pixel_fuse")
            torch.cuda.current_stream().synchronize()
        out.update(counts=counts.cpu().numpy(), labels=labels, points=points, disp_pp=disp_pp, logits=logits_out)
        return out

    def _check_scores(self, sc, upw, upb, disp) -> int:
        for t in (sc, upw, upb, disp):
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise TypeError("score-map mode needs contiguous CUDA float32 tensors")
        if self.height % 8 or self.width % 8:
            raise ValueError("score-map mode needs frame sizes that are multiples of 8")
        b = disp.shape[0]
        if tuple(disp.shape) != (b, 2, self.height, self.width) or tuple(sc.shape) != (b, self.height // 8, self.width // 8, 3):
            raise ValueError(f"expected disp [B,2,{self.height},{self.width}] and scores [B,{self.height // 8},{self.width // 8},3]")
        if tuple(upw.shape) != (16, 16, 3, 3) or tuple(upb.shape) != (3,):
            raise ValueError("expected up-sampling weights [16,16,3,3] (kh, kw, out, in) and bias [3]")
        if b < 1 or b > self.max_frames:
            raise ValueError(f"batch {b} exceeds the engine's max_frames={self.max_frames}")
        return b

    def fuse_frames_scores(self, scores: torch.Tensor, weights: torch.Tensor, bias: torch.Tensor, disp: torch.Tensor,
                           intr: Intrinsics, params: FusionParams | None = None) -> FusionResult:
        """The fused path fed by the unexpanded FCN-8s head (``second_skip`` scores + the transposed-conv kernel of
        fcn8s/fcn.py:207-213): the 12 B/pixel logits tensor is never materialised."""
        with torch.cuda.device(self.device):
            return self.fetch(self.enqueue_scores(scores, weights, bias, disp, intr, params))

    def enqueue_scores(self, scores, weights, bias, disp, intr: Intrinsics, params: FusionParams | None = None) -> int:
        """``enqueue`` for the score-map mode (no host sync; CUDA-graph capturable after one eager call)."""
        params = params or FusionParams()
        sc, upw, upb = scores.contiguous(), weights.contiguous(), bias.contiguous()
        b = self._check_scores(sc, upw, upb, disp)
        cam, ps = camera_struct(intr), params_struct(params)
        _lib.load_ops().fuse_frames_scores(sc, upw, upb, disp, _lib.struct_tensor(cam), _lib.struct_tensor(ps),
                                           int(self._ws.value), self._results)
        return b

    # ------------------------------------------------------------------------------------------
    # per-call cloud ops (SoA device tensors in, device tensors / host scalars out)
    # ------------------------------------------------------------------------------------------
    @_on_device
    def median_mad(self, col: torch.Tensor) -> tuple[np.float32, np.float32]:
        out = (C.c_float * 2)()
        check(self.lib.sd_median_mad(_ptr(col), col.numel(), out, self._ws, _stream_ptr()), "sd_median_mad")
        return np.float32(out[0]), np.float32(out[1])

    @_on_device
    def filter(self, x, y, z, pred: SdPredicate, want_points: bool = True):
        """Stable filter; returns (kept_idx int32 [M], (x,y,z) of the survivors or None)."""
        n = x.numel()
        dev = x.device
        idx = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        ox = oy = oz = None
        if want_points:
            ox, oy, oz = (torch.empty(max(n, 1), dtype=torch.float32, device=dev) for _ in range(3))
        n_out = C.c_int32(0)
        check(self.lib.sd_filter(_ptr(x), _ptr(y), _ptr(z), None, n, C.byref(pred), _ptr(ox), _ptr(oy), _ptr(oz),
                                 _ptr(idx), C.byref(n_out), self._ws, _stream_ptr()), "sd_filter")
        m = n_out.value
        pts = (ox[:m], oy[:m], oz[:m]) if want_points else None
        return idx[:m], pts

    @_on_device
    def plane_fit(self, x, y, z, axis: int):
        coeff = (C.c_double * 3)()
        sing = C.c_int32(0)
        check(self.lib.sd_plane_fit(_ptr(x), _ptr(y), _ptr(z), x.numel(), axis, coeff, C.byref(sing), self._ws,
                                    _stream_ptr()), "sd_plane_fit")
        return np.array(list(coeff), dtype=np.float64), bool(sing.value)

    @_on_device
    def mean_f32(self, col: torch.Tensor) -> np.float32:
        out = C.c_float(0)
        check(self.lib.sd_mean_f32(_ptr(col), col.numel(), C.byref(out), self._ws, _stream_ptr()), "sd_mean_f32")
        return np.float32(out.value)

    @_on_device
    def slab_minmax(self, x, z, lo: float, hi: float, use_f32: bool):
        xmin, xmax, cnt = C.c_float(0), C.c_float(0), C.c_int32(0)
        check(self.lib.sd_slab_minmax(_ptr(x), _ptr(z), x.numel(), lo, hi, int(use_f32), C.byref(xmin), C.byref(xmax),
                                      C.byref(cnt), self._ws, _stream_ptr()), "sd_slab_minmax")
        return np.float32(xmin.value), np.float32(xmax.value), int(cnt.value)

    @_on_device
    def knn_mean_distance(self, x, y, z, k: int, std_ratio: float = 0.5):
        n = x.numel()
        avg = torch.empty(max(n, 1), dtype=torch.float64, device=x.device)
        stats = (C.c_double * 3)()
        check(self.lib.sd_knn_mean_distance(_ptr(x), _ptr(y), _ptr(z), n, k, std_ratio, _ptr(avg), stats, self._ws,
                                            _stream_ptr()), "sd_knn_mean_distance")
        return avg[:n], (stats[0], stats[1], stats[2])

    @_on_device
    def radius_count(self, x, y, z, radius: float, cap: int = -1):
        n = x.numel()
        cnt = torch.empty(max(n, 1), dtype=torch.int32, device=x.device)
        check(self.lib.sd_radius_count(_ptr(x), _ptr(y), _ptr(z), n, radius, cap, _ptr(cnt), self._ws, _stream_ptr()),
              "sd_radius_count")
        return cnt[:n]

    @_on_device
    def ransac_score(self, x, y, z, axis: int, threshold: float, triplets: torch.Tensor):
        k = triplets.shape[0]
        counts = torch.empty(k, dtype=torch.int32, device=x.device)
        best = C.c_int32(-1)
        coeff = (C.c_double * 3)()
        check(self.lib.sd_ransac_score(_ptr(x), _ptr(y), _ptr(z), x.numel(), axis, threshold, _ptr(triplets), k,
                                       _ptr(counts), C.byref(best), coeff, self._ws, _stream_ptr()), "sd_ransac_score")
        return counts, int(best.value), np.array(list(coeff), dtype=np.float64)
