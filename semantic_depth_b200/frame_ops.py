"""Frame-level entry points of the facade: the pixel stage pieces and the fused path.

These are the inline NumPy / OpenCV steps of ``FrameProcessor.process_frame`` and its helpers
(/root/reference/semantic_depth.py:550-564, 656-664, 686-697, 183-324) as single calls.  NumPy in ->
NumPy out, CUDA tensors in -> CUDA tensors out.
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import FusionEngine
from .params import FusionParams, Intrinsics

_frame_engines: dict = {}


def frame_engine(height: int, width: int, max_frames: int = 1, device="cuda:0", max_hypotheses: int = 0) -> FusionEngine:
    key = (height, width, str(device))
    eng = _frame_engines.get(key)
    if eng is None or eng.max_frames < max_frames or eng.max_hyp < max_hypotheses:
        if eng is not None:
            eng.close()
        eng = FusionEngine(height, width, max_frames=max(max_frames, 1), max_hypotheses=max_hypotheses, device=device)
        _frame_engines[key] = eng
    return eng


def _dev(x, dtype=torch.float32):
    if isinstance(x, torch.Tensor):
        return x.to(device="cuda", dtype=dtype).contiguous(), True
    return torch.from_numpy(np.ascontiguousarray(x)).to(device="cuda", dtype=dtype).contiguous(), False


def _back(t, was_torch):
    return t if was_torch else t.cpu().numpy()


def labels_from_logits(logits, shape, prob_thr: float = 0.5):
    """(road_mask, fence_mask) bool [H,W]: ``softmax(logits)[:, c] > prob_thr`` for c = 0, 1
    (semantic_depth.py:555-556,563-564), decided as an fp64 softmax would."""
    h, w = shape
    lg, was_torch = _dev(logits)
    lg = lg.reshape(1, h * w, 3)
    eng = frame_engine(h, w)
    disp = torch.ones((1, 2, h, w), dtype=torch.float32, device=lg.device)
    out = eng.pixel_stage(lg, disp, Intrinsics.synthetic(w), prob_thr=prob_thr)
    lab = out["labels"].reshape(h, w)
    return _back((lab & 1) != 0, was_torch), _back((lab & 2) != 0, was_torch)


def resize_cubic(frame, size):
    """``cv2.resize(frame, size, interpolation=cv2.INTER_CUBIC)`` for uint8 frames [H,W,C] or batches [B,H,W,C]
    (semantic_depth.py:110-112); ``size`` = (width, height) as in OpenCV.  OpenCV's fixed-point definition."""
    from . import _lib
    from ._lib import check
    was_torch = isinstance(frame, torch.Tensor)
    t = (frame if was_torch else torch.from_numpy(np.ascontiguousarray(frame))).to(device="cuda", dtype=torch.uint8).contiguous()
    squeeze_c = t.ndim == 2
    if squeeze_c:
        t = t[:, :, None]
    batched = t.ndim == 4
    if not batched:
        t = t[None]
    b, h, w, c = t.shape
    dw, dh = int(size[0]), int(size[1])
    out = torch.empty((b, dh, dw, c), dtype=torch.uint8, device=t.device)
    with torch.cuda.device(t.device):
        check(_lib.load().sd_resize_cubic_u8(t.data_ptr(), b, h, w, c, out.data_ptr(), dh, dw,
                                             torch.cuda.current_stream().cuda_stream), "sd_resize_cubic_u8")
    out = out if batched else out[0]
    if squeeze_c:
        out = out[..., 0]
    return _back(out, was_torch)


ROAD_RGBA = (128, 64, 128, 64)      # semantic_depth.py:556
FENCE_RGBA = (160, 10, 10, 64)      # semantic_depth.py:564


def _overlay_labels(frames_u8, labels_u8, road_rgba, fence_rgba):
    """frames [B,H,W,3] uint8 and label bytes [B,H*W] (bit0 road, bit1 fence), both on the device -> overlaid frames."""
    import ctypes as C
    from . import _lib
    from ._lib import check
    b, h, w, _ = frames_u8.shape
    out = torch.empty_like(frames_u8)
    scratch = torch.empty(2 * b, dtype=torch.int32, device=frames_u8.device)
    rc_, fc_ = (C.c_int32 * 4)(*[int(v) for v in road_rgba]), (C.c_int32 * 4)(*[int(v) for v in fence_rgba])
    with torch.cuda.device(frames_u8.device):
        check(_lib.load().sd_overlay_masks(frames_u8.data_ptr(), labels_u8.data_ptr(), b, h, w, rc_, fc_, out.data_ptr(),
                                           scratch.data_ptr(), torch.cuda.current_stream().cuda_stream), "sd_overlay_masks")
    return out


def overlay_masks(frame, road_mask, fence_mask, road_rgba=ROAD_RGBA, fence_rgba=FENCE_RGBA):
    """The overlaid frame of ``SegmentFrame.segment_frame`` (semantic_depth.py:547-568): the road mask, then the fence
    mask, pasted as ``toimage(np.dot(mask, [rgba]), mode='RGBA')`` layers with themselves as the paste mask (scipy
    1.2.1 bytescale + PIL paste arithmetic).  frame [H,W,3] or [B,H,W,3] uint8; masks bool [H,W] / [H,W,1] (or batched)."""
    was_torch = isinstance(frame, torch.Tensor)
    t = (frame if was_torch else torch.from_numpy(np.ascontiguousarray(frame)))
    if t.dtype != torch.uint8 or t.ndim not in (3, 4) or t.shape[-1] != 3:
        raise ValueError("frame must be uint8 [H,W,3] or [B,H,W,3]")
    t = t.to(device="cuda").contiguous()
    batched = t.ndim == 4
    if not batched:
        t = t[None]
    b, h, w, _ = t.shape
    rm = _dev(road_mask, torch.bool)[0].reshape(b, h * w)
    fm = _dev(fence_mask, torch.bool)[0].reshape(b, h * w)
    labels = (rm.to(torch.uint8) | (fm.to(torch.uint8) << 1)).contiguous()
    out = _overlay_labels(t, labels, road_rgba, fence_rgba)
    return _back(out if batched else out[0], was_torch)


def segment_frame(frame, logits, prob_thr: float = 0.5):
    """What ``SegmentFrame.segment_frame`` returns after the session run (semantic_depth.py:553-570), from the
    network's logits [H*W,3]: ``(segmentation_road [H,W,1] bool, segmentation_fence [H,W,1] bool, overlaid [H,W,3] uint8)``.
    The label bytes of the pixel stage feed the overlay kernel directly."""
    was_torch = isinstance(frame, torch.Tensor)
    t = (frame if was_torch else torch.from_numpy(np.ascontiguousarray(frame)))
    if t.dtype != torch.uint8 or t.ndim != 3 or t.shape[-1] != 3:
        raise ValueError("frame must be uint8 [H,W,3]")
    t = t.to(device="cuda").contiguous()
    h, w, _ = t.shape
    lg = _dev(logits)[0].reshape(1, h * w, 3)
    eng = frame_engine(h, w)
    disp = torch.ones((1, 2, h, w), dtype=torch.float32, device=lg.device)
    lab = eng.pixel_stage(lg, disp, Intrinsics.synthetic(w), prob_thr=prob_thr)["labels"].reshape(1, h * w).contiguous()
    out = _overlay_labels(t[None], lab, ROAD_RGBA, FENCE_RGBA)[0]
    lab = lab.reshape(h, w, 1)
    return _back((lab & 1) != 0, was_torch), _back((lab & 2) != 0, was_torch), _back(out, was_torch)


def draw_banner(frames, rects=(), texts=()):
    """``cv2.rectangle(img, pt1, pt2, color, -1)`` / ``cv2.putText(img, text, org, 16, font_scale, color, thickness)`` on the
    GPU for the reference's font presets; see ``semantic_depth_b200.banner.draw_banner``.  Returns the frames."""
    from . import banner
    return banner.draw_banner(frames, rects, texts)[0]


def result_banner(segmented_frame, depth, left_pt_rw, right_pt_rw, dist_rw, left_pt_f2f=None, right_pt_f2f=None, dist_f2f=None,
                  is_city=True, approach="both", original_size=None):
    """Section 9 of ``process_frame`` (semantic_depth.py:339-394): the segmented frame up-sampled to the original size with
    ``cv2.resize(..., INTER_CUBIC)`` (when ``original_size`` = (width, height) is given), the grey banner over its top 20 % and
    the result lines.  NumPy in -> NumPy out, CUDA tensor in -> CUDA tensor out."""
    from . import banner
    was_torch = isinstance(segmented_frame, torch.Tensor)
    t = segmented_frame if was_torch else torch.from_numpy(np.ascontiguousarray(segmented_frame))
    t = t.to("cuda")
    t = resize_cubic(t, original_size) if original_size is not None else t.clone()
    h, w = int(t.shape[0]), int(t.shape[1])
    rects, texts = banner.result_banner_spec(h, w, depth, left_pt_rw, right_pt_rw, dist_rw, left_pt_f2f, right_pt_f2f, dist_f2f,
                                             is_city=is_city, approach=approach)
    out, _ = banner.draw_banner(t.contiguous(), rects, texts)
    return _back(out, was_torch)


def sequence_banner(segmented_frame, depth, line_found, left_pt_rw=None, right_pt_rw=None, dist_rw=None, original_size=None):
    """Section 9 of the sequence driver (semantic_depth_cityscapes_sequence.py:304-327): up-sampling, banner over the top 25 %
    and three result lines, or the green "Cannot compute ..." line when no road point fell into the slab."""
    from . import banner
    was_torch = isinstance(segmented_frame, torch.Tensor)
    t = segmented_frame if was_torch else torch.from_numpy(np.ascontiguousarray(segmented_frame))
    t = t.to("cuda")
    t = resize_cubic(t, original_size) if original_size is not None else t.clone()
    h, w = int(t.shape[0]), int(t.shape[1])
    rects, texts = banner.sequence_banner_spec(h, w, depth, bool(line_found), left_pt_rw, right_pt_rw, dist_rw)
    out, _ = banner.draw_banner(t.contiguous(), rects, texts)
    return _back(out, was_torch)


def upsample_scores(scores, weights, bias):
    """FCN-8s' last layer, ``conv2d_transpose(second_skip, 3, 16x16, stride 8, 'same')`` (fcn8s/fcn.py:207-213):
    scores [h,w,3] -> logits [8h*8w, 3] fp32, evaluated by the label kernel (fp32, fixed summation order)."""
    sc, was_torch = _dev(scores)
    h, w = sc.shape[0] * 8, sc.shape[1] * 8
    eng = frame_engine(h, w)
    disp = torch.ones((1, 2, h, w), dtype=torch.float32, device=sc.device)
    out = eng.pixel_stage(None, disp, Intrinsics.synthetic(w), scores=(sc.reshape(1, h // 8, w // 8, 3), _dev(weights)[0], _dev(bias)[0]))
    return _back(out["logits"].reshape(h * w, 3), was_torch)


def post_process_disparity(disp):
    """DepthFrame.post_processing + the fp32 cast (semantic_depth.py:656-664,676): [2,H,W] -> [H,W]."""
    d, was_torch = _dev(disp)
    _, h, w = d.shape
    eng = frame_engine(h, w)
    lg = torch.zeros((1, h * w, 3), dtype=torch.float32, device=d.device)
    out = eng.pixel_stage(lg, d.reshape(1, 2, h, w), Intrinsics.synthetic(w))
    return _back(out["disp_pp"].reshape(h, w), was_torch)


def reproject_to_3d(disparity, intr: Intrinsics):
    """DepthFrame.compute_3D_points (semantic_depth.py:686-697): cv2.reprojectImageTo3D(disp, Q) for the
    fp32 pixel disparity of :145 -> [H,W,3] fp32."""
    d, was_torch = _dev(disparity)
    h, w = d.shape
    eng = frame_engine(h, w)
    lg = torch.zeros((1, h * w, 3), dtype=torch.float32, device=d.device)
    pair = torch.stack([d, d], dim=0).reshape(1, 2, h, w).contiguous()
    out = eng.pixel_stage(lg, pair, intr, raw_disparity=True)
    return _back(out["points"].reshape(h, w, 3), was_torch)


def fuse_frames_from_scores(scores, weights, bias, disp, intrinsics: Intrinsics, params: FusionParams | None = None):
    """``fuse_frames`` fed by the unexpanded FCN-8s head: ``scores`` [B,H/8,W/8,3] (``second_skip``), ``weights``
    [16,16,3,3] and ``bias`` [3] of the last transposed convolution (fcn8s/fcn.py:207-213), ``disp`` [B,2,H,W]."""
    b, _, h, w = disp.shape
    sc, _ = _dev(scores)
    eng = frame_engine(h, w, max_frames=b, device=sc.device)
    return eng.fuse_frames_scores(sc, _dev(weights)[0], _dev(bias)[0], _dev(disp)[0], intrinsics, params)


def fuse_frames(logits, disp, intrinsics: Intrinsics, params: FusionParams | None = None, ransac_hypotheses=None):
    """The fusion section of process_frame (semantic_depth.py:183-324) for a batch of frames.

    ``logits`` [B,H*W,3], ``disp`` [B,2,H,W] (NumPy / CPU tensors -> host entry point with H2D copies;
    CUDA tensors are consumed in place).  Returns a ``FusionResult`` (rw[B], f2f[B], status[B], xl/xr,
    plane coefficients, per-stage counts)."""
    b, _, h, w = disp.shape
    n_hyp = 0
    if ransac_hypotheses:
        n_hyp = max(int(v.shape[1]) for v in ransac_hypotheses.values() if v is not None)
    dev = logits.device if isinstance(logits, torch.Tensor) and logits.is_cuda else "cuda:0"
    eng = frame_engine(h, w, max_frames=b, device=dev, max_hypotheses=n_hyp)
    return eng.fuse_frames(logits, disp, intrinsics, params, ransac_hypotheses)
