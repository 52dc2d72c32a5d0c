"""A ``FrameProcessor``-shaped driver over the fused path (SURVEY.md 8b "Who calls it").

The reference's ``FrameProcessor`` (/root/reference/semantic_depth.py:81-97, live twin
semantic_depth_cityscapes_sequence.py:101-115) is built from two injected producers, a network input shape, the
approach ('rw' | 'both') and the depth, and its ``process_frame`` walks one frame through
resize -> segmentation -> disparity -> 3D points -> filters -> rw / f2f -> overlay -> PLY.  This class keeps that
shape.  The producers stay injected objects; what changes is what they hand over, because everything after the two
networks runs on the GPU in one fused call:

* ``frame_segmenter.logits(frame) -> [H*W, 3] fp32``       the ``logits:0`` tensor ``segment_frame`` feeds to softmax
                                                            (semantic_depth.py:550-552, fcn8s/fcn.py:241)
* ``frame_depther.disparities(frame) -> [2, H, W] fp32``   ``disp_left_est[0]`` of (frame, fliplr(frame)), i.e. the
                                                            input of ``post_processing`` (semantic_depth.py:667-676)

(NumPy arrays or CUDA tensors.)  Steps and the reference lines they replace:

    cv2.resize(..., INTER_CUBIC)                       :110-112   resize_cubic_u8 kernel
    segment_frame: masks + overlaid frame              :547-570   label kernel + overlay kernels
    post_processing, * disparity_mult, reprojectImageTo3D, mask gather, every pcl.* filter, Open3D SOR / ROR,
    end points, plane intersection                     :145-324   one fused call (49 kernels)
    cv2.resize of the segmented frame back             :sequence 304   resize_cubic_u8 kernel
    PointCloud2Ply(road3D, road_colors) + rw line      :sequence 372-376   ply kernels

    cv2.rectangle / cv2.putText of the banner         :339-394, sequence 306-327   banner kernels (``banner=...``)

cv2.imwrite stays with the caller.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import frame_ops
from .params import FusionParams, Intrinsics, STATUS_NO_SLAB_POINTS, status_to_names
from . import pcl_gpu


@dataclass
class FrameOutput:
    """Everything ``process_frame`` computes for one frame (host values; clouds stay on the device)."""
    dist_rw: float | None                 # abs(left_pt_rw[0][0] - right_pt_rw[0][0])    semantic_depth.py:259
    dist_f2f: float | None                # compute_distance_in_3D(left_pt_f2f, right_pt_f2f)      :324
    line_found: bool                      # sequence:232-237
    left_pt_rw: np.ndarray | None         # rows of the road cloud with the smallest / largest x in the slab
    right_pt_rw: np.ndarray | None
    left_pt_f2f: np.ndarray | None        # (1, 3) plane intersections at z = -depth               :317-323
    right_pt_f2f: np.ndarray | None
    road_mask: np.ndarray                 # [H, W] bool
    fence_mask: np.ndarray
    segmented_frame: np.ndarray           # overlay, resized back to the original frame size, uint8
    road3D: torch.Tensor                  # final road cloud [N, 3] fp32 (device) and its colours [N, 3] uint8
    road_colors: torch.Tensor
    status: int = 0
    status_names: list = field(default_factory=list)
    counts: dict = field(default_factory=dict)
    ply_path: str | None = None


class FrameProcessor:
    #: scale of the normalised disparity; None = the original frame width (semantic_depth.py:109), the sequence
    #: driver hard-codes 3800 (sequence:105)
    disp_multiplier = None

    def __init__(self, frame_segmenter, frame_depther, input_shape, approach="both", depth=10.0, verbose=False,
                 intrinsics: Intrinsics | None = None, disp_multiplier: float | None = None,
                 params: FusionParams | None = None, banner: str | None = None, is_city: bool = True):
        if approach not in ("rw", "both"):
            raise ValueError("approach must be 'rw' or 'both'")
        for obj, name in ((frame_segmenter, "logits"), (frame_depther, "disparities")):
            if not callable(getattr(obj, name, None)):
                raise TypeError(f"{type(obj).__name__} must provide {name}(frame)")
        self.frame_segmenter = frame_segmenter
        self.frame_depther = frame_depther
        self.input_shape = (int(input_shape[0]), int(input_shape[1]))
        self.approach = approach
        self.depth = float(depth)
        self.verbose = verbose
        self.intrinsics = intrinsics
        if disp_multiplier is not None:
            self.disp_multiplier = float(disp_multiplier)
        self.params = (params or FusionParams()).replace(depth=self.depth, approach=approach)
        if banner not in (None, "single", "sequence"):
            raise ValueError("banner must be None, 'single' (semantic_depth.py:339-394) or 'sequence' (sequence:304-327)")
        self.banner = banner
        self.is_city = bool(is_city)

    def _intrinsics(self, original_width: int) -> Intrinsics:
        mult = float(original_width) if self.disp_multiplier is None else float(self.disp_multiplier)
        base = self.intrinsics or Intrinsics.cityscapes(self.input_shape[1])
        return Intrinsics(cx=base.cx, cy=base.cy, f=base.f, b=base.b, disparity_mult=mult)

    def process_frame(self, original_frame, output_name: str | None = None, result_ply_dir: str | None = None) -> FrameOutput:
        """``original_frame``: the BGR uint8 image ``cv2.imread`` returned ([H0, W0, 3], NumPy or CUDA tensor)."""
        h, w = self.input_shape
        t = original_frame if isinstance(original_frame, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(original_frame))
        if t.dtype != torch.uint8 or t.ndim != 3 or t.shape[2] != 3:
            raise ValueError("original_frame must be uint8 [H, W, 3]")
        t = t.to("cuda").contiguous()
        h0, w0 = int(t.shape[0]), int(t.shape[1])
        frame = frame_ops.resize_cubic(t, (w, h))                                             # :110-112

        logits = frame_ops._dev(self.frame_segmenter.logits(frame))[0].reshape(1, h * w, 3)
        disp = frame_ops._dev(self.frame_depther.disparities(frame))[0].reshape(1, 2, h, w)
        intr = self._intrinsics(w0)

        road_mask, fence_mask, overlaid = frame_ops.segment_frame(frame, logits[0], self.params.prob_thr)   # :547-570
        res = frame_ops.fuse_frames(logits, disp, intr, self.params)                          # :145-324
        status = int(res.status[0])
        eng = frame_ops.frame_engine(h, w, device=logits.device)
        road3D, src = eng.final_cloud(0, "road")
        colors = torch.flip(frame, dims=(2,)).reshape(h * w, 3)                               # cv2.COLOR_BGR2RGB, :162
        road_colors = colors[src.long()]

        line_found = not (status & STATUS_NO_SLAB_POINTS) and road3D.shape[0] > 0
        left_rw = right_rw = None
        dist_rw = None
        if line_found:
            # float64 like the reference's cloud after the Open3D round trip (:244): the slab bounds stay Python doubles
            left_rw, right_rw = pcl_gpu.get_end_points_of_road(road3D.to(torch.float64), self.depth - 0.02)    # :254-255
            line_found = left_rw is not None
        if line_found:
            left_rw, right_rw = left_rw.cpu().numpy(), right_rw.cpu().numpy()
            dist_rw = float(res.rw[0])
            assert dist_rw == abs(float(left_rw[0][0]) - float(right_rw[0][0]))              # :259, same points
        dist_f2f = left_f2f = right_f2f = None
        if self.approach == "both" and np.isfinite(res.f2f[0]):
            dist_f2f = float(res.f2f[0])
            left_f2f, right_f2f = res.raw["left_pt"][0][None, :].copy(), res.raw["right_pt"][0][None, :].copy()

        segmented = frame_ops.resize_cubic(overlaid, (w0, h0))                                # sequence:304
        if self.banner == "sequence":                                                         # sequence:306-327
            segmented = frame_ops.sequence_banner(segmented, self.depth, line_found, left_rw, right_rw, dist_rw)
        elif self.banner == "single" and line_found and (self.approach == "rw" or dist_f2f is not None):   # :346-394
            segmented = frame_ops.result_banner(segmented, self.depth, left_rw, right_rw, dist_rw, left_f2f, right_f2f, dist_f2f,
                                                is_city=self.is_city, approach=self.approach)
        out = FrameOutput(dist_rw, dist_f2f, line_found, left_rw, right_rw, left_f2f, right_f2f,
                          road_mask[..., 0].cpu().numpy(), fence_mask[..., 0].cpu().numpy(), segmented.cpu().numpy(),
                          road3D, road_colors, status, status_to_names(status), res.counts(0))
        if self.verbose:
            print(f"Road width {dist_rw}  fence to fence {dist_f2f}  status {out.status_names}")

        if result_ply_dir is not None:                                                        # sequence:372-376
            from semantic_depth_lib.point_cloud_2_ply import PointCloud2Ply
            name = "{}/{}_rw".format(result_ply_dir, output_name or "frame")
            cloud = PointCloud2Ply(road3D.cpu().numpy().astype(np.float64), road_colors.cpu().numpy().astype(np.float64), name)
            if line_found:
                line_rw, colors_line_rw = pcl_gpu.create_3Dline_from_3Dpoints(left_rw[:1].astype(np.float64),
                                                                              right_rw[:1].astype(np.float64), [250, 0, 0])
                line_rw[:, 2] += 0.2                                                          # :265
                cloud.add_extra_point_cloud(line_rw, colors_line_rw)
            cloud.prepare_and_save_point_cloud()
            out.ply_path = name + ".ply"
        return out
