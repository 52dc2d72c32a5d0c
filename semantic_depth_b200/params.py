"""Constants of the fusion path, gathered in one place.

The reference hard-codes every threshold as a literal at its call site
(/root/reference/semantic_depth.py:206-219, 234-239, 255, 279-309).  The same literals are the
defaults here so that ``FusionParams()`` reproduces the reference's per-frame pipeline.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Intrinsics:
    """Pinhole camera used by the disparity -> 3D step (semantic_depth.py:686-697).

    The reference builds ``Q`` as a *float32* matrix, so the values OpenCV sees are the float32
    roundings of ``-cx, cy, -f, 1/b``.  ``as_q32`` returns exactly those four numbers.
    ``disparity_mult`` is the scale applied to the normalised disparity (semantic_depth.py:109,145:
    the original image width; the sequence driver uses the constant 3800, sequence:105,146).
    """
    cx: float
    cy: float
    f: float
    b: float
    disparity_mult: float

    def as_q32(self) -> np.ndarray:
        """float32 [q03, q13, q23, q32] = [-cx, cy, -f, 1/b] as stored in the reference's Q."""
        return np.array([-self.cx, self.cy, -self.f, 1 / self.b], dtype=np.float32)

    @staticmethod
    def cityscapes(width: int, f: float | None = None, b: float = 0.6,
                   disparity_mult: float | None = None) -> "Intrinsics":
        """Cityscapes preset (semantic_depth.py:592-599) scaled from the 512-wide network frame.

        The reference's preset is for a 256x512 frame (cx = 1048.64/4, cy = 519.277/4, f = 500);
        larger frames scale cx, cy, f by width/512.
        """
        s = width / 512.0
        return Intrinsics(cx=1048.64 / 4 * s, cy=519.277 / 4 * s,
                          f=(500.0 if f is None else float(f)) * s, b=b,
                          disparity_mult=float(width if disparity_mult is None else disparity_mult))

    @staticmethod
    def munich(f: float = 380.0) -> "Intrinsics":
        """iPhone preset of the Munich test set (semantic_depth.py:600-607), 256x512 frame."""
        return Intrinsics(cx=314.05519001, cy=124.09658151, f=float(f), b=1.0, disparity_mult=512.0)

    @staticmethod
    def synthetic(width: int) -> "Intrinsics":
        """Intrinsics of the synthetic straight-road scene."""
        s = width / 2048.0
        return Intrinsics(cx=1048.64 * s, cy=519.277 * s, f=2000.0 * s, b=0.6,
                          disparity_mult=float(width))


@dataclass(frozen=True)
class FusionParams:
    """Every literal of FrameProcessor.process_frame's fusion section, with its call site."""
    # road chain -----------------------------------------------------------------------------
    road_z_to_meter: float = 7.0        # remove_from_to(road, 2, 0.0, 7.0)        semantic_depth.py:206
    road_mad_y_thr: float = 15.0        # remove_noise_by_mad(road, 1, 15.0)       :209
    road_mad_x_thr: float = 2.0         # remove_noise_by_mad(road, 0, 2.0)        :212
    road_plane_thr: float = 5.0         # remove_noise_by_fitting_plane(axis=1)    :215-219
    sor_nb_neighbors: int = 10          # statistical_outlier_removal              :234-235
    sor_std_ratio: float = 0.5
    ror_nb_points: int = 80             # radius_outlier_removal                   :238-239
    ror_radius: float = 0.5
    use_sor: bool = True                # thesis-era path had neither filter
    use_ror: bool = True
    # answers ----------------------------------------------------------------------------------
    depth: float = 10.0                 # --depth default                          :736-738
    rw_depth_offset: float = 0.02       # get_end_points_of_road(road, depth-0.02) :254-255
    slab_half_width: float = 0.05       # pcl.py:283
    approach: str = "both"              # 'rw' skips the fence chain               :273
    # fence chain ------------------------------------------------------------------------------
    fence_mad_y_thr: float = 5.0        # :279
    fence_abs_z_thr: float = 35.0       # threshold_complete(fence, 2, 35.0)       :283-284
    left_mad_x_thr: float = 5.0         # :291
    right_mad_x_thr: float = 1.0        # :302
    fence_plane_thr: float = 1.0        # :294-298, 305-309
    # labels -----------------------------------------------------------------------------------
    prob_thr: float = 0.5               # softmax > 0.5                            :555-556,563-564
    label_mode: str = "softmax"         # "softmax" (the reference) | "argmax" (north_star's wording of the labelling)

    def slab_bounds(self) -> tuple[float, float]:
        """(lo, hi) with lo < z < hi, computed exactly as the reference's Python doubles do.

        pcl.py:283 evaluates ``-(depth+0.05)`` and ``-(depth-0.05)`` where ``depth`` is the
        already-offset value ``self.depth-0.02`` (semantic_depth.py:254-255).
        """
        d = self.depth - self.rw_depth_offset
        return -(d + self.slab_half_width), -(d - self.slab_half_width)

    def replace(self, **kw) -> "FusionParams":
        return dataclasses.replace(self, **kw)


# Per-frame status bits of the fused path (SURVEY.md section 8b "Error conventions").
STATUS_OK = 0
STATUS_EMPTY_ROAD = 1 << 0          # a road-chain stage received/produced an empty cloud
STATUS_EMPTY_FENCE_LEFT = 1 << 1
STATUS_EMPTY_FENCE_RIGHT = 1 << 2
STATUS_MAD_ZERO = 1 << 3            # some MAD was 0 or NaN (everything dropped, pcl.py:63-67)
STATUS_NO_SLAB_POINTS = 1 << 4      # get_end_points_of_segment -> (None, None)  pcl.py:303-304
STATUS_SINGULAR_PLANES = 1 << 5     # 2x2 system singular (np.linalg.inv raises) pcl.py:232
STATUS_EMPTY_FENCE = 1 << 6
STATUS_SINGULAR_FIT = 1 << 7        # 3x3 normal equations singular (degenerate cloud)

STATUS_NAMES = {
    STATUS_EMPTY_ROAD: "EMPTY_ROAD", STATUS_EMPTY_FENCE_LEFT: "EMPTY_FENCE_LEFT",
    STATUS_EMPTY_FENCE_RIGHT: "EMPTY_FENCE_RIGHT", STATUS_MAD_ZERO: "MAD_ZERO",
    STATUS_NO_SLAB_POINTS: "NO_SLAB_POINTS", STATUS_SINGULAR_PLANES: "SINGULAR_PLANES",
    STATUS_EMPTY_FENCE: "EMPTY_FENCE", STATUS_SINGULAR_FIT: "SINGULAR_FIT",
}


def status_to_names(status: int) -> list[str]:
    return [n for bit, n in STATUS_NAMES.items() if status & bit]
