"""Deterministic synthetic Cityscapes-shaped input frames.

The reference ships no weights and no test images, and random-init producers give degenerate
clouds, so every test and benchmark feeds the fusion path with an analytic straight-road scene
expressed in the reference's world frame (x right, y up, depth = -z, semantic_depth.py:691-694)
mixed with producer-like noise:

* ground plane y = -1.5 m (label 0 = road),
* left wall x = -4.0 m and right wall x = +3.5 m, labelled fence below y < 1.0 m (label 1),
* far clip 80 m; everything else is background (label 2),
* ``logits = N(0,1) + 4*onehot(label)``  (FCN-8s ``logits:0`` layout [H*W, 3] fp32, fcn.py:241),
* ``disp[0] = scene_disp*(1 + 0.02*eps)``, ``disp[1] = fliplr(disp[0])*1.001``
  (monodepth ``disp_left_est[0]`` for (frame, fliplr(frame)), layout [2, H, W] fp32,
  semantic_depth.py:672-676), with ``scene_disp = f*b/depth/W`` normalised like monodepth's.

Frame ``i`` of a stream uses ``seed = i``.  This is input synthesis, not part of the product path
and not part of the oracle: both consume its output.
"""
from __future__ import annotations

import numpy as np

from .params import Intrinsics

GROUND_Y = -1.5
LEFT_WALL_X = -4.0
RIGHT_WALL_X = 3.5
FENCE_TOP_Y = 1.0
FAR_CLIP = 80.0
LOGIT_MARGIN = 4.0
DISP_NOISE = 0.02
FLIP_GAIN = 1.001


def scene_geometry(height: int, width: int, intr: Intrinsics | None = None):
    """Noise-free label map [H,W] uint8 and depth map [H,W] float64 of the analytic scene."""
    intr = intr or Intrinsics.synthetic(width)
    u = np.arange(width, dtype=np.float64)[None, :] - intr.cx
    v = intr.cy - np.arange(height, dtype=np.float64)[:, None]
    f = intr.f
    inf = np.inf
    with np.errstate(divide="ignore", invalid="ignore"):
        t_ground = np.where(v < 0, GROUND_Y / v, inf) * np.ones_like(u)
        t_left = np.where(u < 0, LEFT_WALL_X / u, inf) * np.ones_like(v)
        t_right = np.where(u > 0, RIGHT_WALL_X / u, inf) * np.ones_like(v)
    t_wall = np.minimum(t_left, t_right)
    y_wall = t_wall * v
    wall_ok = np.isfinite(t_wall) & (y_wall > GROUND_Y)
    t_wall = np.where(wall_ok, t_wall, inf)
    t = np.minimum(t_ground, t_wall)
    depth = t * f
    label = np.full((height, width), 2, dtype=np.uint8)
    is_ground = (t_ground <= t_wall) & np.isfinite(t_ground)
    is_fence = (~is_ground) & np.isfinite(t_wall) & (y_wall < FENCE_TOP_Y)
    label[is_ground] = 0
    label[is_fence] = 1
    far = ~(depth <= FAR_CLIP)
    label[far] = 2
    depth = np.where(far, FAR_CLIP, depth)
    return label, depth


def make_frame(height: int, width: int, seed: int = 0, intr: Intrinsics | None = None):
    """One synthetic frame: (logits [H*W,3] fp32, disp [2,H,W] fp32, intrinsics)."""
    intr = intr or Intrinsics.synthetic(width)
    label, depth = scene_geometry(height, width, intr)
    rng = np.random.default_rng(seed)
    logits = rng.standard_normal((height * width, 3), dtype=np.float32)
    logits[np.arange(height * width), label.reshape(-1)] += np.float32(LOGIT_MARGIN)
    eps = rng.standard_normal((height, width), dtype=np.float32)
    scene_disp = (intr.f * intr.b / depth / width).astype(np.float32)
    left = scene_disp * (np.float32(1.0) + np.float32(DISP_NOISE) * eps)
    right = np.ascontiguousarray(left[:, ::-1]) * np.float32(FLIP_GAIN)
    disp = np.stack([left, right], axis=0).astype(np.float32)
    return np.ascontiguousarray(logits), np.ascontiguousarray(disp), intr


def make_frame_scores(height: int, width: int, seed: int = 0, intr: Intrinsics | None = None):
    """The same scene with the FCN-8s head left unexpanded: (scores [H/8, W/8, 3], weights [16,16,3,3],
    bias [3], disp [2,H,W], intrinsics).  Scores are the scene labels at 1/8 resolution plus N(0,1) noise;
    the transposed-conv kernel is a bilinear interpolation kernel per class plus the reference's
    truncated-normal(0.01) initialisation noise (fcn8s/fcn.py:161,207-213)."""
    if height % 8 or width % 8:
        raise ValueError("the score-map mode needs frame sizes that are multiples of 8")
    intr = intr or Intrinsics.synthetic(width)
    label, depth = scene_geometry(height, width, intr)
    rng = np.random.default_rng(seed)
    h, w = height // 8, width // 8
    low = label[4::8, 4::8]
    scores = rng.standard_normal((h, w, 3), dtype=np.float32)
    scores[np.arange(h)[:, None], np.arange(w)[None, :], low] += np.float32(LOGIT_MARGIN)
    k = np.arange(16, dtype=np.float64)
    tri = 1.0 - np.abs(k - 7.5) / 8.0                         # bilinear kernel of a stride-8 upsample
    weights = np.zeros((16, 16, 3, 3), dtype=np.float32)
    for c in range(3):
        weights[:, :, c, c] = np.outer(tri, tri).astype(np.float32)
    weights += np.clip(rng.standard_normal((16, 16, 3, 3)), -2, 2).astype(np.float32) * np.float32(0.01)
    bias = (rng.standard_normal(3) * 0.01).astype(np.float32)
    eps = rng.standard_normal((height, width), dtype=np.float32)
    scene_disp = (intr.f * intr.b / depth / width).astype(np.float32)
    left = scene_disp * (np.float32(1.0) + np.float32(DISP_NOISE) * eps)
    right = np.ascontiguousarray(left[:, ::-1]) * np.float32(FLIP_GAIN)
    disp = np.stack([left, right], axis=0).astype(np.float32)
    return np.ascontiguousarray(scores), weights, bias, np.ascontiguousarray(disp), intr


def make_batch(n_frames: int, height: int, width: int, first_seed: int = 0,
               intr: Intrinsics | None = None):
    """Batch of frames: logits [B,H*W,3], disp [B,2,H,W]; frame i uses seed first_seed+i."""
    intr = intr or Intrinsics.synthetic(width)
    logits = np.empty((n_frames, height * width, 3), dtype=np.float32)
    disp = np.empty((n_frames, 2, height, width), dtype=np.float32)
    for i in range(n_frames):
        logits[i], disp[i], _ = make_frame(height, width, first_seed + i, intr)
    return logits, disp, intr


def make_road_cloud(n_points: int = 2_000_000, seed: int = 0, outlier_frac: float = 0.02):
    """Config 4 stress cloud: planar road strip plus uniform outliers.

    x ~ U(-4, 3.5), z ~ -U(7, 60), y = -1.5 + N(0, 0.03); ``outlier_frac`` of the points are
    replaced by uniform samples in a 10 m box around (0, -1.5, -30).  Returns [N,3] fp32.
    """
    rng = np.random.default_rng(seed)
    pts = np.empty((n_points, 3), dtype=np.float32)
    pts[:, 0] = rng.uniform(-4.0, 3.5, n_points)
    pts[:, 1] = -1.5 + 0.03 * rng.standard_normal(n_points)
    pts[:, 2] = -rng.uniform(7.0, 60.0, n_points)
    n_out = int(round(outlier_frac * n_points))
    where = rng.choice(n_points, n_out, replace=False)
    pts[where] = (rng.uniform(-5.0, 5.0, (n_out, 3)) + np.array([0.0, -1.5, -30.0])).astype(np.float32)
    return pts
