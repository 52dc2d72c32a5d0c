"""CPU oracle (TEST INFRASTRUCTURE ONLY) of the segmentation overlay of SegmentFrame.segment_frame.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this module.

Reference: /root/reference/semantic_depth.py:547-568 (same at semantic_depth_cityscapes_sequence.py:462-483)

    street_im = scipy.misc.toimage(frame)
    road_mask = np.dot(segmentation_road, np.array([[128, 64, 128, 64]]))        # (H, W, 1) bool . (1, 4) -> int64
    road_mask = scipy.misc.toimage(road_mask, mode="RGBA")
    street_im.paste(road_mask, box=None, mask=road_mask)
    ... same for the fence with [160, 10, 10, 64] ...
    return ..., np.array(street_im)

Two third-party pieces are involved:

* scipy.misc.toimage / bytescale -- scipy 1.2.1 (requirements.txt:30; Pillow>=7.1.0 at requirements.txt:23), removed from SciPy >= 1.3 and therefore absent
  here: **restated** from the published 1.2.1 source (scipy/misc/pilutil.py).  `bytescale` returns uint8 input
  unchanged (so the frame is not touched) and stretches everything else from [min, max] of the WHOLE array to
  [0, 255]: `(clip((data - cmin) * (255 / (cmax - cmin)), 0, 255) + 0.5).astype(uint8)`, cmax == cmin -> scale 1.
  A mask with both set and unset pixels therefore becomes RGBA (255, 128, 255, 128) for the road (min 0, max 128)
  and (255, 16, 16, 102) for the fence (min 0, max 160); an empty mask is all zeros; a mask that covers EVERY pixel
  has min = the smallest colour component, which maps alpha 64 -> 0 (road) / 92 (fence).  Parity unpinned for this
  part (no scipy.misc in the container).
* PIL.Image.paste(im, None, mask) with an RGBA mask on an RGB image -- **pinned** against the Pillow installed in the
  build container by tests/golden/make_golden_overlay.py:
  `out = MULDIV255(dst * (255 - a) + src * a)`, `MULDIV255(t) = ((t + 128 >> 8) + t + 128) >> 8`, a = the mask's
  alpha band, src = the RGB bands of the pasted image.
"""
from __future__ import annotations

import numpy as np

ROAD_RGBA = (128, 64, 128, 64)     # semantic_depth.py:556
FENCE_RGBA = (160, 10, 10, 64)     # semantic_depth.py:564


def bytescale(data: np.ndarray, cmin=None, cmax=None, high: int = 255, low: int = 0) -> np.ndarray:
    """scipy 1.2.1 scipy/misc/pilutil.py:bytescale."""
    data = np.asarray(data)
    if data.dtype == np.uint8:
        return data
    if cmin is None:
        cmin = data.min()
    if cmax is None:
        cmax = data.max()
    cscale = cmax - cmin
    if cscale < 0:
        raise ValueError("`cmax` should be larger than `cmin`.")
    if cscale == 0:
        cscale = 1
    scale = float(high - low) / cscale
    bytedata = (data - cmin) * scale + low
    return (bytedata.clip(low, high) + 0.5).astype(np.uint8)


def mask_rgba(mask: np.ndarray, rgba) -> np.ndarray:
    """`toimage(np.dot(mask[..., None], [[r, g, b, a]]), mode='RGBA')` as an (H, W, 4) uint8 array."""
    m = np.asarray(mask, dtype=bool)
    data = np.dot(m.reshape(m.shape[0], m.shape[1], 1), np.array([list(rgba)]))
    return bytescale(data)


def paste_rgba(dst_rgb: np.ndarray, src_rgba: np.ndarray) -> np.ndarray:
    """Image.paste(src, None, mask=src) on an RGB image (Pillow src/libImaging/Paste.c: paste_mask_RGBA + BLEND8)."""
    a = src_rgba[..., 3:4].astype(np.int32)
    tmp = dst_rgb.astype(np.int32) * (255 - a) + src_rgba[..., :3].astype(np.int32) * a + 128
    return (((tmp >> 8) + tmp) >> 8).astype(np.uint8)


def overlay_masks(frame: np.ndarray, road: np.ndarray, fence: np.ndarray,
                  road_rgba=ROAD_RGBA, fence_rgba=FENCE_RGBA) -> np.ndarray:
    """The third return value of segment_frame: frame (H, W, 3) uint8 with the road, then the fence mask pasted."""
    frame = np.asarray(frame)
    if frame.dtype != np.uint8 or frame.ndim != 3 or frame.shape[2] != 3:
        raise ValueError("frame must be (H, W, 3) uint8")
    out = paste_rgba(frame, mask_rgba(road, road_rgba))
    return paste_rgba(out, mask_rgba(fence, fence_rgba))


def overlay_from_labels(frame: np.ndarray, labels: np.ndarray, road_rgba=ROAD_RGBA, fence_rgba=FENCE_RGBA) -> np.ndarray:
    """labels: (H, W) uint8 with bit 0 = road, bit 1 = fence (the pixel stage's label byte)."""
    labels = np.asarray(labels)
    return overlay_masks(frame, (labels & 1) != 0, (labels & 2) != 0, road_rgba, fence_rgba)
