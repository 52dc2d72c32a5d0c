"""CPU oracle (TEST INFRASTRUCTURE ONLY) of the result banner of process_frame: cv2.rectangle + cv2.putText.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this module.

Reference: /root/reference/semantic_depth.py:346-394 and semantic_depth_cityscapes_sequence.py:306-327

    cv2.rectangle(self.segmented_frame, (0, 0), (w, int(0.2*h)), (156, 157, 159), -1)
    cv2.putText(self.segmented_frame, 'At {:.2f}m depth:'.format(self.depth), (int(middle*w), int(h_zero)),
                fontFace=16, fontScale=fontScale, color=(255, 255, 255), thickness=thickness)
    ...

Third-party arithmetic: OpenCV's drawing.cpp (opencv-contrib-python 4.0.0.21 in requirements.txt:21; 4.13 in this
container).  Restated here from its published algorithm as far as the banner needs it:

* ``rectangle(..., thickness=-1)`` with integer corners fills the inclusive box between them, clipped to the image;
* ``putText``: ``hscale = cvRound(fontScale * 65536)``; the pen starts at ``org.x << 16`` and every character first moves it
  by ``-left * hscale``, emits its stroke vertices at ``pen + x * hscale`` and then moves it by ``right * hscale``; so the
  raster of a character relative to the pen's integer pixel depends on the character and on ``frac(pen)`` only.  The stroke
  rasters themselves (ThickLine -> FillConvexPoly + Circle on Hershey glyph tables that live inside the OpenCV binary) are
  DATA here: the bitmaps OpenCV itself rendered for the reference's presets, baked by
  semantic_depth_b200/data/make_hershey_atlas.py into hershey_atlas.npz.

**Pinned**: tests/golden/make_golden_banner.py executes the reference's own banner statements (lifted with ``ast`` from both
drivers) with the real cv2 and asserts byte equality with this module; the atlas generator's --verify does the same for
random strings.  Exactness domain: glyphs fully inside the frame (OpenCV clips stroke segments, not pixels).
"""
from __future__ import annotations

import os

import numpy as np

ATLAS_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "semantic_depth_b200", "data", "hershey_atlas.npz")
FIRST, LAST = 32, 126
SLACK = 1 << 11
_atlas = None


def _load():
    global _atlas
    if _atlas is None:
        z = np.load(ATLAS_PATH)
        _atlas = {k: z[k] for k in z.files}
    return _atlas


def cv_round(v: float) -> int:
    return int(np.rint(v))                           # lrint: half to even


def preset_of(font_scale: float, thickness: int) -> int:
    a = _load()
    for i in range(len([k for k in a if k.endswith("_meta")])):
        m = a[f"p{i}_meta"]
        if int(m[1]) == int(round(font_scale * 1000)) and int(m[2]) == int(thickness):
            return i
    raise ValueError("preset not baked")


def rectangle_filled(img: np.ndarray, pt1, pt2, color) -> None:
    """cv2.rectangle(img, pt1, pt2, color, -1): inclusive box, clipped; in place."""
    h, w = img.shape[:2]
    x0, x1 = sorted((int(pt1[0]), int(pt2[0]))); y0, y1 = sorted((int(pt1[1]), int(pt2[1])))
    x0, y0, x1, y1 = max(x0, 0), max(y0, 0), min(x1, w - 1), min(y1, h - 1)
    if x1 >= x0 and y1 >= y0:
        img[y0:y1 + 1, x0:x1 + 1] = np.asarray(color, dtype=img.dtype)


def put_text(img: np.ndarray, text: str, org, font_scale: float, color, thickness: int) -> None:
    """cv2.putText(img, text, org, 16, font_scale, color, thickness) for a baked preset; in place."""
    a = _load()
    p = preset_of(font_scale, thickness)
    meta, units, bits, index = a[f"p{p}_meta"], a[f"p{p}_units"], a[f"p{p}_bits"], a[f"p{p}_index"]
    hscale, period, cell_h, words, x_off, y_off, cum_max = (int(v) for v in meta[3:10])
    assert hscale == cv_round(font_scale * 65536)
    h, w = img.shape[:2]
    col = np.asarray(color, dtype=img.dtype)
    cum = 0
    for ch in text:
        c = ord(ch)
        if c < FIRST or c > LAST:
            c = ord("?")
        g = c - FIRST
        pen = (int(org[0]) << 16) + hscale * cum
        px = (pen + SLACK) >> 16
        cc = cum
        if cc > cum_max:
            cc -= period * ((cc - cum_max + period - 1) // period)
        bm = bits[int(index[g, cc])]
        rows, wds = np.nonzero(bm)
        if len(rows):
            v = bm[rows, wds]
            for b in range(32):
                sel = ((v >> np.uint32(b)) & np.uint32(1)) != 0
                if sel.any():
                    x = px - x_off + wds[sel] * 32 + b
                    y = int(org[1]) - y_off + rows[sel]
                    ok = (x >= 0) & (x < w) & (y >= 0) & (y < h)
                    img[y[ok], x[ok]] = col
        cum += int(units[g])


def draw(img: np.ndarray, rects, texts) -> np.ndarray:
    """rects: (pt1, pt2, color); texts: (text, org, font_scale, thickness, color).  Returns a new image."""
    out = np.array(img, copy=True)
    for p1, p2, color in rects:
        rectangle_filled(out, p1, p2, color)
    for text, org, scale, thick, color in texts:
        put_text(out, text, org, scale, color, thick)
    return out
