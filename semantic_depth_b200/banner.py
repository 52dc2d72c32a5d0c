"""The result banner of ``process_frame``: ``cv2.rectangle`` + ``cv2.putText`` on the segmented frame, on the GPU.

Reference: /root/reference/semantic_depth.py:339-394 (single frames: font scale 2 / thickness 2 for Cityscapes frames,
4 / 5 for the Munich set) and semantic_depth_cityscapes_sequence.py:304-327 (scale 2 and 2.2, thickness 2), always
``fontFace=16`` (FONT_HERSHEY_SIMPLEX | FONT_ITALIC), 8-connected lines.

OpenCV rasterises Hershey stroke fonts in 16.16 fixed point; the glyph bitmaps used here were rendered by OpenCV itself
for exactly those (scale, thickness) presets (``data/make_hershey_atlas.py``; one bitmap per character for integer scales,
one per character and pen position for scale 2.2) and ship as ``data/hershey_atlas.npz``.  This module is the host glue:
it formats nothing itself (callers pass the strings), reproduces OpenCV's pen arithmetic -- ``hscale = cvRound(scale *
65536)``, pen += (right - left) * hscale per character -- to pick bitmap and position, and hands the placements to
``sd_draw_banner``.  Byte-identical to cv2 whenever every glyph lies inside the frame; a glyph that crosses the frame border
is clipped pixel-wise here, whereas OpenCV clips its stroke segments before rasterising them, which moves a few pixels of
the clipped strokes (``glyphs_inside`` tells the caller).
"""
from __future__ import annotations

import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
ATLAS_PATH = os.path.join(_HERE, "data", "hershey_atlas.npz")
FONT_FACE = 16
_FIRST, _LAST = 32, 126
_SLACK = 1 << 11

_atlas = None
_dev_bits: dict = {}


class _Preset:
    def __init__(self, meta, units, bits, index):
        (self.face, scale_milli, self.thickness, self.hscale, self.period, self.cell_h, self.words, self.x_off, self.y_off,
         self.cum_max) = (int(v) for v in meta)
        self.scale = scale_milli / 1000.0
        self.units = units.astype(np.int64)
        self.bits = np.ascontiguousarray(bits, dtype=np.uint32)
        self.index = index.astype(np.int64)
        # ink box of every bitmap (x0, y0, x1, y1 inclusive, relative to the cell; empty bitmaps: x1 < x0)
        nb = self.bits.shape[0]
        self.ink = np.zeros((nb, 4), np.int64)
        for k in range(nb):
            rows = np.nonzero(self.bits[k].any(axis=1))[0]
            if len(rows) == 0:
                self.ink[k] = (0, 0, -1, -1)
                continue
            cols = np.bitwise_or.reduce(self.bits[k], axis=0)
            xs = [32 * j + b for j in range(self.words) for b in range(32) if (int(cols[j]) >> b) & 1]
            self.ink[k] = (min(xs), int(rows[0]), max(xs), int(rows[-1]))

    def bitmap(self, g: int, cum: int) -> int:
        if cum > self.cum_max:                      # beyond the baked pen positions: the same phase inside them
            cum -= self.period * ((cum - self.cum_max + self.period - 1) // self.period)
        return int(self.index[g, cum])


def atlas():
    global _atlas
    if _atlas is None:
        if not os.path.exists(ATLAS_PATH):
            raise RuntimeError(f"{ATLAS_PATH} is missing (bake it with data/make_hershey_atlas.py where opencv is installed)")
        z = np.load(ATLAS_PATH)
        n = len([k for k in z.files if k.endswith("_meta")])
        _atlas = [_Preset(z[f"p{i}_meta"], z[f"p{i}_units"], z[f"p{i}_bits"], z[f"p{i}_index"]) for i in range(n)]
    return _atlas


def preset_index(font_scale: float, thickness: int, font_face: int = FONT_FACE) -> int:
    for i, p in enumerate(atlas()):
        if p.face == int(font_face) and abs(p.scale - float(font_scale)) < 1e-9 and p.thickness == int(thickness):
            return i
    have = ", ".join(f"(scale {p.scale:g}, thickness {p.thickness})" for p in atlas())
    raise ValueError(f"no baked glyphs for fontFace {font_face}, fontScale {font_scale}, thickness {thickness}; the atlas holds {have} "
                     "(add the preset to data/make_hershey_atlas.py)")


def layout_text(preset: int, text: str, org) -> list:
    """``[(bitmap, x, y)]`` of every character of ``cv2.putText(img, text, org, 16, scale, ..., thickness)``:
    top-left pixel of the glyph cell and which bitmap of the preset's atlas goes there."""
    p = atlas()[preset]
    ox, oy = int(org[0]), int(org[1])
    out, cum = [], 0
    for ch in text:
        c = ord(ch)
        if c < _FIRST or c > _LAST:                 # cv2's readCheck: anything else is drawn as '?'
            c = ord("?")
        g = c - _FIRST
        pen = (ox << 16) + p.hscale * cum
        px = (pen + _SLACK) >> 16
        out.append((p.bitmap(g, cum), px - p.x_off, oy - p.y_off))
        cum += int(p.units[g])
    return out


def _pack(color) -> int:
    c = [int(round(float(v))) for v in color]
    if len(c) != 3 or any(v < 0 or v > 255 for v in c):
        raise ValueError("colour must be three components in [0, 255]")
    return c[0] | (c[1] << 8) | (c[2] << 16)


def _device_bits(preset: int, device) -> torch.Tensor:
    key = (preset, str(device))
    t = _dev_bits.get(key)
    if t is None:
        t = torch.from_numpy(atlas()[preset].bits.view(np.int32)).to(device).contiguous()
        _dev_bits[key] = t
    return t


def draw_banner(frames, rects=(), texts=()):
    """Draw filled rectangles, then text lines, on uint8 frames ``[H,W,3]`` or ``[B,H,W,3]`` (NumPy -> NumPy copy, CUDA
    tensor -> drawn in place and returned).

    ``rects``: ``(frame, pt1, pt2, color)`` as in ``cv2.rectangle(img, pt1, pt2, color, -1)``;
    ``texts``: ``(frame, text, org, font_scale, thickness, color)`` as in ``cv2.putText(img, text, org, 16, font_scale, color,
    thickness)``.  Returns ``(frames, glyphs_inside)``; ``glyphs_inside`` is False when a glyph cell crosses the frame
    border (see the module docstring)."""
    from . import _lib
    from ._lib import check
    was_torch = isinstance(frames, torch.Tensor)
    t = frames if was_torch else torch.from_numpy(np.ascontiguousarray(frames))
    if t.dtype != torch.uint8 or t.ndim not in (3, 4) or t.shape[-1] != 3:
        raise ValueError("frames must be uint8 [H,W,3] or [B,H,W,3]")
    if not t.is_cuda:
        t = t.to("cuda")
    if not t.is_contiguous():
        raise ValueError("frames must be contiguous (they are drawn in place)")
    batched = t.ndim == 4
    tb = t if batched else t[None]
    b, h, w, _ = tb.shape
    rrows = []
    for f, p1, p2, color in rects:
        if not 0 <= int(f) < b:
            raise ValueError("rectangle refers to a frame outside the batch")
        x0, x1 = sorted((int(p1[0]), int(p2[0]))); y0, y1 = sorted((int(p1[1]), int(p2[1])))
        rrows.append([int(f), x0, y0, x1, y1, _pack(color)])
    groups: list = []                                # runs of consecutive lines with the same preset and colour
    inside = True
    for f, text, org, scale, thick, color in texts:
        if not 0 <= int(f) < b:
            raise ValueError("text refers to a frame outside the batch")
        pi = preset_index(scale, thick)
        col = _pack(color)
        pr = atlas()[pi]
        rows = [[int(f), g, x, y, col] for g, x, y in layout_text(pi, str(text), org)]
        for _, g, x, y, _ in rows:
            ix0, iy0, ix1, iy1 = (int(v) for v in pr.ink[g])
            if ix1 >= ix0:
                inside &= (x + ix0 >= 0 and y + iy0 >= 0 and x + ix1 < w and y + iy1 < h)
        if groups and groups[-1][0] == pi and groups[-1][1] == col:
            groups[-1][2].extend(rows)
        else:
            groups.append([pi, col, rows])
    lib = _lib.load()
    with torch.cuda.device(tb.device):
        st = torch.cuda.current_stream().cuda_stream
        d_rects = torch.tensor(rrows, dtype=torch.int32, device=tb.device).reshape(-1, 6) if rrows else None
        if d_rects is not None and not groups:
            check(lib.sd_draw_banner(tb.data_ptr(), b, h, w, d_rects.data_ptr(), len(rrows), None, 0, 0, 0, None, 0, st), "sd_draw_banner")
        for gi, (pi, _, rows) in enumerate(groups):
            pr = atlas()[pi]
            bits = _device_bits(pi, tb.device)
            d_places = torch.tensor(rows, dtype=torch.int32, device=tb.device).reshape(-1, 5)
            first = gi == 0 and d_rects is not None
            check(lib.sd_draw_banner(tb.data_ptr(), b, h, w, d_rects.data_ptr() if first else None, len(rrows) if first else 0,
                                     bits.data_ptr(), int(pr.bits.shape[0]), pr.cell_h, pr.words, d_places.data_ptr(), len(rows), st),
                  "sd_draw_banner")
    out = t if was_torch else t.cpu().numpy()
    return out, bool(inside)


# ---- the reference's two banners ---------------------------------------------------------------------------------------

BANNER_COLOR = (156, 157, 159)


def result_banner_spec(h: int, w: int, depth: float, left_pt_rw, right_pt_rw, dist_rw, left_pt_f2f=None, right_pt_f2f=None,
                       dist_f2f=None, is_city: bool = True, approach: str = "both", frame: int = 0):
    """Rectangles and text lines of semantic_depth.py:346-394 for a frame of ``h`` x ``w`` pixels (``draw_banner`` input)."""
    if is_city:
        thickness, font_scale, left, right, middle = 2, 2, 0.01, 0.68, 0.33
    else:
        thickness, font_scale, left, right, middle = 5, 4, 0.01, 0.67, 0.33
    h_zero, h_first, h_second = 0.05 * h, 0.12 * h, 0.18 * h
    white = (255, 255, 255)
    rects = [(frame, (0, 0), (w, int(0.2 * h)), BANNER_COLOR)]
    texts = [(frame, "At {:.2f}m depth:".format(depth), (int(middle * w), int(h_zero)), font_scale, thickness, white)]
    if approach == "both":
        texts += [
            (frame, "{:.2f}m to l fence".format(-left_pt_f2f[0][0]), (int(left * w), int(h_first)), font_scale, thickness, white),
            (frame, "{:.2f}m to r fence".format(right_pt_f2f[0][0]), (int(right * w), int(h_first)), font_scale, thickness, white),
            (frame, "Fence2Fence: {:.2f}m".format(dist_f2f), (int(middle * w), int(h_first)), font_scale, thickness, white),
        ]
    texts += [
        (frame, "{:.2f}m to road's l".format(-left_pt_rw[0][0]), (int(left * w), int(h_second)), font_scale, thickness, white),
        (frame, "{:.2f}m to road's r".format(right_pt_rw[0][0]), (int(right * w), int(h_second)), font_scale, thickness, white),
        (frame, "Road's width: {:.2f}m".format(dist_rw), (int(middle * w), int(h_second)), font_scale, thickness, white),
    ]
    return rects, texts


def sequence_banner_spec(h: int, w: int, depth: float, line_found: bool, left_pt_rw=None, right_pt_rw=None, dist_rw=None,
                         frame: int = 0):
    """Rectangles and text lines of semantic_depth_cityscapes_sequence.py:306-327."""
    thickness, font_scale = 2, 2
    white = (255, 255, 255)
    if not line_found:
        return [], [(frame, "Cannot compute width of road at {:.2f} m depth:".format(depth), (int(0.28 * w), int(0.035 * h)),
                     font_scale + 0.2, thickness, (0, 255, 0))]
    rects = [(frame, (0, 0), (w, int(0.25 * h)), BANNER_COLOR)]
    texts = [
        (frame, "At {:.2f} m depth:".format(depth), (int(0.36 * w), int(0.05 * h)), font_scale + 0.2, thickness, white),
        (frame, "{:.2f}m to road's left end".format(-left_pt_rw[0][0]), (int(0.05 * w), int(0.13 * h)), font_scale, thickness, white),
        (frame, "{:.2f}m to road's right end".format(right_pt_rw[0][0]), (int(0.5 * w), int(0.13 * h)), font_scale, thickness, white),
        (frame, "Road's width: {:.2f} m".format(dist_rw), (int(0.35 * w), int(0.22 * h)), font_scale, thickness, white),
    ]
    return rects, texts
