"""Bake the glyph atlas of the result banner (BUILD CONTAINER ONLY: needs opencv; the .npz it writes is product data).

``python semantic_depth_b200/data/make_hershey_atlas.py [--verify]``

The reference draws its result banner with ``cv2.putText(..., fontFace=16, fontScale, thickness)``
(/root/reference/semantic_depth.py:350-394: scale 2 / thickness 2 for Cityscapes frames, 4 / 5 for the Munich set;
semantic_depth_cityscapes_sequence.py:306-327: scale 2 and 2.2, thickness 2).  fontFace 16 = FONT_HERSHEY_SIMPLEX |
FONT_ITALIC.  OpenCV rasterises Hershey stroke fonts in 16.16 fixed point: the pen starts at ``org.x << 16`` and moves by
``(right - left) * hscale`` per character, ``hscale = cvRound(fontScale * 65536)``; every stroke vertex is
``pen + unit * hscale``.  So the raster of a character relative to ``floor(pen)`` depends only on the character and on
``frac(pen) = frac(hscale * U / 65536)``, U = the advance units of the characters before it:

* integer scales: frac(pen) = 0 -- one raster per character, translation invariant;
* scale 2.2: hscale = 144179 = 2 * 65536 + 13107 and 5 * 13107 = 65535, so frac(pen) takes five values (U mod 5) up to a drift
  of 1/65536 px per five units, far below anything that moves a pixel in a banner line (checked by --verify on random strings).

The atlas holds, per preset and per phase ``U mod period``, one bitmap per printable ASCII character, rendered by OpenCV
itself (``n`` spaces -- which draw nothing -- in front of the character select the phase), plus the advance units.
The GPU kernel pastes those bitmaps; nothing is rasterised at run time.  Exactness domain: every glyph fully inside the
frame (OpenCV clips stroke *segments* at the border before Bresenham, which moves a few pixels of a clipped stroke).
"""
from __future__ import annotations

import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "hershey_atlas.npz")
FACE = 16
PRESETS = [(2.0, 2), (4.0, 5), (2.2, 2)]       # (fontScale, thickness)
FIRST, LAST = 32, 126
XY_ONE = 1 << 16
CUM_MAX = 1200       # pen positions (advance units) baked for fractional font scales: ~55 characters
SLACK = 1 << 11      # the pen's integer pixel is floor((pen + SLACK) / 65536): frac(pen) sits a hair BELOW its nominal phase (drift)


def advance_units(ch: str) -> int:
    """(right - left) of the glyph in Hershey units: getTextSize at scale 1 / thickness 1 is cvRound(sum + 1)."""
    (w, _), _ = cv2.getTextSize(ch, FACE, 1.0, 1)
    return int(w) - 1


def period_of(hscale: int) -> int:
    for p in range(1, 17):
        r = (hscale * p) % XY_ONE
        if min(r, XY_ONE - r) <= 8:
            return p
    raise ValueError("no short phase period for this font scale")


def prefix_for(units: np.ndarray, target: int, memo: dict) -> str | None:
    """A string that draws nothing near its end and advances the pen by exactly ``target`` units: ink + one trailing space."""
    sp = int(units[0])
    if target % sp == 0:
        return " " * (target // sp)
    rest = target - sp
    if "dp" not in memo:
        coins = {}
        for c in range(FIRST + 1, LAST + 1):
            coins.setdefault(int(units[c - FIRST]), chr(c))
        dp = {0: ""}
        for t in range(1, CUM_MAX + 64):
            for u, ch in coins.items():
                if t - u in dp:
                    dp[t] = dp[t - u] + ch
                    break
        memo["dp"] = dp
    pre = memo["dp"].get(rest)
    return None if pre is None else pre + " "


def bake(scale: float, thick: int):
    hscale = int(round(scale * XY_ONE))            # cvRound: half to even, like Python's round
    period = period_of(hscale)
    units = np.array([advance_units(chr(c)) for c in range(FIRST, LAST + 1)], np.int32)
    # integer scales: one raster per character.  Otherwise one per character and pen position U (in units) up to CUM_MAX:
    # frac(pen) = frac(hscale * U / 65536) drifts away from its nominal phase by a few 1/65536 px per character, which moves
    # the odd edge pixel of a diagonal stroke; equal rasters are stored once
    cums = [0] if period == 1 else list(range(0, CUM_MAX + 1))
    ox, oy, H, W = 8, 200, 288, 3200
    memo: dict = {}
    rasters = {}
    lo = [10 ** 9, 10 ** 9]; hi = [-10 ** 9, -10 ** 9]
    skipped = 0
    for cum in cums:
        pre = prefix_for(units, cum, memo)
        if pre is None:                           # unreachable pen positions are unreachable for any text as well
            skipped += 1
            continue
        pen = (ox << 16) + hscale * cum
        px = (pen + SLACK) >> 16
        img0 = np.zeros((H, W), np.uint8)
        if pre.strip():
            cv2.putText(img0, pre, (ox, oy), FACE, scale, 255, thick)
        x0 = max(px - 8, 0); x1 = min(px + 160, W)
        assert x1 - x0 > 150
        for c in range(FIRST, LAST + 1):
            img = np.zeros((H, x1 - x0 + 0), np.uint8)
            # render on a window that starts at x0 (integer pixel shifts are exact): the prefix ink is far to the left
            cv2.putText(img, pre + chr(c), (ox - x0, oy), FACE, scale, 255, thick)
            if pre.strip():
                img[img0[:, x0:x1] != 0] = 0
                assert not img0[:, max(px - 2, 0):x1].any(), "prefix ink reaches the glyph cell"
            ys, xs = np.nonzero(img)
            xs = xs + x0
            rasters[(cum, c)] = (xs - px, ys - oy)
            if len(xs):
                lo = [min(lo[0], int((xs - px).min())), min(lo[1], int((ys - oy).min()))]
                hi = [max(hi[0], int((xs - px).max())), max(hi[1], int((ys - oy).max()))]
    x_off, y_off = -lo[0], -lo[1]
    cell_w, cell_h = hi[0] - lo[0] + 1, hi[1] - lo[1] + 1
    words = (cell_w + 31) // 32
    nch = LAST - FIRST + 1
    index = np.full((nch, len(cums)), -1, np.int32)
    store: dict = {}
    bitmaps = []
    for (cum, c), (dx, dy) in rasters.items():
        bm = np.zeros((cell_h, words), np.uint32)
        col = dx + x_off; row = dy + y_off
        np.bitwise_or.at(bm, (row, col >> 5), (np.uint32(1) << (col & 31).astype(np.uint32)))
        key = bm.tobytes()
        if key not in store:
            store[key] = len(bitmaps); bitmaps.append(bm)
        index[c - FIRST, cum] = store[key]
    # pen positions no prefix reaches: the raster of the same phase one period earlier
    for ci in range(len(cums)):
        for g in range(nch):
            if index[g, ci] < 0:
                index[g, ci] = index[g, ci - period] if ci >= period else index[g, ci % period + period * 8]
    bits = np.stack(bitmaps)
    meta = np.array([FACE, int(round(scale * 1000)), thick, hscale, period, cell_h, words, x_off, y_off, len(cums) - 1], np.int64)
    print(f"  {len(bitmaps)} distinct rasters for {nch} characters x {len(cums)} pen positions ({skipped} unreachable)")
    return meta, units, bits, index


def glyph_of(meta, index, g: int, cum: int) -> int:
    """Bitmap of character g at pen position cum (units): beyond the baked range the same phase inside it."""
    period, cum_max = int(meta[4]), int(meta[9])
    if cum > cum_max:
        cum -= period * ((cum - cum_max + period - 1) // period)
    return int(index[g, cum])


def compose(atlas, preset: int, img: np.ndarray, text: str, org, color) -> None:
    """Reference composition in NumPy (what the GPU kernel does), used by --verify."""
    meta, units, bits, index = (atlas[f"p{preset}_{k}"] for k in ("meta", "units", "bits", "index"))
    hscale, period, cell_h, words, x_off, y_off = (int(v) for v in meta[3:9])
    cum = 0
    for ch in text:
        g = ord(ch) - FIRST
        pen = (int(org[0]) << 16) + hscale * cum
        px = (pen + SLACK) >> 16
        bm = bits[glyph_of(meta, index, g, cum)]
        rows, wds = np.nonzero(bm)
        for r, wd in zip(rows, wds):
            v = int(bm[r, wd])
            while v:
                b = (v & -v).bit_length() - 1
                v &= v - 1
                x = px + wd * 32 + b - x_off; y = int(org[1]) + r - y_off
                if 0 <= x < img.shape[1] and 0 <= y < img.shape[0]:
                    img[y, x] = color
        cum += int(units[g])


def verify(atlas) -> None:
    rng = np.random.default_rng(7)
    for p, (scale, thick) in enumerate(PRESETS):
        bad = 0
        for trial in range(60):
            n = int(rng.integers(1, 48))
            text = "".join(chr(int(rng.integers(FIRST, LAST + 1))) for _ in range(n))
            if trial == 0: text = "At 10.00m depth:"
            if trial == 1: text = "Cannot compute width of road at 10.00 m depth:"
            if trial == 2: text = "-3.98m to road's left end"
            base = rng.integers(0, 255, (420, 4400, 3), dtype=np.uint8)
            org = (int(rng.integers(5, 300)), int(rng.integers(200, 330)))
            ref = base.copy(); cv2.putText(ref, text, org, fontFace=FACE, fontScale=scale, color=(255, 254, 3), thickness=thick)
            mine = base.copy(); compose(atlas, p, mine, text, org, (255, 254, 3))
            if not np.array_equal(ref, mine):
                bad += 1
                print("MISMATCH", scale, thick, repr(text), int((ref != mine).any(2).sum()))
        print(f"preset {p} scale {scale} thickness {thick}: {60 - bad}/60 random strings byte-identical to cv2.putText")
        if bad:
            raise SystemExit(1)


def main() -> None:
    cv2.setNumThreads(1)
    out = {}
    for p, (scale, thick) in enumerate(PRESETS):
        meta, units, bits, index = bake(scale, thick)
        out[f"p{p}_meta"], out[f"p{p}_units"], out[f"p{p}_bits"], out[f"p{p}_index"] = meta, units, bits, index
        print(f"preset {p}: scale {scale} thickness {thick} hscale {meta[3]} period {meta[4]} cell {meta[5]}x{meta[6] * 32} offsets {meta[7]},{meta[8]}")
    if "--verify" in sys.argv:
        old = np.load(PATH)
        for k, v in out.items():
            assert np.array_equal(old[k], v), f"{k}: committed atlas differs from what OpenCV renders here"
        verify(old)
        print("atlas verified against opencv", cv2.__version__, "-- committed fixtures == live reference")
        return
    np.savez_compressed(PATH, **out)
    verify(np.load(PATH))
    print("wrote", PATH, os.path.getsize(PATH), "bytes")


if __name__ == "__main__":
    main()
