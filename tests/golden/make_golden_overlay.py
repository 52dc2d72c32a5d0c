"""Pin oracle/overlay_ref.py against Pillow's Image.paste.  BUILD CONTAINER ONLY (needs PIL).

``python tests/golden/make_golden_overlay.py`` executes the statements of SegmentFrame.segment_frame
(/root/reference/semantic_depth.py:547-568) with the real PIL paste and a `toimage` shim (scipy.misc is gone from
SciPy >= 1.3; the shim is scipy 1.2.1's 3-D branch: bytescale + Image.frombytes), asserts equality with the oracle
and writes tests/golden/overlay_vectors.npz."""
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import overlay_ref  # noqa: E402


def toimage(arr, mode=None):
    data = np.asarray(arr)
    bytedata = overlay_ref.bytescale(data)
    h, w, ch = bytedata.shape
    if mode is None:
        mode = "RGB" if ch == 3 else "RGBA"
    return Image.frombytes(mode, (w, h), np.ascontiguousarray(bytedata).tobytes())


def segment_frame_overlay(frame, segmentation_road, segmentation_fence, road_rgba, fence_rgba):
    street_im = toimage(frame)
    road_mask = np.dot(segmentation_road, np.array([list(road_rgba)]))
    road_mask = toimage(road_mask, mode="RGBA")
    street_im.paste(road_mask, box=None, mask=road_mask)
    fence_mask = np.dot(segmentation_fence, np.array([list(fence_rgba)]))
    fence_mask = toimage(fence_mask, mode="RGBA")
    street_im.paste(fence_mask, box=None, mask=fence_mask)
    return np.array(street_im)


def make_case(i):
    rng = np.random.default_rng(100 + i)
    h, w = [(64, 128), (48, 80), (33, 57), (64, 128), (64, 128), (40, 72), (31, 45)][i]
    frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    u = rng.random((h, w))
    road, fence = u < 0.35, u > 0.7
    rc, fc = overlay_ref.ROAD_RGBA, overlay_ref.FENCE_RGBA
    if i == 1:
        road[:] = True; fence[:] = False           # full road mask: bytescale maps alpha 64 -> 0
    if i == 2:
        road[:] = False; fence[:] = True           # full fence mask: alpha 64 -> 92
    if i == 3:
        road[:] = False; fence[:] = False
    if i == 4:
        fence = u > 0.2                            # overlapping masks (cannot come out of the label kernel; still defined)
    if i == 5:
        rc, fc = (10, 200, 30, 255), (0, 0, 250, 17)
    if i == 6:
        frame[:] = 255
    labels = (road.astype(np.uint8) | (fence.astype(np.uint8) << 1))
    return frame, labels, np.array(rc), np.array(fc)


def main():
    out = {}
    for i in range(7):
        frame, labels, rc, fc = make_case(i)
        h, w = labels.shape
        road = ((labels & 1) != 0).reshape(h, w, 1)
        fence = ((labels & 2) != 0).reshape(h, w, 1)
        ref = segment_frame_overlay(frame, road, fence, rc, fc)
        mine = overlay_ref.overlay_from_labels(frame, labels, tuple(rc), tuple(fc))
        assert mine.dtype == ref.dtype and mine.shape == ref.shape and (mine == ref).all(), i
        out[f"case{i}_frame"], out[f"case{i}_labels"], out[f"case{i}_out"] = frame, labels, ref
        out[f"case{i}_road_rgba"], out[f"case{i}_fence_rgba"] = rc, fc
        print(f"case {i}: {h}x{w}, {int((ref != frame).sum())} bytes changed, oracle == PIL")
    path = os.path.join(HERE, "overlay_vectors.npz")
    if len(sys.argv) > 1 and sys.argv[1] == "--verify":
        old = np.load(path)
        assert set(old.files) == set(out) and all(np.array_equal(old[k], out[k]) for k in out)
        print("committed fixtures == live reference")
    else:
        np.savez_compressed(path, **out)


if __name__ == "__main__":
    main()
