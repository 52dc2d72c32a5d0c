"""Pin the bicubic-resize oracle against the real ``cv2.resize(..., INTER_CUBIC)`` and store golden vectors.

BUILD CONTAINER ONLY (needs opencv; the fixture it writes is what travels): ``python tests/golden/make_golden_resize.py``.
The reference calls ``cv2.resize(frame, (512, 256), interpolation=cv2.INTER_CUBIC)`` on every frame
(/root/reference/semantic_depth.py:110-112).  cv2's SIMD builds evaluate the vertical pass of the vectorised part of
each row in fp32, so cv2's own output differs from OpenCV's fixed-point definition (``oracle.frame_ref.resize_cubic_u8``)
by at most 1 LSB on a machine-dependent subset of pixels; this script asserts that bound and records cv2's output.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import frame_ref  # noqa: E402

CASES = [(1024, 2048, 256, 512, 3), (300, 400, 256, 512, 3), (100, 37, 64, 128, 3), (64, 128, 256, 512, 1), (48, 64, 31, 47, 4)]


def make_image(h, w, c, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([((xx * (k + 3)) // 7 + (yy * (5 - k)) // 3) % 256 for k in range(c)], axis=-1)
    noise = rng.integers(-20, 21, (h, w, c))
    return np.clip(base + noise, 0, 255).astype(np.uint8)


def main():
    out = {}
    for i, (h, w, dh, dw, c) in enumerate(CASES):
        img = make_image(h, w, c, i)
        ref = cv2.resize(img if c > 1 else img[:, :, 0], (dw, dh), interpolation=cv2.INTER_CUBIC).reshape(dh, dw, c)
        mine = frame_ref.resize_cubic_u8(img, dw, dh)
        d = np.abs(ref.astype(int) - mine.astype(int))
        assert d.max() <= 1, (i, d.max())
        print(f"case {i} {h}x{w}x{c} -> {dh}x{dw}: max |cv2 - fixed point| = {d.max()}, differing {100 * (d > 0).mean():.3f} %")
        out[f"case{i}_shape"] = np.array([h, w, dh, dw, c, i])
        out[f"case{i}_cv2"] = ref
    path = os.path.join(HERE, "resize_vectors.npz")
    if len(sys.argv) > 1 and sys.argv[1] == "--verify":
        old = np.load(path)
        assert set(old.files) == set(out) and all(np.array_equal(old[k], out[k]) for k in out)
        print("committed fixtures == live reference")
    else:
        np.savez_compressed(path, **out)


if __name__ == "__main__":
    main()
