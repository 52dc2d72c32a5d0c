"""Pin the bicubic-resize oracle against the real ``cv2.resize(..., INTER_CUBIC)`` and store golden vectors.

BUILD CONTAINER ONLY (needs opencv; the fixture it writes is what travels): ``python tests/golden/make_golden_resize.py``.
The reference calls ``cv2.resize(frame, (512, 256), interpolation=cv2.INTER_CUBIC)`` on every frame
(/root/reference/semantic_depth.py:110-112).  The oracle restates OpenCV's OWN implementation (resize.cpp: integer
horizontal pass, fp32 SIMD vertical pass on the first 8*floor(n/8) elements of a row, integer tail) and must equal cv2
BYTE FOR BYTE with IPP switched off; this script asserts that and records those bytes (``caseN_cv2``).  With IPP on (the
wheel's default) cv2 calls Intel's closed-source ippiResizeCubic instead: its output (``caseN_cv2_ipp``) is recorded too and
must stay within 1 LSB of OpenCV's own code.
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


CASES = CASES + [(40, 33, 20, 5, 3), (25, 31, 17, 9, 1), (30, 30, 12, 7, 4)]      # rows of 15, 9, 28 elements: fp32 part + integer tail


def main():
    out = {}
    for i, (h, w, dh, dw, c) in enumerate(CASES):
        img = make_image(h, w, c, i)
        src = img if c > 1 else img[:, :, 0]
        ipp_was = cv2.ipp.useIPP()
        cv2.ipp.setUseIPP(False)
        ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_CUBIC).reshape(dh, dw, c)
        cv2.ipp.setUseIPP(ipp_was)
        ref_ipp = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_CUBIC).reshape(dh, dw, c)
        mine = frame_ref.resize_cubic_u8(img, dw, dh)
        assert np.array_equal(ref, mine), (i, int((ref != mine).sum()))
        d = np.abs(ref_ipp.astype(int) - mine.astype(int))
        assert d.max() <= 1, (i, d.max())
        print(f"case {i} {h}x{w}x{c} -> {dh}x{dw}: oracle == cv2 (IPP off) byte for byte; cv2 with IPP {'on' if ipp_was else 'off'}: "
              f"max diff {d.max()}, differing {100 * (d > 0).mean():.3f} %")
        out[f"case{i}_shape"] = np.array([h, w, dh, dw, c, i])
        out[f"case{i}_cv2"] = ref
        out[f"case{i}_cv2_ipp"] = ref_ipp
    path = os.path.join(HERE, "resize_vectors.npz")
    if len(sys.argv) > 1 and sys.argv[1] == "--verify":
        old = np.load(path)
        assert set(old.files) == set(out) and all(np.array_equal(old[k], out[k]) for k in out)
        print("committed fixtures == live reference")
    else:
        np.savez_compressed(path, **out)


if __name__ == "__main__":
    main()
