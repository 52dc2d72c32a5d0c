"""GPU result banner (cv2.rectangle + cv2.putText, semantic_depth.py:339-394, sequence:304-327):
byte-exact against frames the reference's own statements drew with cv2 (tests/golden/banner_vectors.npz) and against the
oracle on random strings, batches and clipped glyphs."""
import os

import numpy as np
import pytest
import torch

from oracle import banner_ref
import semantic_depth_lib.pcl as pcl
from semantic_depth_b200 import banner

pytestmark = pytest.mark.gpu


def test_banner_golden_reference_frames(cuda_device, golden_dir):
    z = np.load(os.path.join(golden_dir, "banner_vectors.npz"))
    from tests_banner_cases import CASES, VALUES, base_frame
    for i, (name, driver, h, w, kw) in enumerate(CASES):
        assert [h, w] == [int(v) for v in z[f"case{i}_shape"][:2]]
        rows = int(z[f"case{i}_shape"][2])
        base = base_frame(h, w, i)
        if driver == "single":
            got = pcl.result_banner(base, kw["depth"], VALUES["left_pt_rw"], VALUES["right_pt_rw"], VALUES["dist_rw"],
                                    VALUES["left_pt_f2f"], VALUES["right_pt_f2f"], VALUES["dist_f2f"],
                                    is_city=kw["is_city"], approach=kw["approach"])
        else:
            got = pcl.sequence_banner(base, kw["depth"], kw["line_found"], VALUES["left_pt_rw"], VALUES["right_pt_rw"], VALUES["dist_rw"])
        assert got.dtype == np.uint8 and got.shape == base.shape
        assert np.array_equal(got[:rows], z[f"case{i}_top"]), name
        assert np.array_equal(got[rows:], base[rows:]), name


@pytest.mark.parametrize("scale,thick", [(2, 2), (4, 5), (2.2, 2)])
def test_banner_random_strings_against_oracle(cuda_device, scale, thick):
    rng = np.random.default_rng(int(scale * 10) + thick)
    b, h, w = 3, 420, 2600
    frames = rng.integers(0, 256, (b, h, w, 3), dtype=np.uint8)
    rects, texts = [], []
    want = frames.copy()
    for f in range(b):
        p1, p2 = (int(rng.integers(-20, 300)), int(rng.integers(-20, 100))), (int(rng.integers(500, 2700)), int(rng.integers(150, 500)))
        rcol = tuple(int(v) for v in rng.integers(0, 256, 3))
        rects.append((f, p1, p2, rcol))
        ts = []
        for line in range(3):
            n = int(rng.integers(1, 40))
            s = "".join(chr(int(rng.integers(32, 127))) for _ in range(n))
            org = (int(rng.integers(5, 300)), 140 + 130 * line)
            col = (255, 255 - 10 * f, 250)
            texts.append((f, s, org, scale, thick, col))
            ts.append((s, org, scale, thick, col))
        want[f] = banner_ref.draw(frames[f], [(p1, p2, rcol)], ts)
    dev = torch.from_numpy(frames).cuda()
    out, inside = banner.draw_banner(dev, rects, texts)
    assert out.data_ptr() == dev.data_ptr()                      # CUDA tensors are drawn in place
    assert np.array_equal(out.cpu().numpy(), want)
    # NumPy in -> NumPy out, input untouched
    src = frames[0].copy()
    got = pcl.draw_banner(src, [(0,) + rects[0][1:]], [(0,) + t[1:] for t in texts[:3]])
    assert np.array_equal(src, frames[0]) and np.array_equal(got, want[0])


def test_banner_clipping_and_odd_characters(cuda_device):
    h, w = 120, 300
    frame = np.full((h, w, 3), 9, np.uint8)
    texts = [(0, "Wgé\t", (-20, 30), 2, 2, (1, 2, 3)), (0, "edge", (250, 118), 2, 2, (200, 100, 50))]
    out, inside = banner.draw_banner(torch.from_numpy(frame).cuda(), [(0, (280, 100), (1000, 1000), (7, 7, 7))], texts)
    assert not inside
    want = banner_ref.draw(frame, [((280, 100), (1000, 1000), (7, 7, 7))], [t[1:] for t in texts])   # pixel-wise clipping, '?' for non-ASCII
    assert np.array_equal(out.cpu().numpy(), want)
    _, inside = banner.draw_banner(torch.from_numpy(frame).cuda(), [], [(0, "ok", (40, 80), 2, 2, (1, 2, 3))])
    assert inside
    with pytest.raises(ValueError):
        banner.draw_banner(torch.from_numpy(frame).cuda(), [], [(0, "x", (0, 50), 3.0, 2, (1, 2, 3))])      # preset not baked
    with pytest.raises(ValueError):
        banner.draw_banner(torch.from_numpy(frame).cuda(), [(1, (0, 0), (5, 5), (1, 2, 3))], [])            # frame outside the batch
