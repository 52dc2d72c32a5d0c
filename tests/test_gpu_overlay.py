"""GPU segmentation overlay (semantic_depth.py:547-568): byte-exact against the PIL-pinned
golden vectors and against the oracle on seeded frames, including the bytescale corner cases (empty / full masks)."""
import os

import numpy as np
import pytest
import torch

from oracle import frame_ref, overlay_ref
import semantic_depth_lib.pcl as pcl

pytestmark = pytest.mark.gpu


def test_overlay_golden_vectors(cuda_device, golden_dir):
    z = np.load(os.path.join(golden_dir, "overlay_vectors.npz"))
    n = len([k for k in z.files if k.endswith("_out")])
    assert n >= 7
    for i in range(n):
        frame, labels = z[f"case{i}_frame"], z[f"case{i}_labels"]
        got = pcl.overlay_masks(frame, (labels & 1) != 0, (labels & 2) != 0,
                                tuple(int(v) for v in z[f"case{i}_road_rgba"]), tuple(int(v) for v in z[f"case{i}_fence_rgba"]))
        assert got.dtype == np.uint8 and np.array_equal(got, z[f"case{i}_out"]), i


@pytest.mark.parametrize("h,w", [(256, 512), (1024, 2048), (37, 53)])
def test_overlay_batch_against_oracle(cuda_device, h, w):
    rng = np.random.default_rng(h + w)
    b = 3
    frames = rng.integers(0, 256, (b, h, w, 3), dtype=np.uint8)
    u = rng.random((b, h, w))
    road, fence = u < 0.3, u > 0.6
    road[1] = True; fence[1] = False          # frame 1: full road mask (alpha collapses to 0), per-frame variant selection
    fence[2] = False                          # frame 2: empty fence mask
    out = pcl.overlay_masks(torch.from_numpy(frames).cuda(), torch.from_numpy(road).cuda(), torch.from_numpy(fence).cuda())
    assert isinstance(out, torch.Tensor) and out.is_cuda
    out = out.cpu().numpy()
    for k in range(b):
        assert np.array_equal(out[k], overlay_ref.overlay_masks(frames[k], road[k], fence[k])), k
    assert np.array_equal(out[1], frames[1])


def test_overlay_unaligned_views(cuda_device):
    rng = np.random.default_rng(5)
    h, w = 64, 96
    frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    labels = rng.integers(0, 4, (h, w), dtype=np.uint8)
    want = overlay_ref.overlay_from_labels(frame, labels)
    # the scalar kernel: odd byte offsets of all three device buffers
    from semantic_depth_b200 import frame_ops
    fbuf = torch.zeros(h * w * 3 + 1, dtype=torch.uint8, device="cuda")
    lbuf = torch.zeros(h * w + 3, dtype=torch.uint8, device="cuda")
    fbuf[1:] = torch.from_numpy(frame).cuda().reshape(-1)
    lbuf[3:] = torch.from_numpy(labels).cuda().reshape(-1)
    got = frame_ops._overlay_labels(fbuf[1:].view(1, h, w, 3), lbuf[3:].view(1, h * w), overlay_ref.ROAD_RGBA, overlay_ref.FENCE_RGBA)
    assert np.array_equal(got[0].cpu().numpy(), want)


def test_segment_frame_triple(cuda_device):
    from semantic_depth_b200.scene import make_frame
    h, w = 128, 256
    logits, _disp = make_frame(h, w, seed=3)[:2]
    rng = np.random.default_rng(0)
    frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    road, fence, over = pcl.segment_frame(frame, logits)
    r0, f0 = (m.reshape(h, w) for m in frame_ref.labels_from_logits(logits))
    assert road.shape == (h, w, 1) and road.dtype == bool
    assert np.array_equal(road[..., 0], r0) and np.array_equal(fence[..., 0], f0)
    assert np.array_equal(over, overlay_ref.overlay_masks(frame, r0, f0))


def test_overlay_rejects_bad_arguments(cuda_device):
    frame = np.zeros((8, 8, 3), dtype=np.uint8)
    m = np.zeros((8, 8), dtype=bool)
    with pytest.raises(Exception):
        pcl.overlay_masks(frame, m, m, road_rgba=(300, 0, 0, 64))
    with pytest.raises(ValueError):
        pcl.overlay_masks(frame.astype(np.float32), m, m)
