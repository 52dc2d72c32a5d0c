"""GPU bicubic resize (semantic_depth.py:110-112): byte-identical to the recorded outputs of cv2's own
implementation (IPP off) and to the oracle; within 1 LSB of what cv2 returns when it dispatches to Intel's closed-source IPP."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import frame_ref
import semantic_depth_lib.pcl as pcl

pytestmark = pytest.mark.gpu


def test_resize_byte_identical_to_cv2(cuda_device, golden_dir):
    sys.path.insert(0, golden_dir)
    from make_golden_resize import make_image
    z = np.load(os.path.join(golden_dir, "resize_vectors.npz"))
    n = len([k for k in z.files if k.endswith("_shape")])
    for i in range(n):
        h, w, dh, dw, c, seed = (int(v) for v in z[f"case{i}_shape"])
        img = make_image(h, w, c, seed)
        got = pcl.resize_cubic(img, (dw, dh))
        assert got.dtype == np.uint8 and got.shape == (dh, dw, c)
        assert np.array_equal(got, frame_ref.resize_cubic_u8(img, dw, dh)), i
        assert np.array_equal(got, z[f"case{i}_cv2"]), (i, int((got != z[f"case{i}_cv2"]).sum()))
        assert np.abs(got.astype(int) - z[f"case{i}_cv2_ipp"].astype(int)).max() <= 1


def test_resize_batch_and_torch(cuda_device):
    rng = np.random.default_rng(3)
    batch = rng.integers(0, 256, (3, 120, 200, 3), dtype=np.uint8)
    out = pcl.resize_cubic(torch.from_numpy(batch).cuda(), (512, 256))
    assert isinstance(out, torch.Tensor) and out.is_cuda and tuple(out.shape) == (3, 256, 512, 3)
    for b in range(3):
        assert np.array_equal(out[b].cpu().numpy(), frame_ref.resize_cubic_u8(batch[b], 512, 256))
    gray = rng.integers(0, 256, (77, 91), dtype=np.uint8)
    assert np.array_equal(pcl.resize_cubic(gray, (40, 30)), frame_ref.resize_cubic_u8(gray, 40, 30)[:, :, 0])
