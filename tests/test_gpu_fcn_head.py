"""FCN-8s decoder head (SURVEY.md 8f rank 2, fcn8s/fcn.py:159-205) on the GPU against its NumPy restatement: second_skip
bit for bit, and -- through the fused score-map path -- the same labels, stage counts and answers as the oracle fed with
the logits the reference would have fetched."""
import numpy as np
import pytest
import torch

from oracle import fcn_ref, frame_ref
from semantic_depth_b200 import scene
from semantic_depth_b200.engine import FusionEngine
from semantic_depth_b200.fcn8s_head import Fcn8sHead, init_head_weights
from semantic_depth_b200.params import FusionParams

pytestmark = pytest.mark.gpu


def features(rng, b, h, w, channels):
    c3, c4, c7 = channels
    return (rng.standard_normal((b, h, w, c3), dtype=np.float32),
            rng.standard_normal((b, h // 2, w // 2, c4), dtype=np.float32),
            rng.standard_normal((b, h // 4, w // 4, c7), dtype=np.float32))


@pytest.mark.parametrize("b,h,w,channels", [(1, 8, 12, (256, 512, 4096)), (2, 16, 32, (256, 512, 4096)), (1, 12, 20, (40, 100, 333)),
                                            (3, 4, 4, (32, 31, 1))])
def test_head_bit_exact(cuda_device, b, h, w, channels):
    rng = np.random.default_rng(b * 1000 + h)
    wts = init_head_weights(seed=h, channels=channels, bias_std=0.05)
    l3, l4, l7 = features(rng, b, h, w, channels)
    want = fcn_ref.fcn8s_head(l3, l4, l7, wts)
    head = Fcn8sHead(wts, device=cuda_device, channels=channels)
    got = head(torch.from_numpy(l3).cuda(), torch.from_numpy(l4).cuda(), torch.from_numpy(l7).cuda()).cpu().numpy()
    assert got.shape == want.shape == (b, h, w, 3)
    assert np.array_equal(got, want), (np.abs(got - want).max(), int((got != want).sum()))


def test_head_rejects_bad_inputs(cuda_device):
    head = Fcn8sHead(device=cuda_device)
    l3 = torch.zeros((1, 8, 8, 256), device="cuda"); l4 = torch.zeros((1, 4, 4, 512), device="cuda"); l7 = torch.zeros((1, 2, 2, 4096), device="cuda")
    assert tuple(head(l3, l4, l7).shape) == (1, 8, 8, 3)
    with pytest.raises(ValueError):
        head(l3, l4, torch.zeros((1, 2, 2, 100), device="cuda"))
    with pytest.raises(TypeError):
        head(l3.cpu(), l4, l7)
    with pytest.raises(ValueError):
        Fcn8sHead({**init_head_weights(), "deconv1_w": np.zeros((4, 4, 3, 2), np.float32)}, device=cuda_device)


def test_head_feeds_the_fused_path(cuda_device):
    """Features -> head -> fuse_frames_scores on the device, against: oracle head -> upsample_scores -> oracle fuse_frame.
    The features are synthesised so that the head's output is the scene prior (the frame is non-degenerate)."""
    H, W, B = 128, 256, 2
    h, w = H // 8, W // 8
    channels = (64, 96, 160)
    wts = init_head_weights(seed=5, channels=channels)
    rng = np.random.default_rng(9)
    P = FusionParams()
    head = Fcn8sHead(wts, device=cuda_device, channels=channels)
    eng = FusionEngine(H, W, max_frames=B, device=cuda_device)
    l3s, l4s, l7s, disps, ups = [], [], [], [], None
    for f in range(B):
        sc, upw, upb, disp, intr = scene.make_frame_scores(H, W, seed=40 + f)
        # layer3 features that make conv_1x1_of_3 reproduce the scene scores (least squares through the random 1x1 kernel);
        # layer4 / layer7 carry small noise, so second_skip = scene scores + the decoder's contribution
        sol = np.linalg.lstsq(wts["conv3_w"].astype(np.float64).T, sc.reshape(-1, 3).astype(np.float64).T, rcond=None)[0].T
        l3s.append(sol.reshape(h, w, channels[0]).astype(np.float32))
        l4s.append(rng.standard_normal((h // 2, w // 2, channels[1]), dtype=np.float32) * np.float32(0.1))
        l7s.append(rng.standard_normal((h // 4, w // 4, channels[2]), dtype=np.float32) * np.float32(0.1))
        disps.append(disp); ups = (upw, upb)
    l3, l4, l7, disp = (np.ascontiguousarray(np.stack(v)) for v in (l3s, l4s, l7s, disps))
    scores = head(torch.from_numpy(l3).cuda(), torch.from_numpy(l4).cuda(), torch.from_numpy(l7).cuda())
    want_scores = fcn_ref.fcn8s_head(l3, l4, l7, wts)
    assert np.array_equal(scores.cpu().numpy(), want_scores)
    res = eng.fuse_frames_scores(scores, torch.from_numpy(ups[0]).cuda(), torch.from_numpy(ups[1]).cuda(), torch.from_numpy(disp).cuda(), intr, P)
    for f in range(B):
        logits = frame_ref.upsample_scores(want_scores[f], ups[0], ups[1])
        o = frame_ref.fuse_frame(logits, disp[f], intr.as_q32(), intr.disparity_mult, P)
        assert res.counts(f) == {k: int(v) for k, v in o["counts"].items()}, f
        assert int(res.status[f]) == o["status"]
        assert (o["rw"] is None and np.isnan(res.rw[f])) or float(res.rw[f]) == o["rw"]
        assert o["counts"]["road_ror"] > 100, "the synthesised features must give a non-degenerate frame"
