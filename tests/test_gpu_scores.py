"""Score-map mode (FCN-8s' last transposed convolution evaluated inside the label kernel) and
the argmax labelling north_star names, against the oracle."""
import numpy as np
import pytest
import torch

from oracle import frame_ref
from semantic_depth_b200 import scene
from semantic_depth_b200.engine import FusionEngine
from semantic_depth_b200.params import FusionParams
import semantic_depth_lib.pcl as pcl

pytestmark = pytest.mark.gpu


def dev(*arrs):
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrs]


@pytest.mark.parametrize("h,w,seed", [(64, 128, 0), (128, 256, 2), (256, 512, 1), (8, 16, 4)])
def test_upsampled_logits_and_labels_bit_exact(cuda_device, h, w, seed):
    sc, wt, bs, disp, intr = scene.make_frame_scores(h, w, seed)
    ref = frame_ref.upsample_scores(sc, wt, bs)
    eng = FusionEngine(h, w, max_frames=1, device=cuda_device)
    dsc, dw, db, dd = dev(sc[None], wt, bs, disp[None])
    out = eng.pixel_stage(None, dd, intr, scores=(dsc, dw, db))
    got = out["logits"][0].cpu().numpy()
    assert np.array_equal(got, ref), int((got != ref).sum())
    road, fence = frame_ref.labels_from_logits(ref)
    lab = out["labels"][0].cpu().numpy()
    assert np.array_equal((lab & 1) != 0, road) and np.array_equal((lab & 2) != 0, fence)
    # the facade spelling
    assert np.array_equal(pcl.upsample_scores(sc, wt, bs), ref)
    eng.close()


@pytest.mark.parametrize("h,w,seed", [(128, 256, 0), (256, 512, 3)])
def test_fused_from_scores_matches_oracle(cuda_device, h, w, seed):
    sc, wt, bs, disp, intr = scene.make_frame_scores(h, w, seed)
    logits = frame_ref.upsample_scores(sc, wt, bs)
    o = frame_ref.fuse_frame(logits, disp, intr.as_q32(), intr.disparity_mult, FusionParams())
    eng = FusionEngine(h, w, max_frames=2, device=cuda_device)
    dsc, dw, db, dd = dev(np.stack([sc, sc]), wt, bs, np.stack([disp, disp]))
    res = eng.fuse_frames_scores(dsc, dw, db, dd, intr, FusionParams())
    for f in range(2):
        assert res.counts(f) == dict(o["counts"]), (res.counts(f), dict(o["counts"]))
        assert int(res.status[f]) == o["status"]
        if o["rw"] is not None:
            assert float(res.rw[f]) == o["rw"]
        if o["f2f"] is not None:
            assert abs(float(res.f2f[f]) - o["f2f"]) <= max(1e-3, 1e-4 * o["f2f"])
    _, src = eng.final_cloud(0, "road")
    assert np.array_equal(src.cpu().numpy(), o["src"]["road_ror"])
    # same answers as the logits path fed with the oracle's upsampled logits
    res2 = eng.fuse_frames(torch.from_numpy(np.stack([logits, logits])).cuda(), dd, intr, FusionParams())
    assert res.raw.tobytes() == res2.raw.tobytes()
    eng.close()


def test_argmax_label_mode(cuda_device):
    h, w = 128, 256
    logits, disp, intr = scene.make_frame(h, w, 5)
    logits[:7] = np.float32([[2, 2, 1], [0, 3, 3], [1, 1, 1], [np.nan, 1, 0], [0, np.nan, 5], [-1, -2, -3], [0, 0, 1]])
    road, fence = frame_ref.labels_argmax(logits)
    eng = FusionEngine(h, w, max_frames=1, device=cuda_device)
    dl, dd = dev(logits[None], disp[None])
    out = eng.pixel_stage(dl, dd, intr, argmax=True)
    lab = out["labels"][0].cpu().numpy()
    assert np.array_equal((lab & 1) != 0, road) and np.array_equal((lab & 2) != 0, fence)
    res = eng.fuse_frames(dl, dd, intr, FusionParams(label_mode="argmax"))
    assert res.counts(0)["road_gather"] == int(road.sum()) and res.counts(0)["fence_gather"] == int(fence.sum())
    # argmax marks at least every pixel the reference's softmax > 0.5 rule marks
    r5, f5 = frame_ref.labels_from_logits(logits)
    assert np.all(road[r5]) and np.all(fence[f5])
    eng.close()
