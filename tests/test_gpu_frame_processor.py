"""The FrameProcessor-shaped driver: one frame from the BGR image and the two producers' raw outputs to
rw / f2f, masks, overlaid frame and the PLY dump -- every piece against the oracle."""
import numpy as np
import pytest

from oracle import frame_ref, overlay_ref, ply_ref
from semantic_depth_b200.frame_processor import FrameProcessor
from semantic_depth_b200.params import FusionParams, Intrinsics
from semantic_depth_b200.scene import make_frame

pytestmark = pytest.mark.gpu

H, W = 256, 512


class _Segmenter:
    def __init__(self, logits):
        self._logits = logits

    def logits(self, frame):
        assert tuple(frame.shape) == (H, W, 3)
        return self._logits


class _Depther:
    def __init__(self, disp):
        self._disp = disp

    def disparities(self, frame):
        return self._disp


def _oracle(logits, disp, intr, params):
    return frame_ref.fuse_frame(logits, disp, intr.as_q32(), intr.disparity_mult, params)


@pytest.mark.parametrize("approach", ["both", "rw"])
def test_process_frame_against_oracle(cuda_device, tmp_path, approach):
    logits, disp, _ = make_frame(H, W, seed=4)
    rng = np.random.default_rng(9)
    original = rng.integers(0, 256, (2 * H, 2 * W, 3), dtype=np.uint8)             # "cv2.imread" result, BGR
    intr = Intrinsics.synthetic(W)
    proc = FrameProcessor(_Segmenter(logits), _Depther(disp), (H, W), approach=approach, depth=10.0, intrinsics=intr,
                          disp_multiplier=W)
    out = proc.process_frame(original, output_name="f0", result_ply_dir=str(tmp_path))

    params = FusionParams(depth=10.0, approach=approach)
    want = _oracle(logits, disp, intr, params)
    assert out.line_found and abs(out.dist_rw - want["rw"]) <= 1e-3
    if approach == "both":
        assert abs(out.dist_f2f - want["f2f"]) <= max(1e-3, 1e-4 * abs(want["f2f"]))
        assert out.left_pt_f2f.shape == (1, 3)
    else:
        assert out.dist_f2f is None
    # masks, overlay, resize chain
    frame = frame_ref.resize_cubic_u8(original, W, H)
    road, fence = (m.reshape(H, W) for m in frame_ref.labels_from_logits(logits))
    assert np.array_equal(out.road_mask, road) and np.array_equal(out.fence_mask, fence)
    over = overlay_ref.overlay_masks(frame, road, fence)
    assert np.array_equal(out.segmented_frame, frame_ref.resize_cubic_u8(over, 2 * W, 2 * H))
    # final road cloud, colours (BGR -> RGB gather), PLY dump with the rw line appended
    road3D = out.road3D.cpu().numpy()
    assert road3D.shape[0] == out.counts["road_ror"] == want["counts"]["road_ror"]
    src = want["src"]["road_ror"]
    assert np.array_equal(out.road_colors.cpu().numpy(), frame[..., ::-1].reshape(-1, 3)[src])
    left, right = out.left_pt_rw[:1].astype(np.float64), out.right_pt_rw[:1].astype(np.float64)
    left[0][1] += 0.01; right[0][1] += 0.01
    t = np.arange(0.0, 1.0, 0.001)
    line = np.concatenate([left, left + t[:, None] * (right - left)], axis=0)
    line[:, 2] += 0.2
    want_ply = ply_ref.prepare_and_save_bytes(np.vstack([road3D.astype(np.float64), line]),
                                              np.vstack([out.road_colors.cpu().numpy().astype(np.float64),
                                                         np.ones(line.shape) * [250, 0, 0]]))
    assert open(out.ply_path, "rb").read() == want_ply


def test_process_frame_rejects_bad_producers(cuda_device):
    with pytest.raises(TypeError):
        FrameProcessor(object(), _Depther(None), (H, W))
    with pytest.raises(ValueError):
        FrameProcessor(_Segmenter(None), _Depther(None), (H, W), approach="f2f")


@pytest.mark.parametrize("mode", ["single", "sequence"])
def test_process_frame_draws_the_banner(cuda_device, mode):
    """Section 9 of process_frame (semantic_depth.py:339-394 / sequence:304-327) on a Cityscapes-sized original."""
    from oracle import banner_ref
    from semantic_depth_b200 import banner
    logits, disp, _ = make_frame(H, W, seed=4)
    rng = np.random.default_rng(11)
    h0, w0 = 1024, 2048
    original = rng.integers(0, 256, (h0, w0, 3), dtype=np.uint8)
    intr = Intrinsics.synthetic(W)
    plain = FrameProcessor(_Segmenter(logits), _Depther(disp), (H, W), approach="both", depth=10.0, intrinsics=intr,
                           disp_multiplier=W).process_frame(original)
    out = FrameProcessor(_Segmenter(logits), _Depther(disp), (H, W), approach="both", depth=10.0, intrinsics=intr,
                         disp_multiplier=W, banner=mode).process_frame(original)
    assert out.line_found and out.dist_rw == plain.dist_rw
    if mode == "single":
        rects, texts = banner.result_banner_spec(h0, w0, 10.0, out.left_pt_rw, out.right_pt_rw, out.dist_rw,
                                                 out.left_pt_f2f, out.right_pt_f2f, out.dist_f2f, is_city=True, approach="both")
    else:
        rects, texts = banner.sequence_banner_spec(h0, w0, 10.0, True, out.left_pt_rw, out.right_pt_rw, out.dist_rw)
    want = banner_ref.draw(plain.segmented_frame, [(p1, p2, c) for _, p1, p2, c in rects],
                           [(t, org, s, th, c) for _, t, org, s, th, c in texts])
    assert np.array_equal(out.segmented_frame, want)
    assert not np.array_equal(out.segmented_frame, plain.segmented_frame)
