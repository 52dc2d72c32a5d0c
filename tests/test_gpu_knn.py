"""Statistical / radius outlier removal (uniform-grid CUDA search) against the cKDTree oracle, and the
RANSAC scorer against its NumPy specification (BASELINE.json configs 4 and 5, scaled to run in seconds)."""
import numpy as np
import pytest
import torch

from oracle import frame_ref, pcl_ref
from semantic_depth_b200 import scene
from semantic_depth_b200.pcl_gpu import engine_for
import semantic_depth_lib.pcl as pcl

pytestmark = pytest.mark.gpu


def soa(pts):
    t = torch.from_numpy(np.ascontiguousarray(pts.T)).cuda()
    return t[0].contiguous(), t[1].contiguous(), t[2].contiguous()


def clouds():
    rng = np.random.default_rng(5)
    out = {}
    out["road200k"] = scene.make_road_cloud(200_000, seed=0)
    out["uniform3d"] = rng.uniform(-1, 1, (20_000, 3)).astype(np.float32)
    out["line"] = np.c_[np.linspace(0, 10, 5000), np.zeros(5000), np.zeros(5000)].astype(np.float32)
    out["dups"] = np.repeat(rng.uniform(-1, 1, (500, 3)).astype(np.float32), 7, axis=0)
    out["tiny"] = rng.uniform(-1, 1, (6, 3)).astype(np.float32)
    out["single"] = np.float32([[1.0, 2.0, 3.0]])
    return out


@pytest.mark.parametrize("name,k", [("road200k", 16), ("road200k", 10), ("uniform3d", 20), ("line", 10), ("dups", 10),
                                    ("tiny", 10), ("single", 10), ("uniform3d", 1), ("uniform3d", 33)])
def test_knn_mean_distance_bit_exact(cuda_device, name, k):
    pts = clouds()[name]
    x, y, z = soa(pts)
    avg, stats = engine_for(pts.shape[0]).knn_mean_distance(x, y, z, k, 0.5)
    ref, _ = frame_ref.knn_mean_distances(pts, k)
    got = avg.cpu().numpy()
    assert np.array_equal(got, ref), (name, k, np.abs(got - ref).max(), int((got != ref).sum()))
    thr, mu, sd = frame_ref.sor_threshold(ref, 0.5)
    if pts.shape[0] > 1:
        assert abs(stats[0] - mu) <= 1e-12 * max(abs(mu), 1e-30) and abs(stats[2] - thr) <= 1e-10 * max(abs(thr), 1e-30)


@pytest.mark.parametrize("name,radius", [("road200k", 0.5), ("road200k", 0.05), ("uniform3d", 0.2), ("line", 0.011),
                                         ("dups", 0.1), ("tiny", 0.5), ("single", 0.5)])
def test_radius_counts_exact(cuda_device, name, radius):
    pts = clouds()[name]
    x, y, z = soa(pts)
    got = engine_for(pts.shape[0]).radius_count(x, y, z, radius, -1).cpu().numpy()
    ref = frame_ref.radius_counts(pts, radius)
    # d == r is the documented tie class (cKDTree counts it, FLANN does not); fp64 makes it measure-zero
    assert np.array_equal(got, ref), (name, radius, int((got != ref).sum()))
    capped = engine_for(pts.shape[0]).radius_count(x, y, z, radius, 80).cpu().numpy()
    assert np.array_equal(capped > 80, ref > 80)


def test_outlier_removal_facade(cuda_device):
    pts = scene.make_road_cloud(150_000, seed=3)
    cols = np.arange(pts.shape[0])
    keep, avg, _ = frame_ref.keep_statistical_outlier_removal(pts, 16, 0.5)
    p, c, k = pcl.statistical_outlier_removal(pts, cols, 16, 0.5, return_index=True)
    # ties within 1e-6 of the threshold are allowed to differ (north_star); there are none in practice
    assert p.dtype == np.float64 and np.array_equal(k.cpu().numpy(), keep)
    assert np.array_equal(p, pts[keep].astype(np.float64))
    keep2 = frame_ref.keep_radius_outlier_removal(p, 80, 0.5)
    p2, c2, k2 = pcl.radius_outlier_removal(p, c, 80, 0.5, return_index=True)
    assert np.array_equal(k2.cpu().numpy(), keep2) and np.array_equal(p2, p[keep2])


@pytest.mark.parametrize("axis,thr,K", [(1, 5.0, 1024), (0, 1.0, 2048), (2, 0.5, 300)])
def test_ransac_counts_bit_exact(cuda_device, axis, thr, K):
    rng = np.random.default_rng(11)
    n = 30_000
    pts = (rng.standard_normal((n, 3)) * np.array([3.0, 0.2, 15.0]) + np.array([0, -1.5, -30.0])).astype(np.float32)
    if axis == 0:
        pts = pts[:, [1, 0, 2]].copy()
    elif axis == 2:
        pts = pts[:, [0, 2, 1]].copy()
    trip = np.random.default_rng(1234).integers(0, n, (K, 3)).astype(np.int32)
    trip[5] = [7, 7, 9]                                  # repeated index -> invalid -> count 0
    x, y, z = soa(pts)
    counts, best, coeff = engine_for(n).ransac_score(x, y, z, axis, thr, torch.from_numpy(trip).cuda())
    ref = frame_ref.ransac_inlier_counts(pts, axis, thr, trip)
    assert np.array_equal(counts.cpu().numpy(), ref), int((counts.cpu().numpy() != ref).sum())
    assert best == int(np.argmax(ref)) and counts.cpu().numpy()[5] == 0
    keep, C, obest, _ = frame_ref.keep_plane_ransac(pts, axis, thr, trip)
    cols = np.arange(n)
    p, c, _, _, coeffs = pcl.remove_noise_by_fitting_plane(pts, cols, axis=axis, threshold=thr, hypotheses=trip)
    exp = pcl_ref.coefficients_dict(axis, C)
    for key in exp:
        assert abs(coeffs[key] - exp[key]) <= 1e-9 * max(1.0, abs(exp[key]))
    assert np.array_equal(c, keep)


def test_fused_with_ransac_hypotheses(cuda_device):
    from semantic_depth_b200.engine import FusionEngine
    from semantic_depth_b200.params import FusionParams
    h, w, K = 128, 256, 512
    logits, disp, intr = scene.make_frame(h, w, 6)
    base = frame_ref.fuse_frame(logits, disp, intr.as_q32(), intr.disparity_mult)
    rng = np.random.default_rng(1234)
    hyp = {"road": rng.integers(0, base["counts"]["road_mad_x"], (K, 3)).astype(np.int32),
           "left": rng.integers(0, base["counts"]["left_mad_x"], (K, 3)).astype(np.int32),
           "right": rng.integers(0, base["counts"]["right_mad_x"], (K, 3)).astype(np.int32)}
    o = frame_ref.fuse_frame(logits, disp, intr.as_q32(), intr.disparity_mult, hypotheses=hyp)
    eng = FusionEngine(h, w, max_frames=1, max_hypotheses=K, device=cuda_device)
    dh = {k: torch.from_numpy(v[None]).cuda() for k, v in hyp.items()}
    res = eng.fuse_frames(torch.from_numpy(logits[None]).cuda(), torch.from_numpy(disp[None]).cuda(), intr,
                          FusionParams(), dh)
    counts = res.counts(0)
    for name, c in o["counts"].items():
        assert counts[name] == c, (name, counts, dict(o["counts"]))
    for i, which in enumerate(("road", "left", "right")):
        assert int(res.raw["ransac_best"][0][i]) == o["ransac"][which]["best"]
    assert float(res.rw[0]) == o["rw"] and abs(float(res.f2f[0]) - o["f2f"]) <= 1e-3
