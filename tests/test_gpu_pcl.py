"""semantic_depth_lib.pcl (CUDA) against the golden vectors of the reference's pcl.py and the oracle.

The calls below read like the reference's own call sites (semantic_depth.py:206-309)."""
import os

import numpy as np
import pytest
import torch

from oracle import frame_ref, pcl_ref
from semantic_depth_b200 import scene
import semantic_depth_lib.pcl as pcl

pytestmark = pytest.mark.gpu

CASES = ["rand1", "rand2", "rand7", "rand8", "rand9", "rand127", "rand128", "rand129", "rand1000", "rand4097",
         "dups", "mad_zero", "with_inf"]


@pytest.fixture(scope="module")
def vec(golden_dir):
    return np.load(os.path.join(golden_dir, "pcl_vectors.npz"))


def colors_for(pts):
    return (np.arange(pts.shape[0] * 3).reshape(-1, 3) % 251).astype(np.uint8)


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and bool(np.all((a == b) | ((a != a) & (b != b))))


@pytest.mark.parametrize("name", CASES)
def test_axis_filters_and_mad(cuda_device, vec, name):
    pts = vec[f"{name}/pts"]
    cols = colors_for(pts)
    p, c = pcl.remove_from_to(pts, cols, 2, 0.0, 7.0)
    k = vec[f"{name}/keep_z7"]
    assert same(p, pts[k]) and same(c, cols[k])
    p, c = pcl.threshold_complete(pts, cols, 2, 35.0)
    k = vec[f"{name}/keep_absz35"]
    assert same(p, pts[k]) and same(c, cols[k])
    for axis, thr in ((1, 15.0), (0, 2.0), (1, 5.0), (0, 5.0), (0, 1.0), (2, 3.0)):
        with np.errstate(all="ignore"):
            ad, m = pcl.mad(pts[:, axis])
        assert same(np.float32(m), np.float32(vec[f"{name}/mad_{axis}"])), (name, axis, m, vec[f"{name}/mad_{axis}"])
        p, c = pcl.remove_noise_by_mad(pts, cols, axis, thr)
        k = vec[f"{name}/keep_mad_{axis}_{thr}"]
        assert same(p, pts[k]) and same(c, cols[k]), (name, axis, thr)


@pytest.mark.parametrize("name", CASES)
def test_extract_pcls_numpy_mean(cuda_device, vec, name):
    pts = vec[f"{name}/pts"]
    cols = colors_for(pts)
    with np.errstate(all="ignore"):
        l, lc, r, rc = pcl.extract_pcls(pts, cols)
    kl, kr = vec[f"{name}/keep_left"], vec[f"{name}/keep_right"]
    assert same(l, pts[kl]) and same(r, pts[kr]) and same(lc, cols[kl]) and same(rc, cols[kr])


@pytest.mark.parametrize("n", [1, 5, 8, 100, 129, 130, 1000, 5000, 70001, 300007])
def test_mean_matches_numpy_pairwise(cuda_device, n):
    from semantic_depth_b200.pcl_gpu import engine_for
    rng = np.random.default_rng(n)
    col = (rng.standard_normal(n) * 3 - 0.3).astype(np.float32)
    got = engine_for(n).mean_f32(torch.from_numpy(col).cuda())
    assert np.float32(got) == np.mean(col), (n, got, np.mean(col))


@pytest.mark.parametrize("n", [1, 2, 3, 10, 511, 512, 4096, 4097, 100000, 333333])
def test_median_mad_matches_numpy(cuda_device, n):
    from semantic_depth_b200.pcl_gpu import engine_for
    rng = np.random.default_rng(n + 1)
    for kind in ("normal", "concentrated", "ties", "negzero"):
        col = rng.standard_normal(n).astype(np.float32)
        if kind == "concentrated":
            col = (-1.5 + 1e-4 * col).astype(np.float32)
        elif kind == "ties":
            col = np.round(col * 2).astype(np.float32)
        elif kind == "negzero":
            col[::3] = -0.0
            col[1::3] = 0.0
        med, m = engine_for(n).median_mad(torch.from_numpy(col).cuda())
        emed = np.median(col)
        emad = np.median(abs(col - emed))
        assert med == emed and m == emad, (n, kind, med, emed, m, emad)


@pytest.mark.parametrize("name", ["rand7", "rand9", "rand129", "rand1000", "rand4097", "dups", "fp64"])
def test_plane_fit(cuda_device, vec, name):
    pts = vec[f"{name}/pts"]
    cols = colors_for(pts)
    for axis, thr in ((1, 5.0), (0, 1.0), (2, 2.0), (1, 0.3)):
        p, c, plane3D, colors_plane, coeff = pcl.remove_noise_by_fitting_plane(pts, cols, axis=axis, threshold=thr,
                                                                               plane_color=[40, 70, 40])
        C = vec[f"{name}/coef_plane_{axis}"]
        exp = pcl_ref.coefficients_dict(axis, C)
        assert list(coeff) == ["Cx", "Cy", "Cz", "C"]
        for key in exp:
            assert abs(coeff[key] - exp[key]) <= 1e-9 * max(1.0, abs(exp[key])), (name, axis, key, coeff[key], exp[key])
        k = vec[f"{name}/keep_plane_{axis}_{thr}"]
        # inlier sets agree unless a residual sits within ~1e-9 of the threshold (documented tie class)
        res = np.abs(pcl_ref.plane_residual(pts, axis, C))
        if np.min(np.abs(res - thr)) > 1e-8:
            assert same(p, pts[k]) and same(c, cols[k]), (name, axis, thr)
        ref_plane, ref_cols = pcl_ref.plane_mesh(pts, axis, C, [40, 70, 40])
        assert plane3D.shape == ref_plane.shape and colors_plane.shape == ref_cols.shape
        np.testing.assert_allclose(plane3D, ref_plane, rtol=0, atol=1e-7)


def test_plane_fit_errors(cuda_device):
    with pytest.raises(ValueError):
        pcl.remove_noise_by_fitting_plane(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint8), axis=1)
    with pytest.raises(ValueError):
        pcl.remove_from_to(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint8), 2, 0.0, 7.0)


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_plane_fit_rank_deficient_is_min_norm(cuda_device, axis):
    """scipy.linalg.lstsq (pcl.py:120,154,186) answers a rank-deficient design matrix with the minimum-norm solution;
    the filter that follows must keep the same points and report the same coefficients."""
    rng = np.random.default_rng(17 + axis)
    ua, va = [(1, 2), (0, 2), (0, 1)][axis]                      # the two regressor columns of this axis
    clouds = {}
    p = (rng.standard_normal((3000, 3)) * [2.0, 0.5, 9.0] + [0.3, -1.5, -20.0]).astype(np.float32)
    q = p.copy(); q[:, va] = np.float32(-12.5); clouds["constant_v"] = q
    q = p.copy(); q[:, ua] = np.float32(0.75); clouds["constant_u"] = q
    q = p.copy(); q[:, ua] = np.float32(3.0); q[:, va] = np.float32(-7.0); clouds["constant_u_and_v"] = q
    t = rng.integers(-64, 64, 3000).astype(np.float32)
    q = p.copy(); q[:, ua] = 2.0 * t + 1.0; q[:, va] = t - 3.0; clouds["exactly_collinear"] = q      # u = 2 v + 7, exact in fp32
    clouds["one_point"] = p[:1].copy()
    clouds["two_points"] = p[:2].copy()
    clouds["same_point_twice"] = np.repeat(p[:1], 2, axis=0)
    for name, pts in clouds.items():
        for thr in (0.05, 0.5, 5.0):
            keep, C = pcl_ref.keep_plane(pts, axis, thr)
            if np.abs(C).max() > 1e6:
                # gelsd kept a rounding-level singular value (s_min ~ 1e-14 s_max > rcond = eps): the reference's own answer is
                # noise of size 1e11 here, not a minimum-norm plane -- nothing to be equal to (tie classes)
                continue
            got_p, _, _, _, coeff = pcl.remove_noise_by_fitting_plane(pts, colors_for(pts), axis=axis, threshold=thr)
            want = pcl_ref.coefficients_dict(axis, C)
            for k in ("Cx", "Cy", "Cz", "C"):
                assert abs(coeff[k] - want[k]) <= 1e-9 * max(1.0, abs(want[k])), (name, thr, k, coeff, want)
            if name != "exactly_collinear":          # collinear only up to rounding sits on gelsd's own rank cut-off
                assert same(got_p, pts[keep]), (name, thr, got_p.shape, int(keep.size))


@pytest.mark.parametrize("name", ["rand1000", "rand4097", "fp64", "with_inf", "rand1"])
def test_end_points_of_road(cuda_device, vec, name):
    pts = vec[f"{name}/pts"]
    for depth in (9.98, 30.0):
        a = pcl_ref.get_end_points_of_road(pts, depth)
        b = pcl.get_end_points_of_road(pts, depth)
        assert (a[0] is None) == (b[0] is None), (name, depth)
        if a[0] is not None:
            assert same(a[0], b[0]) and same(a[1], b[1]), (name, depth)


def test_end_points_of_segment_with_non_finite_rows(cuda_device):
    """pcl.py:307-308 take np.amin / np.amax over every row of the segment: rows whose z is +-inf or NaN count too."""
    rng = np.random.default_rng(3)
    seg = (rng.standard_normal((3000, 3)) * [3.0, 0.1, 10.0]).astype(np.float32)
    seg[np.argmin(seg[:, 0]), 2] = -np.inf            # the leftmost row has z = -inf
    seg[np.argmax(seg[:, 0]), 2] = np.nan             # the rightmost row has z = NaN
    seg[7, 2] = np.inf
    a, b = pcl_ref.get_end_points_of_segment(seg), pcl.get_end_points_of_segment(seg)
    assert same(a[0], b[0]) and same(a[1], b[1])
    seg[11, 0] = np.nan                               # NaN in x: amin / amax are NaN and `x == nan` selects no row
    with np.errstate(invalid="ignore"):
        a = pcl_ref.get_end_points_of_segment(seg)
    b = pcl.get_end_points_of_segment(seg)
    assert a[0].shape == b[0].shape == (0, 3) and a[1].shape == b[1].shape == (0, 3)


def test_thresholds_on_float64_clouds_compare_in_double(cuda_device):
    """After the Open3D round trip the reference's clouds are float64 (semantic_depth.py:244) and NumPy compares them with
    the Python threshold in double: a value equal to fl32(0.7) = 0.699999988 is < 0.7, and -fl32(7.1) is > -7.1."""
    lo, hi = np.float32(0.7), np.nextafter(np.float32(0.7), np.float32(1))
    assert float(lo) < 0.7 < float(hi)
    for dtype in (np.float64, np.float32):
        pts = np.zeros((6, 3), dtype)
        pts[:, 2] = [lo, hi, -lo, -hi, 0.5, np.float32(0.9)]          # float32-born values, as after the Open3D round trip
        cols = colors_for(pts)
        want = pts[pcl_ref.keep_threshold_complete(pts, 2, 0.7)]
        got, _ = pcl.threshold_complete(pts, cols, 2, 0.7)
        assert same(got, want), (dtype, got[:, 2], want[:, 2])
        t = np.float32(7.1)
        pts[:, 2] = [-t, np.nextafter(-t, np.float32(0)), np.nextafter(-t, np.float32(-10)), -7.5, -6.0, 1.0]
        want = pts[pcl_ref.keep_remove_from_to(pts, 2, 7.1)]
        got, _ = pcl.remove_from_to(pts, cols, 2, 0.0, 7.1)
        assert same(got, want), (dtype, got[:, 2], want[:, 2])


def test_intersection_distance_line(cuda_device, vec):
    road = {"Cx": 0.01, "Cy": -1.0, "Cz": 0.002, "C": -1.5}
    left = {"Cx": -1.0, "Cy": 0.03, "Cz": 0.001, "C": -4.0}
    got = pcl.planes_intersection_at_certain_depth(road, left, 10.0)
    assert same(got, vec["intersect/expected"])
    with pytest.raises(np.linalg.LinAlgError):
        pcl.planes_intersection_at_certain_depth(road, road, 10.0)
    pa, pb = np.float64([[1.0, 2.0, -10.0]]), np.float64([[-3.0, 2.5, -10.0]])
    assert pcl.compute_distance_in_3D(pa, pb) == np.linalg.norm(pa - pb)
    line, _ = pcl.create_3Dline_from_3Dpoints(pa, pb, [250, 0, 0])
    assert same(line, vec["line/expected"]) and pa[0][1] == 2.01     # in-place lift kept (pcl.py:322-323)


def test_torch_inputs_return_torch(cuda_device, vec):
    pts = vec["rand1000/pts"]
    t = torch.from_numpy(pts).cuda()
    cols = torch.from_numpy(colors_for(pts)).cuda()
    p, c = pcl.remove_noise_by_mad(t, cols, 1, 15.0)
    k = vec["rand1000/keep_mad_1_15.0"]
    assert p.is_cuda and c.is_cuda and same(p.cpu().numpy(), pts[k]) and same(c.cpu().numpy(), colors_for(pts)[k])


def test_reference_call_sequence_on_a_frame(cuda_device):
    """The reference's road + fence call sequence (semantic_depth.py:206-324) through the facade."""
    h, w = 256, 512
    logits, disp, intr = scene.make_frame(h, w, 0)
    o = frame_ref.fuse_frame(logits, disp, intr.as_q32(), intr.disparity_mult)
    road_mask, fence_mask = pcl.labels_from_logits(logits, (h, w))
    disparity = pcl.post_process_disparity(disp) * np.float32(intr.disparity_mult)
    points3D = pcl.reproject_to_3d(disparity, intr)
    colors = np.arange(h * w, dtype=np.int32).reshape(h, w)
    road3D, road_colors = points3D[road_mask], colors[road_mask]
    fence3D, fence_colors = points3D[fence_mask], colors[fence_mask]
    road3D, road_colors = pcl.remove_from_to(road3D, road_colors, 2, 0.0, 7.0)
    road3D, road_colors = pcl.remove_noise_by_mad(road3D, road_colors, 1, 15.0)
    road3D, road_colors = pcl.remove_noise_by_mad(road3D, road_colors, 0, 2.0)
    road3D, road_colors, _, _, road_coeff = pcl.remove_noise_by_fitting_plane(road3D, road_colors, axis=1, threshold=5.0,
                                                                              plane_color=[200, 200, 200])
    assert np.array_equal(road_colors, o["src"]["road_plane"])
    road3D, road_colors = pcl.statistical_outlier_removal(road3D, road_colors, nb_neighbors=10, std_ratio=0.5)
    assert np.array_equal(road_colors.astype(np.int64), o["src"]["road_sor"])
    road3D, road_colors = pcl.radius_outlier_removal(road3D, road_colors, nb_points=80, radius=0.5)
    assert road3D.dtype == np.float64 and np.array_equal(road_colors.astype(np.int64), o["src"]["road_ror"])
    left_pt, right_pt = pcl.get_end_points_of_road(road3D, 10.0 - 0.02)
    assert abs(left_pt[0][0] - right_pt[0][0]) == o["rw"]
    fence3D, fence_colors = pcl.remove_noise_by_mad(fence3D, fence_colors, 1, 5.0)
    fence3D, fence_colors = pcl.threshold_complete(fence3D, fence_colors, 2, 35.0)
    fl, flc, fr, frc = pcl.extract_pcls(fence3D, fence_colors)
    assert np.array_equal(flc, o["src"]["left_split"]) and np.array_equal(frc, o["src"]["right_split"])
    fl, flc = pcl.remove_noise_by_mad(fl, flc, 0, 5.0)
    fl, flc, _, _, left_coeff = pcl.remove_noise_by_fitting_plane(fl, flc, axis=0, threshold=1.0, plane_color=[40, 70, 40])
    fr, frc = pcl.remove_noise_by_mad(fr, frc, 0, 1.0)
    fr, frc, _, _, right_coeff = pcl.remove_noise_by_fitting_plane(fr, frc, axis=0, threshold=1.0, plane_color=[40, 70, 40])
    assert np.array_equal(flc, o["src"]["left_plane"]) and np.array_equal(frc, o["src"]["right_plane"])
    lp = pcl.planes_intersection_at_certain_depth(road_coeff, left_coeff, z=10.0)
    rp = pcl.planes_intersection_at_certain_depth(road_coeff, right_coeff, z=10.0)
    f2f = pcl.compute_distance_in_3D(lp, rp)
    assert abs(f2f - o["f2f"]) <= max(1e-3, 1e-4 * o["f2f"])
