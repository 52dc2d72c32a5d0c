"""CPU suite: the oracle against the committed golden vectors.

The vectors under tests/golden were produced by tests/golden/make_golden.py in the build container, where the
reference's own semantic_depth_lib/pcl.py (imported unmodified), its DepthFrame.post_processing /
compute_3D_points (lifted with ast, real cv2) and the oracle were asserted bit-equal.  Here the oracle alone is
re-checked against those vectors (the reference does not exist on the GPU box)."""
import glob
import os
import sys

import numpy as np
import pytest

from oracle import frame_ref, pcl_ref
from semantic_depth_b200 import scene
from semantic_depth_b200.params import FusionParams, Intrinsics

CASES = ["rand1", "rand2", "rand7", "rand8", "rand9", "rand127", "rand128", "rand129", "rand1000", "rand4097",
         "dups", "mad_zero", "with_inf", "fp64"]


@pytest.fixture(scope="module")
def vec(golden_dir):
    return np.load(os.path.join(golden_dir, "pcl_vectors.npz"))


@pytest.mark.parametrize("name", CASES)
def test_pcl_keep_indices(vec, name):
    pts = vec[f"{name}/pts"]
    with np.errstate(all="ignore"):
        assert np.array_equal(pcl_ref.keep_remove_from_to(pts, 2, 7.0), vec[f"{name}/keep_z7"])
        assert np.array_equal(pcl_ref.keep_threshold_complete(pts, 2, 35.0), vec[f"{name}/keep_absz35"])
        for axis, thr in ((1, 15.0), (0, 2.0), (1, 5.0), (0, 5.0), (0, 1.0), (2, 3.0)):
            assert np.array_equal(pcl_ref.keep_mad(pts, axis, thr), vec[f"{name}/keep_mad_{axis}_{thr}"])
            _, m = pcl_ref.mad(pts[:, axis])
            a, b = np.asarray(m), vec[f"{name}/mad_{axis}"]
            assert (a == b) or (np.isnan(a) and np.isnan(b))
        kl, kr, mean = pcl_ref.keep_extract_pcls(pts)
        assert np.array_equal(kl, vec[f"{name}/keep_left"]) and np.array_equal(kr, vec[f"{name}/keep_right"])
        for depth in (9.98, 30.0):
            assert np.array_equal(pcl_ref.keep_slab(pts, depth), vec[f"{name}/keep_slab_{depth}"])
        if f"{name}/coef_plane_1" in vec:
            for axis, thr in ((1, 5.0), (0, 1.0), (2, 2.0), (1, 0.3)):
                keep, C = pcl_ref.keep_plane(pts, axis, thr)
                np.testing.assert_allclose(C, vec[f"{name}/coef_plane_{axis}"], rtol=1e-9, atol=1e-11)
                res = np.abs(pcl_ref.plane_residual(pts, axis, vec[f"{name}/coef_plane_{axis}"]))
                if np.min(np.abs(res - thr)) > 1e-8:
                    assert np.array_equal(keep, vec[f"{name}/keep_plane_{axis}_{thr}"])


def test_pcl_surface_shapes(vec):
    pts = vec["rand1000/pts"]
    cols = (np.arange(pts.shape[0] * 3).reshape(-1, 3) % 251).astype(np.uint8)
    out = pcl_ref.remove_noise_by_fitting_plane(pts, cols, axis=1, threshold=5.0, plane_color=[200, 200, 200])
    assert len(out) == 5 and list(out[4]) == ["Cx", "Cy", "Cz", "C"] and out[4]["Cy"] == -1.0
    assert len(pcl_ref.extract_pcls(pts, cols)) == 4
    assert pcl_ref.get_end_points_of_road(pts[:0], 9.98) == (None, None)
    got = pcl_ref.planes_intersection_at_certain_depth({"Cx": 0.01, "Cy": -1.0, "Cz": 0.002, "C": -1.5},
                                                       {"Cx": -1.0, "Cy": 0.03, "Cz": 0.001, "C": -4.0}, 10.0)
    assert np.array_equal(got, vec["intersect/expected"])
    pa, pb = np.float64([[1.0, 2.0, -10.0]]), np.float64([[-3.0, 2.5, -10.0]])
    line, _ = pcl_ref.create_3Dline_from_3Dpoints(pa, pb, [250, 0, 0])
    assert np.array_equal(line, vec["line/expected"])
    with pytest.raises(ValueError):
        pcl_ref.remove_from_to(pts[:0], cols[:0], 2, 0.0, 7.0)


def test_pixel_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "pixel_vectors.npz"))
    for tag in ("city", "munich", "synth"):
        blend = frame_ref.post_process_disparity(g[f"{tag}/disp"])
        assert blend.dtype == np.float32 and np.array_equal(blend.view(np.uint32), g[f"{tag}/blend"].view(np.uint32))
        with np.errstate(all="ignore"):
            pts = frame_ref.reproject_to_3d(blend * np.float32(g[f"{tag}/mult"]), g[f"{tag}/q32"])
        ref = g[f"{tag}/points"]
        assert np.all((pts == ref) | (np.isnan(pts) & np.isnan(ref)))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "frame_*.npz"))))
def test_frame_fixture(path):
    g = np.load(path)
    h, w, seed = int(g["h"]), int(g["w"]), int(g["seed"])
    if h * w > 256 * 512:
        pytest.skip("large fixture is exercised by the GPU suite (keeps the CPU suite short)")
    logits, disp, intr = scene.make_frame(h, w, seed)
    o = frame_ref.fuse_frame(logits, disp, intr.as_q32(), intr.disparity_mult, FusionParams())
    for name, c in o["counts"].items():
        assert c == int(g[f"count/{name}"]), name
        if f"src/{name}" in g:
            assert np.array_equal(o["src"][name], g[f"src/{name}"]), name
    assert o["status"] == int(g["status"])
    rw = float(g["rw"])
    assert (o["rw"] is None and np.isnan(rw)) or o["rw"] == rw
    assert o["f2f"] == float(g["f2f"])


def test_ransac_oracle_reduces_to_least_squares_shape():
    rng = np.random.default_rng(0)
    pts = (rng.standard_normal((2000, 3)) * np.array([3.0, 0.05, 10.0]) + np.array([0, -1.5, -30.0])).astype(np.float32)
    trip = rng.integers(0, 2000, (64, 3))
    trip[3] = [5, 5, 9]
    counts = frame_ref.ransac_inlier_counts(pts, 1, 0.1, trip)
    assert counts.shape == (64,) and counts[3] == 0 and counts.max() > 1000
    keep, C, best, _ = frame_ref.keep_plane_ransac(pts, 1, 0.1, trip)
    assert best == int(np.argmax(counts)) and abs(C[2] + 1.5) < 0.05 and keep.size > 1000


def test_sor_ror_oracle_small():
    pts = scene.make_road_cloud(3000, seed=1)
    keep, avg, (thr, mu, sd) = frame_ref.keep_statistical_outlier_removal(pts, 10, 0.5)
    assert avg.shape == (3000,) and np.all(avg > 0) and thr == mu + 0.5 * sd and 0 < keep.size < 3000
    cnt = frame_ref.radius_counts(pts, 0.5)
    assert cnt.min() >= 1 and np.array_equal(frame_ref.keep_radius_outlier_removal(pts, 80, 0.5), np.flatnonzero(cnt > 80))


def test_upsample_oracle_matches_scatter_formulation():
    """the gather form of the 16x16 / stride-8 transposed convolution (what the kernel evaluates)
    against the textbook scatter form in fp64."""
    from semantic_depth_b200 import scene
    sc, w, b, disp, intr = scene.make_frame_scores(64, 128, 3)
    lg = frame_ref.upsample_scores(sc, w, b)
    h, wd, _ = sc.shape
    ref = np.zeros((8 * h + 16, 8 * wd + 16, 3))
    for iy in range(h):
        for ix in range(wd):
            ref[8 * iy:8 * iy + 16, 8 * ix:8 * ix + 16, :] += np.einsum("c,yxoc->yxo", sc[iy, ix].astype(np.float64), w.astype(np.float64))
    ref = ref[4:4 + 8 * h, 4:4 + 8 * wd] + b
    assert lg.dtype == np.float32 and lg.shape == (64 * 128, 3)
    assert np.abs(ref.reshape(-1, 3) - lg).max() < 1e-5
    # the scene stays usable in this mode: both classes present and the fused oracle answers
    o = frame_ref.fuse_frame(lg, disp, intr.as_q32(), intr.disparity_mult)
    assert o["counts"]["road_gather"] > 500 and o["counts"]["fence_gather"] > 500


def test_argmax_labels_oracle():
    lg = np.float32([[1, 0, 0], [0, 1, 0], [0, 0, 1], [2, 2, 1], [0, 3, 3], [np.nan, 1, 0]])
    road, fence = frame_ref.labels_argmax(lg)
    assert road.tolist() == [True, False, False, True, False, True]      # first maximum wins; NaN wins like np.argmax
    assert fence.tolist() == [False, True, False, False, True, False]


def test_resize_oracle_against_cv2_golden(golden_dir):
    """the bicubic-resize oracle equals the recorded bytes of cv2's own implementation (IPP off) and stays
    within 1 LSB of what cv2 returned through Intel's closed-source IPP path (see make_golden_resize.py)."""
    sys.path.insert(0, golden_dir)
    from make_golden_resize import make_image
    z = np.load(os.path.join(golden_dir, "resize_vectors.npz"))
    n = len([k for k in z.files if k.endswith("_shape")])
    assert n >= 4
    for i in range(n):
        h, w, dh, dw, c, seed = (int(v) for v in z[f"case{i}_shape"])
        got = frame_ref.resize_cubic_u8(make_image(h, w, c, seed), dw, dh)
        assert got.shape == (dh, dw, c) and np.array_equal(got, z[f"case{i}_cv2"]), i
        d = np.abs(got.astype(int) - z[f"case{i}_cv2_ipp"].astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 0.1


def test_ply_oracle_against_reference_digests(golden_dir):
    """oracle.ply_ref against the digests of the files the reference's own PointCloud2Ply wrote."""
    import hashlib
    from oracle import ply_ref
    sys.path.insert(0, golden_dir)
    from make_golden_ply import make_cloud
    z = np.load(os.path.join(golden_dir, "ply_vectors.npz"))
    for i in range(len([k for k in z.files if k.endswith("_sha256")])):
        seed, n, is64, nbytes = (int(v) for v in z[f"case{i}"])
        p, c = make_cloud(seed, n, np.float64 if is64 else np.float32, is64 == 2)
        got = ply_ref.prepare_and_save_bytes(p, c)
        assert len(got) == nbytes and hashlib.sha256(got).digest() == z[f"case{i}_sha256"].tobytes()


def test_overlay_oracle_against_pil_golden(golden_dir):
    """oracle.overlay_ref against frames PIL's own Image.paste produced (make_golden_overlay.py)."""
    from oracle import overlay_ref
    z = np.load(os.path.join(golden_dir, "overlay_vectors.npz"))
    n = len([k for k in z.files if k.endswith("_out")])
    assert n >= 7
    for i in range(n):
        got = overlay_ref.overlay_from_labels(z[f"case{i}_frame"], z[f"case{i}_labels"],
                                              tuple(z[f"case{i}_road_rgba"]), tuple(z[f"case{i}_fence_rgba"]))
        assert np.array_equal(got, z[f"case{i}_out"]), i
    # the bytescale corner cases the layer constants come from
    assert overlay_ref.mask_rgba(np.array([[True, False]]), overlay_ref.ROAD_RGBA)[0, 0].tolist() == [255, 128, 255, 128]
    assert overlay_ref.mask_rgba(np.array([[True, False]]), overlay_ref.FENCE_RGBA)[0, 0].tolist() == [255, 16, 16, 102]
    assert overlay_ref.mask_rgba(np.array([[True, True]]), overlay_ref.ROAD_RGBA)[0, 0].tolist() == [255, 0, 255, 0]
    assert overlay_ref.mask_rgba(np.array([[True, True]]), overlay_ref.FENCE_RGBA)[0, 0].tolist() == [255, 0, 0, 92]
    assert not overlay_ref.mask_rgba(np.array([[False, False]]), overlay_ref.ROAD_RGBA).any()


def test_banner_oracle_against_cv2_golden(golden_dir):
    """SURVEY 8f rank 4 (banner + putText): oracle.banner_ref + the baked glyph atlas + the product's own line / position glue
    against frames the reference's own statements drew with cv2 (make_golden_banner.py)."""
    from oracle import banner_ref
    from semantic_depth_b200 import banner
    from tests_banner_cases import CASES, VALUES, base_frame
    z = np.load(os.path.join(golden_dir, "banner_vectors.npz"))
    assert [str(n) for n in z["names"]] == [c[0] for c in CASES]
    for i, (name, driver, h, w, kw) in enumerate(CASES):
        if driver == "single":
            rects, texts = banner.result_banner_spec(h, w, kw["depth"], VALUES["left_pt_rw"], VALUES["right_pt_rw"], VALUES["dist_rw"],
                                                     VALUES["left_pt_f2f"], VALUES["right_pt_f2f"], VALUES["dist_f2f"],
                                                     is_city=kw["is_city"], approach=kw["approach"])
        else:
            rects, texts = banner.sequence_banner_spec(h, w, kw["depth"], kw["line_found"], VALUES["left_pt_rw"],
                                                       VALUES["right_pt_rw"], VALUES["dist_rw"])
        rows = int(z[f"case{i}_shape"][2])
        base = base_frame(h, w, i)
        got = banner_ref.draw(base, [(p1, p2, c) for _, p1, p2, c in rects], [(t, org, s, th, c) for _, t, org, s, th, c in texts])
        assert np.array_equal(got[:rows], z[f"case{i}_top"]), name
        assert np.array_equal(got[rows:], base[rows:]), name
        # the product's placement arithmetic is the oracle's (same pen pixel, same bitmap) for every line
        for _, text, org, scale, thick, _ in texts:
            pi = banner.preset_index(scale, thick)
            assert pi == banner_ref.preset_of(scale, thick)
            assert len(banner.layout_text(pi, text, org)) == len(text)
    assert banner.layout_text(0, "\u00e9", (0, 50)) == banner.layout_text(0, "?", (0, 50))      # cv2 draws '?' for non-ASCII
    with pytest.raises(ValueError):
        banner.preset_index(3.0, 2)


def test_config4_fixture_reproducible(golden_dir):
    """BASELINE.json configs[3]: the committed answers of the 2 M-point statistical filter are what the oracle computes."""
    import hashlib
    from oracle import frame_ref
    g = np.load(os.path.join(golden_dir, "config4_sor.npz"))
    pts = scene.make_road_cloud(int(g["n"]), seed=int(g["seed"]))
    keep, avg, (thr, mu, sd) = frame_ref.keep_statistical_outlier_removal(pts, int(g["k"]), float(g["ratio"]), workers=-1)
    assert hashlib.sha256(np.ascontiguousarray(avg).tobytes()).digest() == g["avg_sha256"].tobytes()
    assert thr == float(g["thr"]) and mu == float(g["mean"]) and sd == float(g["std"])
    assert keep.size == int(g["kept"]) and np.array_equal(avg[g["sample_idx"]], g["sample_avg"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/semantic_depth_lib"), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("script,args", [("make_golden.py", ["--verify"]), ("make_golden_ply.py", ["--verify"]),
                                         ("make_golden_overlay.py", ["--verify"]), ("make_golden_resize.py", ["--verify"]),
                                         ("make_golden_banner.py", ["--verify"]),
                                         ("../../semantic_depth_b200/data/make_hershey_atlas.py", ["--verify"])])
def test_committed_fixtures_equal_live_reference(golden_dir, script, args):
    """Build container only: re-run the generators against the reference's own code (its unmodified pcl.py, the
    DepthFrame methods lifted from semantic_depth.py, its PointCloud2Ply; PIL's paste for the overlay) and compare with the
    committed fixtures.  A subprocess keeps the reference's `semantic_depth_lib` out of this process."""
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(golden_dir, script), *args], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-1500:])
    assert "committed fixtures == live reference" in out.stdout
