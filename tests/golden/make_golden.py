"""Pin the oracle against the reference and write the golden fixtures.  BUILD CONTAINER ONLY.

Run from the repo root:  ``python tests/golden/make_golden.py``

What it does (needs ``/root/reference``, which does not exist on the GPU box -- the fixtures it
writes are what travels):

1. imports the reference's ``semantic_depth_lib/pcl.py`` UNMODIFIED (sys.path) and lifts
   ``DepthFrame.post_processing`` / ``DepthFrame.compute_3D_points`` (semantic_depth.py:656-664,
   686-697) with ``ast`` -- the module itself cannot be imported (tensorflow / open3d imports);
2. runs reference and oracle on identical inputs and ASSERTS bit-equality function by function
   (the only deliberate deviation: ``planes_intersection_at_certain_depth`` raises on NumPy >= 1.24
   in the reference, pcl.py:235, so it is compared against ``inv(A) * B`` computed by the
   reference's own lines 226-233);
3. drives the reference functions in the order/constants of semantic_depth.py:183-324 on synthetic
   frames (Open3D SOR/ROR replaced by the oracle's cKDTree stand-ins -- Open3D is not installable)
   and asserts the oracle's ``fuse_frame`` reproduces every per-stage cloud bit for bit;
4. stores inputs' seeds + expected outputs under tests/golden/*.npz.
"""
from __future__ import annotations

import ast
import io
import contextlib
import os
import sys
import textwrap

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import frame_ref, pcl_ref  # noqa: E402
from semantic_depth_b200 import scene  # noqa: E402
from semantic_depth_b200.params import FusionParams, Intrinsics  # noqa: E402


def load_reference():
    sys.path.insert(0, REF)
    import semantic_depth_lib.pcl as ref_pcl  # the reference module, unmodified
    assert ref_pcl.__file__.startswith(REF), ref_pcl.__file__
    src = open(os.path.join(REF, "semantic_depth.py")).read()
    tree = ast.parse(src)
    methods = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == "DepthFrame":
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name in ("post_processing", "compute_3D_points"):
                    methods[fn.name] = textwrap.dedent(ast.get_source_segment(src, fn))
    import cv2
    ns = {"np": np, "cv2": cv2}
    for code in methods.values():
        exec(code, ns)
    return ref_pcl, ns["post_processing"], ns["compute_3D_points"]


class _Self:
    def __init__(self, intr):
        self.cx, self.cy, self.f, self.b = intr.cx, intr.cy, intr.f, intr.b


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and bool(np.all((a == b) | ((a != a) & (b != b))))


def check_functions(ref_pcl, rng):
    """Function-level bit-equality on random and edge-case clouds; returns the vectors to store."""
    vec = {}
    cases = {}
    for n in (1, 2, 7, 8, 9, 127, 128, 129, 1000, 4097):
        pts = (rng.standard_normal((n, 3)) * np.array([3.0, 0.5, 20.0]) + np.array([0.0, -1.5, -30.0])).astype(np.float32)
        cases[f"rand{n}"] = pts
    dup = np.repeat(cases["rand9"], 3, axis=0)
    cases["dups"] = dup
    const = np.tile(np.float32([[1.0, -1.5, -10.0]]), (50, 1))
    const[:5, 0] = np.float32([0.5, 0.7, 0.9, 1.1, 1.3])
    cases["mad_zero"] = const
    withinf = cases["rand1000"].copy()
    withinf[3] = np.float32([np.inf, -np.inf, -np.inf])
    cases["with_inf"] = withinf
    cases["fp64"] = cases["rand1000"].astype(np.float64)
    for name, pts in cases.items():
        cols = (np.arange(pts.shape[0] * 3).reshape(-1, 3) % 251).astype(np.uint8)
        vec[f"{name}/pts"] = pts
        with np.errstate(all="ignore"):
            # remove_from_to
            a = ref_pcl.remove_from_to(pts, cols, 2, 0.0, 7.0)
            b = pcl_ref.remove_from_to(pts, cols, 2, 0.0, 7.0)
            assert same(a[0], b[0]) and same(a[1], b[1]), name
            vec[f"{name}/keep_z7"] = pcl_ref.keep_remove_from_to(pts, 2, 7.0).astype(np.int32)
            # MAD on each axis with the reference's thresholds
            for axis, thr in ((1, 15.0), (0, 2.0), (1, 5.0), (0, 5.0), (0, 1.0), (2, 3.0)):
                a = ref_pcl.remove_noise_by_mad(pts, cols, axis, thr)
                b = pcl_ref.remove_noise_by_mad(pts, cols, axis, thr)
                assert same(a[0], b[0]) and same(a[1], b[1]), (name, axis, thr)
                vec[f"{name}/keep_mad_{axis}_{thr}"] = pcl_ref.keep_mad(pts, axis, thr).astype(np.int32)
                ad_r, m_r = ref_pcl.mad(pts[:, axis])
                ad_o, m_o = pcl_ref.mad(pts[:, axis])
                assert same(ad_r, ad_o) and same(m_r, m_o)
                vec[f"{name}/median_{axis}"] = np.asarray(np.median(pts[:, axis]))
                vec[f"{name}/mad_{axis}"] = np.asarray(m_o)
            # threshold_complete, extract_pcls
            a = ref_pcl.threshold_complete(pts, cols, 2, 35.0)
            b = pcl_ref.threshold_complete(pts, cols, 2, 35.0)
            assert same(a[0], b[0]) and same(a[1], b[1]), name
            vec[f"{name}/keep_absz35"] = pcl_ref.keep_threshold_complete(pts, 2, 35.0).astype(np.int32)
            a = ref_pcl.extract_pcls(pts, cols)
            b = pcl_ref.extract_pcls(pts, cols)
            assert all(same(x, y) for x, y in zip(a, b)), name
            kl, kr, mean = pcl_ref.keep_extract_pcls(pts)
            vec[f"{name}/keep_left"], vec[f"{name}/keep_right"] = kl.astype(np.int32), kr.astype(np.int32)
            vec[f"{name}/mean_x"] = np.asarray(mean)
            # plane fit on the three axes
            if pts.shape[0] >= 3 and np.all(np.isfinite(pts)) and name != "mad_zero":
                for axis, thr in ((1, 5.0), (0, 1.0), (2, 2.0), (1, 0.3)):
                    a = ref_pcl.remove_noise_by_fitting_plane(pts, cols, axis=axis, threshold=thr, plane_color=[40, 70, 40])
                    b = pcl_ref.remove_noise_by_fitting_plane(pts, cols, axis=axis, threshold=thr, plane_color=[40, 70, 40])
                    assert all(same(x, y) for x, y in zip(a[:4], b[:4])), (name, axis)
                    assert a[4] == b[4] and list(a[4]) == list(b[4]), (name, axis)
                    keep, C = pcl_ref.keep_plane(pts, axis, thr)
                    vec[f"{name}/keep_plane_{axis}_{thr}"] = keep.astype(np.int32)
                    vec[f"{name}/coef_plane_{axis}"] = C
            # slab end points at the reference's depth
            for depth in (9.98, 30.0):
                a = ref_pcl.get_end_points_of_road(pts, depth)
                b = pcl_ref.get_end_points_of_road(pts, depth)
                assert (a[0] is None) == (b[0] is None)
                if a[0] is not None:
                    assert same(a[0], b[0]) and same(a[1], b[1])
                vec[f"{name}/keep_slab_{depth}"] = pcl_ref.keep_slab(pts, depth).astype(np.int32)
    # plane-plane intersection: reference lines 217-233 (the array build at 235 raises on NumPy>=1.24)
    road = {"Cx": 0.01, "Cy": -1.0, "Cz": 0.002, "C": -1.5}
    left = {"Cx": -1.0, "Cy": 0.03, "Cz": 0.001, "C": -4.0}
    z = -10.0
    A = np.matrix([[road["Cx"], road["Cy"]], [left["Cx"], left["Cy"]]])
    B = np.matrix([[-(road["Cz"] * z + road["C"])], [-(left["Cz"] * z + left["C"])]])
    X = np.linalg.inv(A) * B
    got = pcl_ref.planes_intersection_at_certain_depth(road, left, 10.0)
    assert got.shape == (1, 3) and got.dtype == np.float64
    assert got[0, 0] == X[0, 0] and got[0, 1] == X[1, 0] and got[0, 2] == z
    vec["intersect/expected"] = got
    try:
        ref_pcl.planes_intersection_at_certain_depth(road, left, 10.0)
        print("note: reference intersection did not raise under this NumPy")
    except ValueError:
        pass
    # distance + line
    pa, pb = np.float64([[1.0, 2.0, -10.0]]), np.float64([[-3.0, 2.5, -10.0]])
    assert ref_pcl.compute_distance_in_3D(pa, pb) == pcl_ref.compute_distance_in_3D(pa, pb)
    la, lb = ref_pcl.create_3Dline_from_3Dpoints(pa.copy(), pb.copy(), [250, 0, 0])
    oa, ob = pcl_ref.create_3Dline_from_3Dpoints(pa.copy(), pb.copy(), [250, 0, 0])
    assert same(la, oa) and same(lb, ob)
    vec["line/expected"] = oa
    return vec


def reference_chain(ref_pcl, post_processing, compute_3D_points, logits, disp, intr, P):
    """The reference's own functions in the order of semantic_depth.py:183-324 (Open3D -> oracle stand-in)."""
    h, w = disp.shape[1:]
    road_mask, fence_mask = frame_ref.labels_from_logits(logits)
    road_mask, fence_mask = road_mask.reshape(h, w), fence_mask.reshape(h, w)
    disparity = post_processing(None, disp).astype(np.float32)                     # :676
    disparity = disparity * int(intr.disparity_mult)                               # :145
    with contextlib.redirect_stdout(io.StringIO()):
        points3D = compute_3D_points(_Self(intr), disparity)                       # :160
    colors = np.arange(h * w, dtype=np.int64).reshape(h, w)                        # colors := source pixel index
    stages = {}
    road3D, road_src = points3D[road_mask], colors[road_mask]                      # :183-184
    fence3D, fence_src = points3D[fence_mask], colors[fence_mask]                  # :186-187
    stages["road_gather"], stages["fence_gather"] = (road3D, road_src), (fence3D, fence_src)
    road3D, road_src = ref_pcl.remove_from_to(road3D, road_src, 2, 0.0, 7.0)       # :206
    stages["road_z"] = (road3D, road_src)
    road3D, road_src = ref_pcl.remove_noise_by_mad(road3D, road_src, 1, 15.0)      # :209
    stages["road_mad_y"] = (road3D, road_src)
    road3D, road_src = ref_pcl.remove_noise_by_mad(road3D, road_src, 0, 2.0)       # :212
    stages["road_mad_x"] = (road3D, road_src)
    road3D, road_src, _, _, road_coeff = ref_pcl.remove_noise_by_fitting_plane(    # :215-219
        road3D, road_src, axis=1, threshold=5.0, plane_color=[200, 200, 200])
    stages["road_plane"] = (road3D, road_src)
    k = frame_ref.keep_statistical_outlier_removal(road3D, 10, 0.5)[0]             # :234-236 (stand-in)
    road3D, road_src = road3D.astype(np.float64)[k], road_src[k]
    stages["road_sor"] = (road3D, road_src)
    k = frame_ref.keep_radius_outlier_removal(road3D, 80, 0.5)                     # :238-241 (stand-in)
    road3D, road_src = road3D[k], road_src[k]
    stages["road_ror"] = (road3D, road_src)
    left_pt, right_pt = ref_pcl.get_end_points_of_road(road3D, P.depth - 0.02)     # :254-255
    rw = None if left_pt is None else abs(left_pt[0][0] - right_pt[0][0])          # :259
    fence3D, fence_src = ref_pcl.remove_noise_by_mad(fence3D, fence_src, 1, 5.0)   # :279
    stages["fence_mad_y"] = (fence3D, fence_src)
    fence3D, fence_src = ref_pcl.threshold_complete(fence3D, fence_src, 2, 35.0)   # :283-284
    stages["fence_abs_z"] = (fence3D, fence_src)
    fl, fl_src, fr, fr_src = ref_pcl.extract_pcls(fence3D, fence_src)              # :286-287
    stages["left_split"], stages["right_split"] = (fl, fl_src), (fr, fr_src)
    fl, fl_src = ref_pcl.remove_noise_by_mad(fl, fl_src, 0, 5.0)                   # :291
    stages["left_mad_x"] = (fl, fl_src)
    fl, fl_src, _, _, left_coeff = ref_pcl.remove_noise_by_fitting_plane(fl, fl_src, axis=0, threshold=1.0,
                                                                         plane_color=[40, 70, 40])  # :294-298
    stages["left_plane"] = (fl, fl_src)
    fr, fr_src = ref_pcl.remove_noise_by_mad(fr, fr_src, 0, 1.0)                   # :302
    stages["right_mad_x"] = (fr, fr_src)
    fr, fr_src, _, _, right_coeff = ref_pcl.remove_noise_by_fitting_plane(fr, fr_src, axis=0, threshold=1.0,
                                                                          plane_color=[40, 70, 40])  # :305-309
    stages["right_plane"] = (fr, fr_src)
    # :317-324 via the reference's lines 217-233 (see check_functions for the NumPy>=1.24 note)
    def intersect(c1, c2, z):
        z = -z
        A = np.matrix([[c1["Cx"], c1["Cy"]], [c2["Cx"], c2["Cy"]]])
        B = np.matrix([[-(c1["Cz"] * z + c1["C"])], [-(c2["Cz"] * z + c2["C"])]])
        X = np.linalg.inv(A) * B
        return np.array([[X[0, 0], X[1, 0], z]], np.float64)
    f2f = ref_pcl.compute_distance_in_3D(intersect(road_coeff, left_coeff, P.depth),
                                         intersect(road_coeff, right_coeff, P.depth))
    return {"stages": stages, "rw": rw, "f2f": f2f, "points3D": points3D.reshape(-1, 3),
            "disparity": disparity, "coeff": {"road": road_coeff, "left": left_coeff, "right": right_coeff}}


FRAME_CASES = ((64, 128, 0), (128, 256, 0), (128, 256, 3), (256, 512, 0), (256, 512, 1), (512, 1024, 0),
               (1024, 2048, 0))       # the last one is BASELINE.json's full size (about a minute of CPU)


def _store(path, fix, verify):
    """Write the fixture, or (verify) assert that the committed file holds exactly these arrays."""
    if not verify:
        np.savez_compressed(path, **fix)
        return
    old = np.load(path)
    assert set(old.files) == set(fix), (path, set(old.files) ^ set(fix))
    for k, v in fix.items():
        assert same(old[k], np.asarray(v)), (path, k)


def check_frames(ref_pcl, post_processing, compute_3D_points, cases=FRAME_CASES, verify=False):
    P = FusionParams()
    for (h, w, seed) in cases:
        logits, disp, intr = scene.make_frame(h, w, seed)
        ref = reference_chain(ref_pcl, post_processing, compute_3D_points, logits, disp, intr, P)
        # pixel-stage oracle functions against the lifted reference methods
        pp = frame_ref.post_process_disparity(disp)
        assert same(pp * np.float32(intr.disparity_mult), ref["disparity"])
        assert same(frame_ref.reproject_to_3d(ref["disparity"], intr.as_q32()).reshape(-1, 3), ref["points3D"])
        o = frame_ref.fuse_frame(logits, disp, intr.as_q32(), intr.disparity_mult, P)
        for name, (pts, src) in ref["stages"].items():
            assert o["counts"][name] == pts.shape[0], (h, w, seed, name)
            assert np.array_equal(o["src"][name], src), (h, w, seed, name)
        assert (ref["rw"] is None and o["rw"] is None) or ref["rw"] == o["rw"], (ref["rw"], o["rw"])
        assert ref["f2f"] == o["f2f"], (ref["f2f"], o["f2f"])
        for which in ("road", "left", "right"):
            assert ref["coeff"][which] == o["coeff"][which]
        fix = {"h": h, "w": w, "seed": seed, "rw": np.float64(np.nan if o["rw"] is None else o["rw"]),
               "f2f": np.float64(o["f2f"]), "status": np.int64(o["status"]),
               "labels": np.packbits(o["labels"] == 1), "labels_fence": np.packbits(o["labels"] == 2),
               "sor_thr": np.float64(o["sor"]["thr"]), "fence_mean_x": np.float32(o["fence_mean_x"]),
               "xl": np.float64(o.get("xl", np.nan)), "xr": np.float64(o.get("xr", np.nan))}
        for name, c in o["counts"].items():
            fix[f"count/{name}"] = np.int64(c)
        if h * w <= 256 * 512:
            for name, s in o["src"].items():
                fix[f"src/{name}"] = s.astype(np.int32)
        else:  # large frame: a checksum of the index lists keeps the fixture small
            for name, s in o["src"].items():
                fix[f"srcsum/{name}"] = np.uint64(int(np.sum(s.astype(np.uint64) * (np.arange(s.size, dtype=np.uint64) % 65521 + 1)) % (1 << 63)))
        for which in ("road", "left", "right"):
            c = o["coeff"][which]
            fix[f"coeff/{which}"] = np.float64([c["Cx"], c["Cy"], c["Cz"], c["C"]])
        _store(os.path.join(HERE, f"frame_{h}x{w}_seed{seed}.npz"), fix, verify)
        print(f"frame {h}x{w} seed {seed}: reference == oracle on {len(ref['stages'])} stages; "
              f"rw={o['rw']} f2f={o['f2f']} status={o['status']} counts={dict(o['counts'])}")


def check_pixel_vectors(post_processing, compute_3D_points):
    """Small pixel-stage vectors incl. d = 0 / negative disparities and both intrinsics presets."""
    rng = np.random.default_rng(7)
    out = {}
    for tag, (h, w, intr) in {"city": (24, 64, Intrinsics.cityscapes(64)), "munich": (16, 40, Intrinsics.munich()),
                              "synth": (32, 128, Intrinsics.synthetic(128))}.items():
        disp = rng.uniform(0.001, 0.2, (2, h, w)).astype(np.float32)
        disp[0, 0, :3] = 0.0
        disp[1, 0, -3:] = 0.0
        disp[0, 1, 5] = -0.01
        pp = post_processing(None, disp).astype(np.float32)
        assert same(pp, frame_ref.post_process_disparity(disp))
        scaled = pp * int(intr.disparity_mult)
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
            pts = compute_3D_points(_Self(intr), scaled)
        assert same(pts, frame_ref.reproject_to_3d(scaled, intr.as_q32()))
        out[f"{tag}/disp"], out[f"{tag}/blend"], out[f"{tag}/points"] = disp, pp, pts
        out[f"{tag}/q32"], out[f"{tag}/mult"] = intr.as_q32(), np.float32(intr.disparity_mult)
    return out


def main():
    ref_pcl, post_processing, compute_3D_points = load_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "--only-full-size":
        # optional seeds after the flag: the frames of bench.py's first batch are seeds 0..4
        seeds = [int(v) for v in sys.argv[2:]] or [0]
        check_frames(ref_pcl, post_processing, compute_3D_points, cases=tuple((1024, 2048, sd) for sd in seeds))
        return
    # --verify: recompute everything from the live reference and compare with the committed files instead of writing
    verify = len(sys.argv) > 1 and sys.argv[1] == "--verify"
    rng = np.random.default_rng(20260101)
    vec = check_functions(ref_pcl, rng)
    _store(os.path.join(HERE, "pcl_vectors.npz"), vec, verify)
    print(f"pcl function vectors: {len(vec)} arrays, reference == oracle")
    pix = check_pixel_vectors(post_processing, compute_3D_points)
    _store(os.path.join(HERE, "pixel_vectors.npz"), pix, verify)
    print(f"pixel vectors: {len(pix)} arrays, reference (cv2) == oracle")
    check_frames(ref_pcl, post_processing, compute_3D_points, cases=FRAME_CASES[:5] if verify else FRAME_CASES, verify=verify)
    if verify:
        print("committed fixtures == live reference")


if __name__ == "__main__":
    main()
