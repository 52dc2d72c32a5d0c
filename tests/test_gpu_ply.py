"""GPU PLY writer against the oracle restatement of PointCloud2Ply and the golden digests of the
files the reference's own class wrote (tests/golden/make_golden_ply.py)."""
import hashlib
import os
import sys

import numpy as np
import pytest

from oracle import ply_ref
from semantic_depth_lib.point_cloud_2_ply import PointCloud2Ply

pytestmark = pytest.mark.gpu


def test_ply_files_byte_identical(cuda_device, golden_dir, tmp_path):
    sys.path.insert(0, golden_dir)
    from make_golden_ply import make_cloud
    z = np.load(os.path.join(golden_dir, "ply_vectors.npz"))
    ncases = len([k for k in z.files if k.endswith("_sha256")])
    assert ncases >= 4
    for i in range(ncases):
        seed, n, is64, nbytes = (int(v) for v in z[f"case{i}"])
        p, c = make_cloud(seed, n, np.float64 if is64 else np.float32, is64 == 2)
        w = PointCloud2Ply(p.copy(), c.copy(), str(tmp_path / f"cloud{i}"))
        w.prepare_and_save_point_cloud()
        got = open(tmp_path / f"cloud{i}.ply", "rb").read()
        assert got == ply_ref.prepare_and_save_bytes(p, c), i
        assert len(got) == nbytes and hashlib.sha256(got).digest() == z[f"case{i}_sha256"].tobytes()
        assert got[:400] == z[f"case{i}_head"].tobytes()


def test_ply_special_values_and_extra_cloud(cuda_device, tmp_path):
    vals = np.float32([0.0, -0.0, 0.5, 1.5, 2.5e-6, 3.5e-6, 4.9999999e-7, 5.0000001e-7, -5e-7, 1e-45, 0.9999995, 0.99999994,
                       123456.7890625, -1e6, 16777216.0, 1e10, 3.4028235e38, -3.4028235e38, 1.17549435e-38, 1234.5678,
                       np.inf, -np.inf, np.nan, 7.0000005])
    p = np.stack([vals, vals[::-1], np.linspace(-5, 5, len(vals), dtype=np.float32)], axis=1)
    c = (np.arange(len(vals) * 3).reshape(-1, 3) * 7 % 256).astype(np.uint8)
    w = PointCloud2Ply(p, c, str(tmp_path / "special"))
    w.write_ply(str(tmp_path / "special.ply"))
    assert open(tmp_path / "special.ply", "rb").read() == ply_ref.ply_bytes(p, c)
    extra_p = np.float32([[1, 2, 3], [4, 5, 6]]); extra_c = np.uint8([[9, 99, 199], [0, 10, 255]])
    w.add_extra_point_cloud(extra_p, extra_c)
    w.write_ply(str(tmp_path / "special2.ply"))
    assert open(tmp_path / "special2.ply", "rb").read() == ply_ref.ply_bytes(np.vstack([p, extra_p]), np.vstack([c, extra_c]))


def test_ply_float64_rows(cuda_device, tmp_path):
    """Genuinely float64 rows (the reference's plane meshes and lines): exact '%f' of a double, every exponent range."""
    rng = np.random.default_rng(11)
    bits = rng.integers(0, 2 ** 63, 60_000, dtype=np.int64).astype(np.uint64)
    expo = rng.integers(1023 - 80, 1023 + 127, 60_000).astype(np.uint64)           # 2^-80 .. 2^127
    v = ((bits & np.uint64((1 << 52) - 1)) | (expo << np.uint64(52))).view(np.float64)
    v[::2] *= -1.0
    v[:8] = [5e-324, 2.2250738585072014e-308, 1e-7, 0.0078125, 0.0234375, 1.7976931348623157e308, np.inf, np.nan]
    v[5] = 1.5 * 2.0 ** 127
    p = v.reshape(-1, 3).copy()
    c = rng.integers(0, 256, p.shape).astype(np.uint8)
    w = PointCloud2Ply(p, c, str(tmp_path / "f64"))
    w.write_ply(str(tmp_path / "f64.ply"))
    assert open(tmp_path / "f64.ply", "rb").read() == ply_ref.ply_bytes(p, c)
    # values of 2^128 and above are refused loudly
    p[0, 0] = 2.0 ** 128
    with pytest.raises(NotImplementedError):
        PointCloud2Ply(p, c, str(tmp_path / "big")).write_ply(str(tmp_path / "big.ply"))


def test_ply_cloud_plus_float64_line(cuda_device, tmp_path):
    """The reference's dump: a float32-born road cloud with the float64 rw line appended (sequence:372-376)."""
    import semantic_depth_lib.pcl as pcl
    rng = np.random.default_rng(2)
    road = (rng.standard_normal((5000, 3)) * [3.0, 0.05, 20.0] + [0.0, -1.5, -30.0]).astype(np.float32).astype(np.float64)
    colors = rng.integers(0, 256, road.shape).astype(np.float64)          # colours are float64 after Open3D (:244)
    left, right = np.array([[-3.7123, -1.52, -9.98]]), np.array([[3.4119, -1.49, -9.97]])
    line, line_colors = pcl.create_3Dline_from_3Dpoints(left.copy(), right.copy(), [250, 0, 0])
    line[:, 2] += 0.2
    w = PointCloud2Ply(road, colors, str(tmp_path / "rw"))
    w.add_extra_point_cloud(line, line_colors)
    w.prepare_and_save_point_cloud()
    want = ply_ref.prepare_and_save_bytes(np.vstack([road, line]), np.vstack([colors, line_colors]))
    assert open(tmp_path / "rw.ply", "rb").read() == want


def test_ply_infinity_filter_with_minus_inf_rows(cuda_device, tmp_path):
    """Disparity-0 pixels reproject to z = -inf: ``z > z.min()`` (point_cloud_2_ply.py:87-89) must drop exactly those rows
    and keep every finite one, including the row at the finite minimum (the filter's minimum is taken over ALL rows)."""
    rng = np.random.default_rng(5)
    p = (rng.standard_normal((4000, 3)) * [3.0, 0.2, 15.0] + [0.0, -1.5, -30.0]).astype(np.float32)
    p[::37, 2] = -np.inf
    p[5, 2] = np.float32(p[np.isfinite(p[:, 2]), 2].min())          # two rows share the finite minimum
    c = rng.integers(0, 256, p.shape).astype(np.uint8)
    w = PointCloud2Ply(p.copy(), c.copy(), str(tmp_path / "inf"))
    w.prepare_and_save_point_cloud()
    got = open(tmp_path / "inf.ply", "rb").read()
    assert got == ply_ref.prepare_and_save_bytes(p, c)
    assert w.points3D.shape[0] == int(np.isfinite(p[:, 2]).sum())
    # no -inf row: the reference drops the rows at the (finite) minimum -- same here
    q = p[np.isfinite(p[:, 2])]
    w2 = PointCloud2Ply(q.copy(), c[: len(q)].copy(), str(tmp_path / "fin"))
    w2.prepare_and_save_point_cloud()
    assert open(tmp_path / "fin.ply", "rb").read() == ply_ref.prepare_and_save_bytes(q, c[: len(q)])
    assert w2.points3D.shape[0] == int((q[:, 2] > q[:, 2].min()).sum())
