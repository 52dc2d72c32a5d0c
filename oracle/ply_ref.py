"""ASCII PLY writer of the reference -- TEST INFRASTRUCTURE ONLY (see oracle/__init__).

Restates ``semantic_depth_lib/point_cloud_2_ply.py:33-93`` (``PointCloud2Ply``): the "infinity" filter
``z > z.min()`` (:87-89), the header (:38-49, including the four leading spaces the reference's triple-quoted string
puts in front of every header line after the first) and ``np.savetxt(f, hstack([points, colors]), '%f %f %f %d %d %d')``
(:64-70).  ``tests/golden/make_golden_ply.py`` asserts byte-equality with the real class in the build container.
"""
from __future__ import annotations

import io

import numpy as np

PLY_HEADER = ('''ply
    format ascii 1.0
    element vertex {vertex_count}
    property float x
    property float y
    property float z
    property uchar red
    property uchar green
    property uchar blue
    end_header
    ''')


def infinity_filter(points3D, colors):
    """point_cloud_2_ply.py:87-89: drop the rows at the minimum z (the far clip of the reprojection)."""
    p = np.asarray(points3D).reshape(-1, 3)
    c = np.asarray(colors).reshape(-1, 3)
    keep = p[:, 2] > p[:, 2].min()
    return p[keep], c[keep]


def ply_bytes(points3D, colors) -> bytes:
    """point_cloud_2_ply.py:62-70: header + one '%f %f %f %d %d %d' row per point."""
    p = np.asarray(points3D).reshape(-1, 3)
    c = np.asarray(colors).reshape(-1, 3)
    rows = np.hstack([p, c])
    buf = io.StringIO()
    buf.write(PLY_HEADER.format(vertex_count=len(rows)))
    np.savetxt(buf, rows, '%f %f %f %d %d %d')
    return buf.getvalue().encode("ascii")


def prepare_and_save_bytes(points3D, colors) -> bytes:
    """point_cloud_2_ply.py:83-93 without touching the disk."""
    return ply_bytes(*infinity_filter(points3D, colors))
