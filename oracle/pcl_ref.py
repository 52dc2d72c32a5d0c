"""NumPy restatement of ``semantic_depth_lib/pcl.py`` -- TEST INFRASTRUCTURE ONLY (see oracle/__init__).

Same function names, argument order and return arities as the reference module so the parity
tests read like calls into the reference.  Every function is written from the reference's
*behaviour* (vectorised, index-returning helpers underneath) and cites the lines it follows;
``tests/golden/make_golden.py`` asserts bit-equality with the real module in the build container.

dtype rules that matter (NumPy >= 2, NEP 50):
  * Python-float literals are weak: fp32 clouds are compared/multiplied in fp32, fp64 clouds in fp64.
  * ``np.float64`` scalars (lstsq coefficients) are strong: plane residuals are fp64.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

GRID_SIZE = 0.05  # pcl.py:100


# ----------------------------------------------------------------------------------------------
# index-returning kernels of each filter (what the CUDA path is compared with)
# ----------------------------------------------------------------------------------------------
def keep_remove_from_to(points3D, axis, to_meter):
    """pcl.py:35-37: keep rows with ``p[axis] < -to_meter`` (from_meter is ignored, pcl.py:30-43)."""
    return np.flatnonzero(points3D[:, axis] < -to_meter)


def mad(points1D):
    """pcl.py:76-81: (abs deviations from the median, median of those)."""
    with np.errstate(all="ignore"):
        med = np.median(points1D)
        abs_diffs = abs(points1D - med)
        return abs_diffs, np.median(abs_diffs)


def keep_mad(points3D, axis, threshold):
    """pcl.py:56-67: keep rows whose penalty ``0.6745*|v-med|/mad`` is below ``threshold``."""
    if points3D.shape[0] == 0:
        return np.zeros(0, dtype=np.int64)
    abs_diffs, mad_axis = mad(points3D[:, axis])
    with np.errstate(all="ignore"):
        penalty = 0.6745 * abs_diffs / mad_axis
    return np.flatnonzero(penalty < threshold)


_REGRESSORS = {0: (1, 2), 1: (0, 2), 2: (0, 1)}  # pcl.py:118, 152, 184


def fit_plane(points3D, axis):
    """pcl.py:118-120 / 152-154 / 184-186: least squares ``w = C0*u + C1*v + C2`` (lstsq, fp64)."""
    iu, iv = _REGRESSORS[axis]
    A = np.c_[points3D[:, iu], points3D[:, iv], np.ones(points3D.shape[0])]
    b = points3D[:, axis]
    C, _, _, _ = scipy.linalg.lstsq(A, b)
    return C


def plane_residual(points3D, axis, C):
    """pcl.py:130 / 163 / 196: ``((C0*u + C1*v) - w) + C2`` (fp64 because C[k] is np.float64)."""
    iu, iv = _REGRESSORS[axis]
    return C[0] * points3D[:, iu] + C[1] * points3D[:, iv] - points3D[:, axis] + C[2]


def coefficients_dict(axis, C):
    """pcl.py:135 / 168 / 201: reorder into Cx*x + Cy*y + Cz*z + C = 0, regressed axis = -1."""
    names = ("Cx", "Cy", "Cz")
    iu, iv = _REGRESSORS[axis]
    out = {names[axis]: -1.0, names[iu]: C[0], names[iv]: C[1], "C": C[2]}
    return {k: out[k] for k in ("Cx", "Cy", "Cz", "C")}


def keep_plane(points3D, axis, threshold, C=None):
    """Kept row indices of remove_noise_by_fitting_plane and the coefficient vector used."""
    if C is None:
        C = fit_plane(points3D, axis)
    a = plane_residual(points3D, axis, C)
    return np.flatnonzero(abs(a) < threshold), C


def plane_mesh(points3D, axis, C, plane_color):
    """pcl.py:107-113,123-126 (and the axis 1/2 twins): 0.05 m visualisation mesh of the plane."""
    iu, iv = _REGRESSORS[axis]
    u_min, u_max = np.amin(points3D[:, iu]), np.amax(points3D[:, iu])
    v_min, v_max = np.amin(points3D[:, iv]), np.amax(points3D[:, iv])
    U, V = np.meshgrid(np.arange(u_min, u_max, GRID_SIZE), np.arange(v_min, v_max, GRID_SIZE))
    Wm = C[0] * U + C[1] * V + C[2]
    cols = [None, None, None]
    cols[iu], cols[iv], cols[axis] = U.flatten(), V.flatten(), Wm.flatten()
    plane3D = np.c_[cols[0], cols[1], cols[2]]
    return plane3D, np.ones(plane3D.shape) * plane_color


def keep_threshold_complete(points3D, axis, threshold):
    """pcl.py:245-247: keep rows with ``abs(p[axis]) < threshold``."""
    return np.flatnonzero(abs(points3D[:, axis]) < threshold)


def keep_extract_pcls(points3D, axis=0):
    """pcl.py:257-264: split around ``np.mean`` of the column; rows equal to the mean are dropped."""
    col = points3D[:, axis]
    with np.errstate(all="ignore"):
        mean = np.mean(col)
    return np.flatnonzero(col < mean), np.flatnonzero(col > mean), mean


def keep_slab(points3D, depth):
    """pcl.py:280-283: rows with ``-(depth+0.05) < z < -(depth-0.05)`` (strict both sides)."""
    z = points3D[:, 2]
    return np.flatnonzero((z < -(depth - 0.05)) & (z > -(depth + 0.05)))


# ----------------------------------------------------------------------------------------------
# reference call surface (pcl.py:30-331)
# ----------------------------------------------------------------------------------------------
def remove_from_to(points3D, colors, axis, from_meter, to_meter):
    if points3D.shape[0] == 0:
        raise ValueError("min() arg is an empty sequence")  # pcl.py:33 builtin min on empty column
    k = keep_remove_from_to(points3D, axis, to_meter)
    return points3D[k], colors[k]


def remove_noise_by_mad(points3D, colors, axis, threshold=15.0):
    k = keep_mad(points3D, axis, threshold)
    return points3D[k], colors[k]


def remove_noise_by_fitting_plane(points3D, colors, axis=0, threshold=1.0, plane_color=[255, 255, 255]):
    if points3D.shape[0] == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")  # pcl.py:107
    C = fit_plane(points3D, axis)
    plane3D, colors_plane = plane_mesh(points3D, axis, C, plane_color)
    k, _ = keep_plane(points3D, axis, threshold, C)
    return points3D[k], colors[k], plane3D, colors_plane, coefficients_dict(axis, C)


def planes_intersection_at_certain_depth(C_p1, C_p2, z):
    """pcl.py:212-237 with the NumPy>=1.24 shim: the reference builds a ragged array at line 235.

    Semantics kept: ``X = inv(A) @ B`` for the 2x2 system in (x, y) at ``z = -depth``; returns
    ``[[x, y, z]]`` float64 of shape (1, 3).  A singular A raises numpy.linalg.LinAlgError (line 232).
    """
    z = -z
    A = np.array([[C_p1["Cx"], C_p1["Cy"]], [C_p2["Cx"], C_p2["Cy"]]], dtype=np.float64)
    B = np.array([[-(C_p1["Cz"] * z + C_p1["C"])], [-(C_p2["Cz"] * z + C_p2["C"])]], dtype=np.float64)
    X = np.linalg.inv(A) @ B
    return np.array([[X[0, 0], X[1, 0], z]], dtype=np.float64)


def threshold_complete(points3D, colors, axis, threshold=15.0):
    k = keep_threshold_complete(points3D, axis, threshold)
    return points3D[k], colors[k]


def extract_pcls(points3D, colors, axis=0):
    left, right, _ = keep_extract_pcls(points3D, axis)
    return points3D[left], colors[left], points3D[right], colors[right]


def get_end_points_of_segment(segment):
    """pcl.py:293-313: rows holding the min / max x of the slab, or (None, None) if it is empty."""
    seg_x = segment[:, 0]
    if seg_x.size == 0:
        return None, None
    return segment[np.flatnonzero(seg_x == np.amin(seg_x))], segment[np.flatnonzero(seg_x == np.amax(seg_x))]


def get_end_points_of_road(points3D, depth):
    return get_end_points_of_segment(points3D[keep_slab(points3D, depth)])


def compute_distance_in_3D(pt3D_A, pt3D_B):
    """pcl.py:316-318."""
    return np.linalg.norm(pt3D_A - pt3D_B)


def create_3Dline_from_3Dpoints(left_pt, right_pt, color):
    """pcl.py:321-331: 1001-row line; lifts both end points by 1 cm *in place* like the reference."""
    left_pt[0][1] += 0.01
    right_pt[0][1] += 0.01
    v = right_pt - left_pt
    t = np.arange(0.0, 1.0, 0.001)
    line = np.concatenate([left_pt, left_pt + t[:, None] * v], axis=0)
    return line, np.ones(line.shape) * color
