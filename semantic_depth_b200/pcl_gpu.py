"""GPU implementation of the ``semantic_depth_lib.pcl`` call surface.

Same function names, positional order, defaults and return arities as the reference module
(/root/reference/semantic_depth_lib/pcl.py:30-331).  Each filter computes its kept row indices with
the CUDA kernels (stable compaction => NumPy order) and applies them to the caller's own arrays, so
points and colors come back with the caller's dtypes, exactly like the reference's fancy indexing.

Inputs may be NumPy arrays (results are NumPy) or CUDA torch tensors (results are CUDA tensors).
fp64 clouds (what Open3D hands back, semantic_depth.py:244) are accepted where the reference's
arithmetic does not depend on the dtype (axis cuts, plane fit, slab end points) provided every value
is exactly representable in fp32; MAD and extract_pcls are fp32-only (their NumPy arithmetic differs
in fp64) and raise NotImplementedError otherwise.

Reference exceptions are reproduced: ValueError for empty clouds where the reference hits
``min()`` / ``np.amin`` (pcl.py:33,107), ``(None, None)`` for an empty slab (pcl.py:303-304),
numpy.linalg.LinAlgError for parallel planes (pcl.py:232).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import SdPredicate
from .engine import FusionEngine

GRID_SIZE = 0.05   # pcl.py:100
_engines: dict = {}


def engine_for(n_points: int, device=None) -> FusionEngine:
    """Cached per-device engine whose workspace holds clouds of at least ``n_points`` points."""
    dev = torch.device(device or "cuda:0")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
    need = max(int(n_points), 1)
    eng = _engines.get(dev)
    if eng is None or eng.capacity < need:
        cap = 1 << 16
        while cap < need:
            cap <<= 1
        if eng is not None:
            eng.close()
        eng = FusionEngine(height=cap // 4096, width=4096, max_frames=1, max_hypotheses=1 << 14, device=dev)
        _engines[dev] = eng
    return eng


class _Cloud:
    """SoA fp32 device view of an [N,3] cloud given as NumPy / torch, remembering how to answer."""

    def __init__(self, points3D, allow_f64: bool = True, what: str = ""):
        self.is_torch = isinstance(points3D, torch.Tensor)
        self.src = points3D
        if self.is_torch:
            if not points3D.is_cuda:
                raise TypeError("torch inputs must live on a CUDA device (NumPy arrays cover the host case)")
            t = points3D
        else:
            arr = np.asarray(points3D)
            if arr.ndim != 2 or arr.shape[1] != 3:
                raise ValueError("points3D must have shape [N, 3]")
            self.src = arr
            t = torch.from_numpy(np.ascontiguousarray(arr)).cuda() if arr.size else torch.zeros((0, 3), dtype=torch.float32, device="cuda")
        if t.ndim != 2 or t.shape[1] != 3:
            raise ValueError("points3D must have shape [N, 3]")
        self.f64 = t.dtype == torch.float64
        if self.f64:
            if not allow_f64:
                raise NotImplementedError(f"{what}: float64 clouds change NumPy's arithmetic here; pass float32")
            t32 = t.to(torch.float32)
            if not bool(torch.all((t32.to(torch.float64) == t) | (t != t))):
                raise NotImplementedError(f"{what}: float64 cloud is not exactly representable in float32")
            t = t32
        elif t.dtype != torch.float32:
            raise TypeError("points3D must be float32 or float64")
        self.n = t.shape[0]
        soa = t.t().contiguous()
        self.x, self.y, self.z = soa[0], soa[1], soa[2]
        self.device = t.device
        self.eng = engine_for(self.n, self.device)

    def col(self, axis):
        return (self.x, self.y, self.z)[axis]

    def take(self, arr, idx: torch.Tensor):
        """arr[idx] in the caller's array kind."""
        if isinstance(arr, torch.Tensor):
            return arr[idx.to(arr.device).long()]
        return np.asarray(arr)[idx.cpu().numpy()]


def _pred(kind, axis=0, **kw) -> SdPredicate:
    p = SdPredicate()
    p.kind, p.axis = kind, axis
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _thr32(c: "_Cloud", thr: float, upper: bool = True) -> float:
    """The fp32 constant whose comparison with the fp32 device column decides like NumPy does.

    fp32 cloud: the Python threshold is a weak scalar and becomes ``np.float32(thr)`` (nearest).  fp64 cloud (the
    reference's clouds after the Open3D round trip, semantic_depth.py:244): NumPy compares in double, and for an fp32-born
    value ``x < thr``  <=>  ``x < (smallest fp32 >= thr)``; ``x > thr``  <=>  ``x > (largest fp32 <= thr)``."""
    t = np.float32(thr)
    if c.f64 and np.isfinite(t):
        if upper and float(t) < float(thr):
            t = np.nextafter(t, np.float32(np.inf))
        elif not upper and float(t) > float(thr):
            t = np.nextafter(t, np.float32(-np.inf))
    return float(t)


def _keep(c: _Cloud, pred: SdPredicate) -> torch.Tensor:
    idx, _ = c.eng.filter(c.x, c.y, c.z, pred, want_points=False)
    return idx


# ----------------------------------------------------------------------------------------------
# reference call surface
# ----------------------------------------------------------------------------------------------
def remove_from_to(points3D, colors, axis, from_meter, to_meter):
    """pcl.py:30-43: keeps rows with ``p[axis] < -to_meter`` (``from_meter`` is unused there too)."""
    c = _Cloud(points3D)
    if c.n == 0:
        raise ValueError("min() arg is an empty sequence")          # pcl.py:33
    k = _keep(c, _pred(_lib.PRED_LT, axis, fa=_thr32(c, float(-to_meter))))
    return c.take(c.src, k), c.take(colors, k)


def mad(points1D):
    """pcl.py:76-81: (abs deviations, their median) of a 1-D fp32 column."""
    is_torch = isinstance(points1D, torch.Tensor)
    col = points1D if is_torch else torch.from_numpy(np.ascontiguousarray(points1D)).cuda()
    if col.dtype != torch.float32:
        raise NotImplementedError("mad: float32 columns only")
    col = col.contiguous()
    med, m = engine_for(col.numel(), col.device).median_mad(col)
    if is_torch:
        return (points1D - float(med)).abs(), m
    return abs(np.asarray(points1D) - med), m


def remove_noise_by_mad(points3D, colors, axis, threshold=15.0):
    """pcl.py:46-73: keep rows with ``0.6745*|v-median|/MAD < threshold`` (all fp32)."""
    c = _Cloud(points3D, allow_f64=False, what="remove_noise_by_mad")
    if c.n == 0:
        return c.take(c.src, torch.zeros(0, dtype=torch.int32, device=c.device)), c.take(colors, torch.zeros(0, dtype=torch.int32, device=c.device))
    med, m = c.eng.median_mad(c.col(axis).contiguous())
    k = _keep(c, _pred(_lib.PRED_MAD, axis, fa=float(threshold), f0=float(med), f1=float(m)))
    return c.take(c.src, k), c.take(colors, k)


_REGRESSORS = {0: (1, 2), 1: (0, 2), 2: (0, 1)}   # pcl.py:118,152,184
_NAMES = ("Cx", "Cy", "Cz")


def _coefficients(axis, C):
    iu, iv = _REGRESSORS[axis]
    d = {_NAMES[axis]: -1.0, _NAMES[iu]: np.float64(C[0]), _NAMES[iv]: np.float64(C[1]), "C": np.float64(C[2])}
    return {k: d[k] for k in ("Cx", "Cy", "Cz", "C")}


def _plane_mesh(c: _Cloud, axis, C, plane_color):
    """The 0.05 m visualisation mesh (pcl.py:107-113,123-126): host/torch glue, not on the hot path."""
    iu, iv = _REGRESSORS[axis]
    if c.is_torch:
        src = c.src
        u, v = src[:, iu], src[:, iv]
        U, V = torch.meshgrid(torch.arange(float(u.min()), float(u.max()), GRID_SIZE, dtype=torch.float64, device=src.device),
                              torch.arange(float(v.min()), float(v.max()), GRID_SIZE, dtype=torch.float64, device=src.device),
                              indexing="xy")
        Wm = C[0] * U + C[1] * V + C[2]
        cols = [None, None, None]
        cols[iu], cols[iv], cols[axis] = U.flatten(), V.flatten(), Wm.flatten()
        plane3D = torch.stack(cols, dim=1)
        return plane3D, torch.ones_like(plane3D) * torch.tensor(plane_color, dtype=torch.float64, device=src.device)
    src = c.src
    u_min, u_max = np.amin(src[:, iu]), np.amax(src[:, iu])
    v_min, v_max = np.amin(src[:, iv]), np.amax(src[:, iv])
    U, V = np.meshgrid(np.arange(u_min, u_max, GRID_SIZE), np.arange(v_min, v_max, GRID_SIZE))
    Wm = C[0] * U + C[1] * V + C[2]
    cols = [None, None, None]
    cols[iu], cols[iv], cols[axis] = U.flatten(), V.flatten(), Wm.flatten()
    plane3D = np.c_[cols[0], cols[1], cols[2]]
    return plane3D, np.ones(plane3D.shape) * plane_color


def remove_noise_by_fitting_plane(points3D, colors, axis=0, threshold=1.0, plane_color=[255, 255, 255],
                                  hypotheses=None):
    """pcl.py:84-209: least-squares plane, keep ``abs(residual) < threshold``; returns the 5-tuple
    (points, colors, plane3D, colors_plane, coefficients).

    ``hypotheses`` (additive; int triplets [K,3]) selects the RANSAC variant of north_star row 8-R:
    score the K planes, refit on the best one's inliers, filter with the refit.  ``None`` is the
    reference behaviour."""
    if axis not in (0, 1, 2):
        raise UnboundLocalError("plane3D")   # the reference falls through its if/elif chain (pcl.py:104-209)
    c = _Cloud(points3D, what="remove_noise_by_fitting_plane")
    if c.n == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")   # pcl.py:107
    if hypotheses is not None:
        trip = hypotheses if isinstance(hypotheses, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(hypotheses, dtype=np.int32))
        trip = trip.to(device=c.device, dtype=torch.int32).contiguous()
        _, best, bc = c.eng.ransac_score(c.x, c.y, c.z, axis, float(threshold), trip)
        inl = _keep(c, _pred(_lib.PRED_PLANE, axis, da=float(threshold), d0=bc[0], d1=bc[1], d2=bc[2])).long()
        C, singular = c.eng.plane_fit(c.x[inl].contiguous(), c.y[inl].contiguous(), c.z[inl].contiguous(), axis)
    else:
        C, singular = c.eng.plane_fit(c.x, c.y, c.z, axis)
    if singular:
        raise np.linalg.LinAlgError("plane fit: rank-deficient cloud (the reference's lstsq would return a minimum-norm "
                                    "solution; the CUDA path reports it instead)")
    plane3D, colors_plane = _plane_mesh(c, axis, C, plane_color)
    k = _keep(c, _pred(_lib.PRED_PLANE, axis, da=float(threshold), d0=C[0], d1=C[1], d2=C[2]))
    return c.take(c.src, k), c.take(colors, k), plane3D, colors_plane, _coefficients(axis, C)


def planes_intersection_at_certain_depth(C_p1, C_p2, z):
    """pcl.py:212-237: host fp64 2x2 solve (``inv(A) @ B``) at ``z = -depth`` -> [[x, y, z]].

    (The reference's array construction at line 235 raises on NumPy >= 1.24; the semantics of lines
    217-233 are kept.)  A singular system raises numpy.linalg.LinAlgError like ``np.linalg.inv``."""
    z = -z
    A = np.array([[C_p1["Cx"], C_p1["Cy"]], [C_p2["Cx"], C_p2["Cy"]]], dtype=np.float64)
    B = np.array([[-(C_p1["Cz"] * z + C_p1["C"])], [-(C_p2["Cz"] * z + C_p2["C"])]], dtype=np.float64)
    X = np.linalg.inv(A) @ B
    return np.array([[X[0, 0], X[1, 0], z]], dtype=np.float64)


def threshold_complete(points3D, colors, axis, threshold=15.0):
    """pcl.py:240-250: keep rows with ``abs(p[axis]) < threshold``."""
    c = _Cloud(points3D)
    k = _keep(c, _pred(_lib.PRED_ABS_LT, axis, fa=_thr32(c, float(threshold))))
    return c.take(c.src, k), c.take(colors, k)


def extract_pcls(points3D, colors, axis=0):
    """pcl.py:253-268: split around ``np.mean`` of the column (NumPy's pairwise fp32 sum reproduced)."""
    c = _Cloud(points3D, allow_f64=False, what="extract_pcls")
    mean = c.eng.mean_f32(c.col(axis).contiguous())
    left = _keep(c, _pred(_lib.PRED_LT, axis, fa=float(mean)))
    right = _keep(c, _pred(_lib.PRED_GT, axis, fa=float(mean)))
    return c.take(c.src, left), c.take(colors, left), c.take(c.src, right), c.take(colors, right)


def get_end_points_of_segment(segment):
    """pcl.py:293-313: rows with the min / max x, or (None, None) for an empty segment."""
    c = _Cloud(segment)
    if c.n == 0:
        return None, None
    xmin, xmax, _ = c.eng.slab_minmax(c.x, c.z, 0.0, 0.0, 2)          # every row, whatever its z (np.amin / np.amax)
    if bool(torch.isnan(c.x).any()):                                  # np.amin / np.amax propagate NaN, and x == NaN selects no row
        xmin = xmax = np.float32(np.nan)
    return _rows_equal(c, xmin), _rows_equal(c, xmax)


def _rows_equal(c: _Cloud, value) -> object:
    # rows whose x equals the extreme: |x| < value+ and > value- cannot be expressed with one predicate;
    # x == v  <=>  not (x < v) and not (x > v); both extremes are attained, so use two cheap filters.
    if value != value:                                                # NaN extreme: `x == nan` selects nothing
        return c.take(c.src, torch.empty(0, dtype=torch.int32, device=c.device))
    lt = _keep(c, _pred(_lib.PRED_LT, 0, fa=float(value)))
    gt = _keep(c, _pred(_lib.PRED_GT, 0, fa=float(value)))
    mask = torch.ones(c.n, dtype=torch.bool, device=c.device)
    mask[lt.long()] = False
    mask[gt.long()] = False
    return c.take(c.src, torch.nonzero(mask).flatten().to(torch.int32))


def get_end_points_of_road(points3D, depth):
    """pcl.py:271-290: end points of the slab ``-(depth+0.05) < z < -(depth-0.05)``.

    Bounds follow NumPy's rule for the cloud's dtype: Python doubles for an fp64 cloud (the
    reference's case, after Open3D), rounded to fp32 for an fp32 cloud."""
    c = _Cloud(points3D)
    lo, hi = -(depth + 0.05), -(depth - 0.05)
    use_f32 = not c.f64
    if use_f32:
        lo, hi = float(np.float32(lo)), float(np.float32(hi))
    xmin, xmax, cnt = c.eng.slab_minmax(c.x, c.z, lo, hi, use_f32)
    if cnt == 0:
        return None, None
    slab = _keep(c, _pred(_lib.PRED_SLAB, 2, use_f32=int(use_f32), f0=lo, f1=hi, d0=lo, d1=hi)).long()
    sx = c.x[slab]
    left = slab[sx == float(xmin)].to(torch.int32)
    right = slab[sx == float(xmax)].to(torch.int32)
    return c.take(c.src, left), c.take(c.src, right)


def compute_distance_in_3D(pt3D_A, pt3D_B):
    """pcl.py:316-318 (host scalar)."""
    if isinstance(pt3D_A, torch.Tensor):
        return torch.linalg.norm(pt3D_A - pt3D_B)
    return np.linalg.norm(pt3D_A - pt3D_B)


def create_3Dline_from_3Dpoints(left_pt, right_pt, color):
    """pcl.py:321-331: 1001-row visualisation line; lifts both end points 1 cm IN PLACE like the
    reference does (viz glue on the host, not on the hot path)."""
    left_pt[0][1] += 0.01
    right_pt[0][1] += 0.01
    if isinstance(left_pt, torch.Tensor):
        v = right_pt - left_pt
        t = torch.arange(0.0, 1.0, 0.001, dtype=torch.float64, device=left_pt.device)
        line = torch.cat([left_pt.to(torch.float64), left_pt + t[:, None] * v], dim=0)
        return line, torch.ones_like(line) * torch.tensor(color, dtype=torch.float64, device=line.device)
    v = right_pt - left_pt
    t = np.arange(0.0, 1.0, 0.001)
    line = np.concatenate([left_pt, left_pt + t[:, None] * v], axis=0)
    return line, np.ones(line.shape) * color


# ----------------------------------------------------------------------------------------------
# additive entry points: what the reference delegates to Open3D / OpenCV / inline NumPy
# ----------------------------------------------------------------------------------------------
def _as_f64(arr):
    if isinstance(arr, torch.Tensor):
        return arr.to(torch.float64)
    return np.asarray(arr, dtype=np.float64)


def statistical_outlier_removal(points3D, colors, nb_neighbors, std_ratio, return_index=False):
    """Open3D ``statistical_outlier_removal`` + ``select_down_sample`` (semantic_depth.py:227-236):
    keep ``0 < mean_knn_dist < mean + std_ratio*std``; like the Open3D round trip the survivors come
    back as float64 (semantic_depth.py:244-245)."""
    c = _Cloud(points3D, what="statistical_outlier_removal")
    if c.n == 0:
        k = torch.zeros(0, dtype=torch.int32, device=c.device)
    else:
        avg, (mean, std, thr) = c.eng.knn_mean_distance(c.x, c.y, c.z, int(nb_neighbors), float(std_ratio))
        k = _keep(c, _pred(_lib.PRED_SOR, 0, da=thr, d_aux=avg.data_ptr()))
    out = _as_f64(c.take(c.src, k)), _as_f64(c.take(colors, k))
    return out + (k,) if return_index else out


def radius_outlier_removal(points3D, colors, nb_points, radius, return_index=False):
    """Open3D ``radius_outlier_removal`` + ``select_down_sample`` (semantic_depth.py:238-245): keep
    points with more than ``nb_points`` points (self included) within ``radius``."""
    c = _Cloud(points3D, what="radius_outlier_removal")
    if c.n == 0:
        k = torch.zeros(0, dtype=torch.int32, device=c.device)
    else:
        cnt = c.eng.radius_count(c.x, c.y, c.z, float(radius), int(nb_points))
        k = _keep(c, _pred(_lib.PRED_ROR, 0, ia=int(nb_points), d_aux=cnt.data_ptr()))
    out = _as_f64(c.take(c.src, k)), _as_f64(c.take(colors, k))
    return out + (k,) if return_index else out
