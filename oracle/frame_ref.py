"""Per-frame fusion oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__).

Restates the inline NumPy/OpenCV/Open3D steps of ``FrameProcessor.process_frame``
(``/root/reference/semantic_depth.py:98-460``) and chains them with ``pcl_ref`` in exactly the
order and with exactly the constants of lines 183-324.  Every cloud carries the flat source pixel
index of each point so that compaction indices can be compared stage by stage.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree

from . import pcl_ref

# status bits: same numeric values as semantic_depth_b200/params.py (checked by a test)
EMPTY_ROAD, EMPTY_FENCE_LEFT, EMPTY_FENCE_RIGHT, MAD_ZERO = 1, 2, 4, 8
NO_SLAB_POINTS, SINGULAR_PLANES, EMPTY_FENCE, SINGULAR_FIT = 16, 32, 64, 128


# ----------------------------------------------------------------------------------------------
# pixel stage
# ----------------------------------------------------------------------------------------------
def labels_from_logits(logits, prob_thr=0.5):
    """semantic_depth.py:550-556,563-564: ``softmax(logits)[:, c] > 0.5`` for c = road(0), fence(1).

    The reference evaluates the softmax in a TF1 fp32 session (irreproducible here); the contract
     is an fp64 softmax on both sides.  Returns two flat bool arrays.
    """
    l = np.asarray(logits, dtype=np.float64).reshape(-1, 3)
    e = np.exp(l - l.max(axis=1, keepdims=True))
    p = e / e.sum(axis=1, keepdims=True)
    return p[:, 0] > prob_thr, p[:, 1] > prob_thr


def upsample_scores(scores, weights, bias):
    """FCN-8s' last layer (fcn8s/fcn.py:207-213): ``conv2d_transpose(second_skip, 3, 16x16, stride 8, 'same')``.

    ``scores`` [h, w, 3] fp32 (``second_skip``), ``weights`` [16, 16, 3(out), 3(in)] fp32 (TF layout
    kh, kw, out, in), ``bias`` [3] fp32  ->  logits [8h*8w, 3] fp32, the tensor the reference fetches as
    ``logits:0`` (fcn.py:241).  Output pixel (y, x) receives the taps ``ky = y + 4 - 8*iy`` in [0, 16): two
    low-res rows and two low-res columns (zero outside the map), three input channels each.  TF's own
    summation order is not reproducible; the contract is fp32, no FMA, accumulated
    from 0.0 in the order (iy ascending, ix ascending, ci ascending), bias added last.
    """
    s = np.asarray(scores, dtype=np.float32)
    w = np.asarray(weights, dtype=np.float32)
    b = np.asarray(bias, dtype=np.float32)
    h, wd, _ = s.shape
    H, W = 8 * h, 8 * wd
    y = np.arange(H); x = np.arange(W)
    iy0 = (y + 4) // 8 - 1; ix0 = (x + 4) // 8 - 1
    pad = np.zeros((h + 2, wd + 2, 3), dtype=np.float32)
    pad[1:-1, 1:-1] = s
    out = np.empty((H, W, 3), dtype=np.float32)
    for co in range(3):
        acc = np.zeros((H, W), dtype=np.float32)
        for dy in range(2):
            iy = iy0 + dy
            ky = y + 4 - 8 * iy                      # in [0, 16)
            for dx in range(2):
                ix = ix0 + dx
                kx = x + 4 - 8 * ix
                for ci in range(3):
                    sv = pad[iy[:, None] + 1, ix[None, :] + 1, ci]
                    wv = w[ky[:, None], kx[None, :], co, ci]
                    acc = acc + sv * wv               # fp32 product, fp32 sum
        out[:, :, co] = acc + b[co]
    return out.reshape(H * W, 3)


def _cubic_tables(src_n, dst_n):
    """Tap indices [dst,4] (border-replicated) and 11-bit fixed-point weights [dst,4] of cv2.resize(INTER_CUBIC)."""
    scale = float(src_n) / dst_n
    d = np.arange(dst_n, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    x = (f - s.astype(np.float32)).astype(np.float32)
    one, A = np.float32(1), np.float32(-0.75)
    c0 = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A
    c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
    xm = one - x
    c2 = ((A + np.float32(2)) * xm - (A + np.float32(3))) * xm * xm + one
    c3 = one - c0 - c1 - c2
    co = np.stack([c0, c1, c2, c3], axis=-1).astype(np.float32)
    ico = np.clip(np.rint(co * np.float32(2048)), -32768, 32767).astype(np.int64)   # saturate_cast<short>(c * 2^11)
    idx = np.clip(s[:, None] - 1 + np.arange(4)[None, :], 0, src_n - 1)
    return idx, ico


def resize_cubic_u8(img, dst_w, dst_h):
    """``cv2.resize(frame, (w, h), interpolation=cv2.INTER_CUBIC)`` on uint8 (semantic_depth.py:110-112).

    OpenCV's own implementation (imgproc/src/resize.cpp, 4.13, the code that runs with ``cv2.ipp.setUseIPP(False)``):
    bicubic weights (A = -0.75) evaluated in fp32 and rounded to 11 fractional bits (``saturate_cast<short>``, no
    sum fix-up), an integer horizontal pass into a row buffer of ``dst_w * channels`` ints, and a vertical pass that is

    * fp32 for the first ``8 * floor(dst_w * channels / 8)`` elements of every row (``VResizeCubicVec_32s8u``, 8 x int16
      lanes at the SSE baseline the wheel is built for): ``S0*b0 + (S1*b1 + (S2*b2 + S3*b3))`` with ``b = beta * 2^-22``,
      every product and sum rounded to fp32 (no FMA), round-half-even, saturation;
    * integer for the tail of the row: ``(sum + 2^21) >> 22``, saturation.

    Pinned byte for byte against cv2 with IPP off (tests/golden/make_golden_resize.py).  With IPP on (the wheel's
    default) cv2 dispatches to Intel's closed-source ippiResizeCubic, which differs from OpenCV's own code by 1 LSB in
    up to 4 % of the bytes and depends on the CPU it runs on: that path has no definition to be equal to.
    """
    img = np.asarray(img, dtype=np.uint8)
    if img.ndim == 2:
        img = img[:, :, None]
    h, w, c = img.shape
    xi, xa = _cubic_tables(w, dst_w)
    yi, yb = _cubic_tables(h, dst_h)
    S = img.astype(np.int64)
    hor = np.zeros((h, dst_w, c), np.int64)
    for k in range(4):
        hor += S[:, xi[:, k], :] * xa[None, :, k, None]
    n = dst_w * c
    hor = hor.reshape(h, n)
    out = np.zeros((dst_h, n), np.int64)
    for k in range(4):
        out += hor[yi[:, k], :] * yb[:, k, None]
    out = np.clip((out + (1 << 21)) >> 22, 0, 255)
    nv = (n // 8) * 8
    if nv:
        b = yb.astype(np.float32) * np.float32(2.0 ** -22)                 # exact
        rows = [hor[yi[:, k], :nv].astype(np.float32) for k in range(4)]   # |values| < 2^24: exact
        acc = rows[3] * b[:, 3, None]
        for k in (2, 1, 0):
            acc = (rows[k] * b[:, k, None]).astype(np.float32) + acc       # fp32 product, fp32 sum
        out[:, :nv] = np.clip(np.rint(acc.astype(np.float32)), 0, 255).astype(np.int64)
    return out.astype(np.uint8).reshape(dst_h, dst_w, c)


def labels_argmax(logits):
    """north_star's wording of the labelling: ``argmax`` over the three classes (first maximum wins, like
    ``np.argmax``); road = class 0, fence = class 1.  An alternative to the reference's ``softmax > 0.5``."""
    l = np.asarray(logits, dtype=np.float32).reshape(-1, 3)
    a = np.argmax(l, axis=1)
    return a == 0, a == 1


def label_margin(logits):
    """|p_c - 0.5| of the closest class per pixel; pixels below ~1e-12 are the documented tie class."""
    l = np.asarray(logits, dtype=np.float64).reshape(-1, 3)
    e = np.exp(l - l.max(axis=1, keepdims=True))
    p = e / e.sum(axis=1, keepdims=True)
    return np.abs(p[:, :2] - 0.5).min(axis=1)


def post_process_disparity(disp):
    """semantic_depth.py:656-664 + the cast at 676: monodepth left/flipped blend -> fp32 [H,W].

    ``m`` stays fp32 (0.5 is a weak scalar); the ramps are fp64, so the blend is evaluated in fp64
    as ``((rm*l) + (lm*r)) + (((1-lm)-rm)*m)`` and rounded once.
    """
    _, h, w = disp.shape
    l_disp = disp[0]
    r_disp = disp[1][:, ::-1]
    m_disp = 0.5 * (l_disp + r_disp)
    ramp = np.linspace(0, 1, w)
    l_mask = (1.0 - np.clip(20 * (ramp - 0.05), 0, 1))[None, :]
    r_mask = l_mask[:, ::-1]
    return (r_mask * l_disp + l_mask * r_disp + (1.0 - l_mask - r_mask) * m_disp).astype(np.float32)


def blend_ramps(w):
    """fp64 (l_mask, r_mask) row vectors of post_processing (semantic_depth.py:661-663)."""
    ramp = np.linspace(0, 1, w)
    l_mask = 1.0 - np.clip(20 * (ramp - 0.05), 0, 1)
    return l_mask, l_mask[::-1].copy()


def reproject_to_3d(disp, q32):
    """semantic_depth.py:691-696: cv2.reprojectImageTo3D(disp, Q) restated in NumPy.

    ``q32`` = float32 [-cx, cy, -f, 1/b] (the non-trivial entries of the reference's float32 Q).
    OpenCV evaluates ``[X Y Z W]^T = Q [u v d 1]^T`` in fp64 and stores ``(X,Y,Z)/W`` as fp32;
    verified bit-identical to cv2 4.13 on 1e9 values (division and reciprocal-multiply agree).
    """
    q = np.asarray(q32, dtype=np.float32).astype(np.float64)
    h, w = disp.shape
    d = disp.astype(np.float64)
    u = np.arange(w, dtype=np.float64)[None, :]
    v = np.arange(h, dtype=np.float64)[:, None]
    with np.errstate(all="ignore"):
        wp = q[3] * d
        out = np.empty((h, w, 3), dtype=np.float32)
        out[..., 0] = (u + q[0]) / wp
        out[..., 1] = (q[1] - v) / wp
        out[..., 2] = q[2] / wp
    return out


# ----------------------------------------------------------------------------------------------
# Open3D stand-ins (parity unpinned: Open3D is not vendored by the reference nor installable here)
# ----------------------------------------------------------------------------------------------
def knn_mean_distances(points, nb_neighbors, workers=1):
    """Mean of the distances to the k nearest neighbours (self included, distance 0), fp64.

    Open3D <= 0.7 ``RemoveStatisticalOutliers``: FLANN KNN on the fp64 cloud returns squared
    distances in ascending order; each is sqrt'ed and they are summed by ``std::accumulate`` from
    0.0, i.e. sequentially in ascending order, then divided by the number found.  The squared
    distance is ``((dx*dx + dy*dy) + dz*dz)`` without FMA (FLANN L2 remainder loop; scipy's
    ``sqeuclidean_distance_double`` has the same order).  The neighbour *set* comes from cKDTree;
    the arithmetic is redone here so that it does not depend on how scipy was compiled.
    """
    p = np.ascontiguousarray(points, dtype=np.float64)
    n = p.shape[0]
    k = min(int(nb_neighbors), n)
    if n == 0 or k == 0:
        return np.zeros(n, dtype=np.float64), k
    _, idx = cKDTree(p).query(p, k=k, workers=workers)
    idx = idx.reshape(n, k)
    diff = p[idx] - p[:, None, :]
    d2 = (diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]) + diff[..., 2] * diff[..., 2]
    d = np.sort(np.sqrt(d2), axis=1)
    acc = np.zeros(n, dtype=np.float64)
    for j in range(k):
        acc = acc + d[:, j]
    return acc / k, k


def sor_threshold(avg, std_ratio):
    """Open3D: mean over avg>0 divided by the number of valid points, Bessel std, mean+ratio*std."""
    n = avg.shape[0]
    pos = avg[avg > 0]
    cloud_mean = (np.add.accumulate(pos)[-1] if pos.size else 0.0) / n
    dev = pos - cloud_mean
    sq = dev * dev
    sq_sum = np.add.accumulate(sq)[-1] if sq.size else 0.0
    with np.errstate(all="ignore"):
        std_dev = np.sqrt(sq_sum / (n - 1)) if n > 1 else np.nan
    return cloud_mean + std_ratio * std_dev, cloud_mean, std_dev


def keep_statistical_outlier_removal(points, nb_neighbors, std_ratio, workers=1):
    """Kept indices (ascending) of Open3D statistical_outlier_removal; semantic_depth.py:234-236."""
    n = np.asarray(points).shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0), (np.nan, np.nan, np.nan)
    avg, _ = knn_mean_distances(points, nb_neighbors, workers)
    thr, mu, sd = sor_threshold(avg, std_ratio)
    return np.flatnonzero((avg > 0) & (avg < thr)), avg, (thr, mu, sd)


def radius_counts(points, radius, workers=1):
    """Number of points within ``radius`` (self included; boundary d == r counted, cKDTree rule)."""
    p = np.ascontiguousarray(points, dtype=np.float64)
    if p.shape[0] == 0:
        return np.zeros(0, dtype=np.int64)
    return np.asarray(cKDTree(p).query_ball_point(p, radius, return_length=True, workers=workers), dtype=np.int64)


def keep_radius_outlier_removal(points, nb_points, radius, workers=1):
    """Kept indices of Open3D radius_outlier_removal: count > nb_points; semantic_depth.py:238-241."""
    return np.flatnonzero(radius_counts(points, radius, workers) > nb_points)


def statistical_outlier_removal(points, colors, nb_neighbors, std_ratio):
    """Open3D round trip of semantic_depth.py:227-236,244-245: returns fp64 points and colors."""
    k, _, _ = keep_statistical_outlier_removal(points, nb_neighbors, std_ratio)
    return np.asarray(points, dtype=np.float64)[k], np.asarray(colors, dtype=np.float64)[k]


def radius_outlier_removal(points, colors, nb_points, radius):
    k = keep_radius_outlier_removal(points, nb_points, radius)
    return np.asarray(points, dtype=np.float64)[k], np.asarray(colors, dtype=np.float64)[k]


# ----------------------------------------------------------------------------------------------
# RANSAC variant of the plane fit (north_star / SURVEY.md row 8-R; no reference code exists)
# ----------------------------------------------------------------------------------------------
def ransac_hypothesis_planes(points3D, axis, triplets):
    """fp64 (C0, C1, C2, valid) of the plane through each index triplet, in regression form.

    ``a = p1-p0``, ``b = p2-p0``, ``n = a x b`` with each component ``a_i*b_j - a_j*b_i`` (two
    products, one subtraction, no FMA); ``C0 = -n_u/n_w``, ``C1 = -n_v/n_w``,
    ``C2 = (w0 - C0*u0) - C1*v0``.  Invalid (count forced to 0) when an index repeats or n_w == 0.
    """
    iu, iv = pcl_ref._REGRESSORS[axis]
    p = np.asarray(points3D, dtype=np.float64)
    t = np.asarray(triplets, dtype=np.int64)
    p0, p1, p2 = p[t[:, 0]], p[t[:, 1]], p[t[:, 2]]
    a, b = p1 - p0, p2 - p0
    au, av, aw = a[:, iu], a[:, iv], a[:, axis]
    bu, bv, bw = b[:, iu], b[:, iv], b[:, axis]
    # normal in (u, v, w) coordinates
    n_u = av * bw - aw * bv
    n_v = aw * bu - au * bw
    n_w = au * bv - av * bu
    valid = (t[:, 0] != t[:, 1]) & (t[:, 0] != t[:, 2]) & (t[:, 1] != t[:, 2]) & (n_w != 0)
    with np.errstate(all="ignore"):
        c0 = -n_u / n_w
        c1 = -n_v / n_w
        c2 = (p0[:, axis] - c0 * p0[:, iu]) - c1 * p0[:, iv]
    return c0, c1, c2, valid


def ransac_inlier_counts(points3D, axis, threshold, triplets, chunk=64):
    """Inlier count of every hypothesis: #{abs((((C0*u)+(C1*v))-w)+C2) < thr}, fp64, no FMA."""
    iu, iv = pcl_ref._REGRESSORS[axis]
    p = np.asarray(points3D, dtype=np.float64)
    u, v, w = p[:, iu], p[:, iv], p[:, axis]
    c0, c1, c2, valid = ransac_hypothesis_planes(points3D, axis, triplets)
    counts = np.zeros(len(c0), dtype=np.int64)
    for s in range(0, len(c0), chunk):
        e = min(s + chunk, len(c0))
        with np.errstate(all="ignore"):
            r = ((c0[s:e, None] * u[None, :] + c1[s:e, None] * v[None, :]) - w[None, :]) + c2[s:e, None]
            counts[s:e] = (np.abs(r) < threshold).sum(axis=1)
    counts[~valid] = 0
    return counts


def keep_plane_ransac(points3D, axis, threshold, triplets):
    """best hypothesis (max count, lowest index on ties) -> refit (pcl.py lstsq) on its inliers ->
    final residual filter with the refit coefficients.  Returns (kept idx, C_refit, best, counts)."""
    counts = ransac_inlier_counts(points3D, axis, threshold, triplets)
    best = int(np.argmax(counts))
    c0, c1, c2, _ = ransac_hypothesis_planes(points3D, axis, np.asarray(triplets)[best:best + 1])
    p64 = np.asarray(points3D, dtype=np.float64)
    iu, iv = pcl_ref._REGRESSORS[axis]
    with np.errstate(all="ignore"):
        r = ((c0[0] * p64[:, iu] + c1[0] * p64[:, iv]) - p64[:, axis]) + c2[0]
    inl = np.flatnonzero(np.abs(r) < threshold)
    C = pcl_ref.fit_plane(points3D[inl], axis)
    keep, _ = pcl_ref.keep_plane(points3D, axis, threshold, C)
    return keep, C, best, counts


# ----------------------------------------------------------------------------------------------
# the whole per-frame path (semantic_depth.py:183-324)
# ----------------------------------------------------------------------------------------------
class _Cloud:
    __slots__ = ("pts", "src")

    def __init__(self, pts, src):
        self.pts, self.src = pts, src

    def take(self, keep):
        return _Cloud(self.pts[keep], self.src[keep])

    def __len__(self):
        return self.pts.shape[0]


def fuse_frame(logits, disp, q32, disparity_mult, params=None, hypotheses=None, keep_clouds=False, workers=1):
    """Run the reference's fusion section on one frame and return every observable of it.

    ``params`` is any object with the attribute names of ``semantic_depth_b200.params.FusionParams``
    (duck-typed so the oracle does not import the product); ``None`` = the reference's literals.
    ``hypotheses`` = None (reference behaviour: all-points least squares) or a dict
    ``{'road'|'left'|'right': int triplets [K,3]}`` selecting the RANSAC variant (row 8-R).

    Returns a dict: ``rw``, ``f2f`` (float or None), ``status`` bitfield, ``counts`` (ordered dict of
    per-stage point counts), ``src`` (per-stage flat pixel indices), plane coefficients, xl/xr.
    Errors the reference would raise (empty cloud into np.amin, None end points, singular 2x2)
    are reported through ``status`` exactly like the fused CUDA path does.
    """
    P = params
    g = lambda name, default: getattr(P, name, default) if P is not None else default
    hyp = hypotheses or {}
    h, w = disp.shape[1:]
    out = {"rw": None, "f2f": None, "status": 0, "counts": {}, "src": {}, "coeff": {}}
    counts, src = out["counts"], out["src"]

    road_mask, fence_mask = labels_from_logits(logits, g("prob_thr", 0.5))          # :555-556,563-564
    disp_pp = post_process_disparity(disp)                                          # :676
    disp_px = disp_pp * np.float32(disparity_mult)                                  # :145 (fp32 product)
    points3D = reproject_to_3d(disp_px, q32).reshape(-1, 3)                         # :160
    out["labels"] = (road_mask.astype(np.uint8) | (fence_mask.astype(np.uint8) << 1))
    road = _Cloud(points3D[road_mask], np.flatnonzero(road_mask))                   # :183
    fence = _Cloud(points3D[fence_mask], np.flatnonzero(fence_mask))                # :186
    counts["road_gather"], counts["fence_gather"] = len(road), len(fence)
    src["road_gather"], src["fence_gather"] = road.src, fence.src

    def stage(cloud, keep, name):
        c = cloud.take(keep)
        counts[name], src[name] = len(c), c.src
        return c

    def mad_stage(cloud, axis, thr, name):
        if len(cloud):
            _, m = pcl_ref.mad(cloud.pts[:, axis])
            if not (m > 0):
                out["status"] |= MAD_ZERO
        return stage(cloud, pcl_ref.keep_mad(cloud.pts, axis, thr), name)

    def plane_stage(cloud, axis, thr, name, which, empty_bit):
        if len(cloud) == 0:
            out["status"] |= empty_bit
            counts[name], src[name] = 0, cloud.src
            return cloud, None
        if which in hyp and hyp[which] is not None:
            keep, C, best, hc = keep_plane_ransac(cloud.pts, axis, thr, hyp[which])
            out.setdefault("ransac", {})[which] = {"best": best, "counts": hc}
        else:
            keep, C = pcl_ref.keep_plane(cloud.pts, axis, thr)
        if not np.all(np.isfinite(C)):
            out["status"] |= SINGULAR_FIT
        out["coeff"][which] = pcl_ref.coefficients_dict(axis, C)
        return stage(cloud, keep, name), C

    # ---- road chain: :206 -> :209 -> :212 -> :215-219 -> :234-236 -> :238-241
    road = stage(road, pcl_ref.keep_remove_from_to(road.pts, 2, g("road_z_to_meter", 7.0)), "road_z")
    road = mad_stage(road, 1, g("road_mad_y_thr", 15.0), "road_mad_y")
    road = mad_stage(road, 0, g("road_mad_x_thr", 2.0), "road_mad_x")
    road, _ = plane_stage(road, 1, g("road_plane_thr", 5.0), "road_plane", "road", EMPTY_ROAD)
    if g("use_sor", True):
        keep, avg, thr3 = keep_statistical_outlier_removal(road.pts, g("sor_nb_neighbors", 10), g("sor_std_ratio", 0.5), workers)
        out["sor"] = {"avg": avg, "thr": thr3[0], "mean": thr3[1], "std": thr3[2]}
        road = stage(road, keep, "road_sor")
    if g("use_ror", True):
        road = stage(road, keep_radius_outlier_removal(road.pts, g("ror_nb_points", 80), g("ror_radius", 0.5), workers), "road_ror")
    road = _Cloud(road.pts.astype(np.float64), road.src)                            # :244 Open3D -> fp64
    if len(road) == 0:
        out["status"] |= EMPTY_ROAD

    # ---- rw: :254-259
    depth = g("depth", 10.0)
    d_rw = depth - g("rw_depth_offset", 0.02)
    slab = pcl_ref.keep_slab(road.pts, d_rw) if g("slab_half_width", 0.05) == 0.05 else None
    if slab is None:
        hw_ = g("slab_half_width", 0.05)
        z = road.pts[:, 2]
        slab = np.flatnonzero((z < -(d_rw - hw_)) & (z > -(d_rw + hw_)))
    counts["road_slab"] = int(slab.size)
    if slab.size == 0:
        out["status"] |= NO_SLAB_POINTS
    else:
        xs = road.pts[slab, 0]
        out["xl"], out["xr"] = float(np.amin(xs)), float(np.amax(xs))
        out["rw"] = float(abs(out["xl"] - out["xr"]))                               # :259

    # ---- fence chain: :279 -> :283 -> :286 -> {:291 -> :294-298} | {:302 -> :305-309}
    if g("approach", "both") == "both":
        fence = mad_stage(fence, 1, g("fence_mad_y_thr", 5.0), "fence_mad_y")
        fence = stage(fence, pcl_ref.keep_threshold_complete(fence.pts, 2, g("fence_abs_z_thr", 35.0)), "fence_abs_z")
        if len(fence) == 0:
            out["status"] |= EMPTY_FENCE
        kl, kr, mean_x = pcl_ref.keep_extract_pcls(fence.pts)
        out["fence_mean_x"] = mean_x
        left, right = stage(fence, kl, "left_split"), stage(fence, kr, "right_split")
        left = mad_stage(left, 0, g("left_mad_x_thr", 5.0), "left_mad_x")
        left, _ = plane_stage(left, 0, g("fence_plane_thr", 1.0), "left_plane", "left", EMPTY_FENCE_LEFT)
        right = mad_stage(right, 0, g("right_mad_x_thr", 1.0), "right_mad_x")
        right, _ = plane_stage(right, 0, g("fence_plane_thr", 1.0), "right_plane", "right", EMPTY_FENCE_RIGHT)
        # ---- f2f: :317-324
        cf = out["coeff"]
        if all(k in cf for k in ("road", "left", "right")) and not (out["status"] & SINGULAR_FIT):
            try:
                pl = pcl_ref.planes_intersection_at_certain_depth(cf["road"], cf["left"], depth)
                pr = pcl_ref.planes_intersection_at_certain_depth(cf["road"], cf["right"], depth)
                if np.all(np.isfinite(pl)) and np.all(np.isfinite(pr)):
                    out["left_pt_f2f"], out["right_pt_f2f"] = pl, pr
                    out["f2f"] = float(pcl_ref.compute_distance_in_3D(pl, pr))
                else:
                    out["status"] |= SINGULAR_PLANES
            except np.linalg.LinAlgError:
                out["status"] |= SINGULAR_PLANES
    if keep_clouds:
        out["road_cloud"] = road.pts
    return out
