"""BASELINE.json's full sizes (configs 2, 4, 5) where the CPU oracle would need minutes per case.

The oracle-vs-CUDA comparisons of the other test files run at sizes the oracle finishes in seconds; here
the same C-ABI entry points run at 1024x2048 / 2 M points / 16 384 hypotheses and are checked through
size-independent properties and through brute-force restatements of the *same arithmetic* evaluated
with plain torch fp64 ops on the device (separate kernels per operator: no FMA contraction), on sampled
queries:

  * fused path: run-to-run determinism, stable compaction (strictly ascending source indices), stage
    counts never grow, idempotence of every filter (re-applying it to its own output keeps everything),
    rw == |xl - xr|, answers in the physical range of the synthetic scene;
  * statistical filter: bit-exact mean k-NN distance for 4 096 sampled queries against all points;
  * radius filter: exact counts for sampled queries;
  * RANSAC: bit-exact inlier counts for K in {1 024 .. 16 384} on the 1024x2048 road / fence clouds.
"""
import numpy as np
import pytest
import torch

from semantic_depth_b200 import scene
from semantic_depth_b200._lib import PRED_MAD, PRED_PLANE, SdPredicate
from semantic_depth_b200.engine import FusionEngine
from semantic_depth_b200.params import FusionParams
from semantic_depth_b200.pcl_gpu import engine_for

pytestmark = pytest.mark.gpu

H, W = 1024, 2048


@pytest.fixture(scope="module")
def full_frame(cuda_device):
    logits, disp, intr = scene.make_frame(H, W, 0)
    eng = FusionEngine(H, W, max_frames=1, max_hypotheses=1 << 14, device=cuda_device)
    dl, dd = torch.from_numpy(logits[None]).cuda(), torch.from_numpy(disp[None]).cuda()
    res = eng.fuse_frames(dl, dd, intr, FusionParams())
    yield eng, dl, dd, intr, res
    eng.close()


def brute_knn_avg(pts64, q64, k):
    """Open3D's per-point quantity with torch fp64: ((dx*dx + dy*dy) + dz*dz), sqrt, ascending sum from 0, / k."""
    out = []
    for s in range(0, q64.shape[0], 256):
        q = q64[s:s + 256]
        dx = pts64[None, :, 0] - q[:, None, 0]
        dy = pts64[None, :, 1] - q[:, None, 1]
        dz = pts64[None, :, 2] - q[:, None, 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        small = torch.topk(d2, k, dim=1, largest=False, sorted=True).values
        dist = torch.sqrt(small)
        acc = torch.zeros(q.shape[0], dtype=torch.float64, device=q.device)
        for j in range(k):
            acc = acc + dist[:, j]
        out.append(acc / torch.full_like(acc, float(k)))     # tensor / tensor: a true division (a scalar divisor becomes a reciprocal multiply)
    return torch.cat(out)


def brute_radius_count(pts64, q64, r):
    out = []
    for s in range(0, q64.shape[0], 256):
        q = q64[s:s + 256]
        dx = pts64[None, :, 0] - q[:, None, 0]
        dy = pts64[None, :, 1] - q[:, None, 1]
        dz = pts64[None, :, 2] - q[:, None, 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        out.append((d2 <= r * r).sum(dim=1))
    return torch.cat(out)


def test_fullsize_fused_properties(full_frame):
    eng, dl, dd, intr, res = full_frame
    c = res.counts(0)
    assert int(res.status[0]) == 0
    # the synthetic scene: road 7.5 m wide, fences 7.5 m apart
    assert 7.0 < float(res.rw[0]) < 8.2 and 7.3 < float(res.f2f[0]) < 7.7
    assert float(res.rw[0]) == abs(float(res.raw["xl"][0]) - float(res.raw["xr"][0]))
    chain = ["road_gather", "road_z", "road_mad_y", "road_mad_x", "road_plane", "road_sor", "road_ror"]
    assert all(c[a] >= c[b] for a, b in zip(chain, chain[1:])), c
    assert c["fence_gather"] >= c["fence_mad_y"] >= c["fence_abs_z"] >= c["left_split"] + c["right_split"]
    assert c["left_split"] >= c["left_mad_x"] >= c["left_plane"] and c["right_split"] >= c["right_mad_x"] >= c["right_plane"]
    assert c["road_gather"] > 500_000 and c["fence_gather"] > 500_000      # the full-size workload really ran
    # stable compaction: source pixel indices strictly ascending in every retained cloud
    for which in ("road", "left", "right"):
        pts, src = eng.final_cloud(0, which)
        s = src.to(torch.int64)
        assert bool(torch.all(s[1:] > s[:-1])), which
        assert int(s[0]) >= 0 and int(s[-1]) < H * W
    # determinism: a second run gives the same bytes
    res2 = eng.fuse_frames(dl, dd, intr, FusionParams())
    assert res.raw.tobytes() == res2.raw.tobytes()


def test_fullsize_filters_are_idempotent(full_frame):
    eng, dl, dd, intr, res = full_frame
    P = FusionParams()
    road, _ = eng.final_cloud(0, "road")
    left, _ = eng.final_cloud(0, "left")
    e1 = engine_for(H * W)
    # left fence after its plane filter: the same residual test with the same coefficients keeps every point
    x, y, z = (left[:, i].contiguous() for i in range(3))
    Cx, Cy, Cz, C0 = (float(v) for v in res.raw["left_coeff"][0])     # x = Cy*y + Cz*z + C  (Cx = -1)
    pred = SdPredicate(kind=PRED_PLANE, axis=0, da=P.fence_plane_thr, d0=Cy, d1=Cz, d2=C0)
    idx, _ = e1.filter(x, y, z, pred, want_points=False)
    assert idx.numel() == left.shape[0] and Cx == -1.0
    # MAD filter on the road's x column: applying the *recorded* median / MAD again keeps every survivor
    xr, yr, zr = (road[:, i].contiguous() for i in range(3))
    med, mad = float(res.raw["median"][0][1]), float(res.raw["mad"][0][1])
    pred = SdPredicate(kind=PRED_MAD, axis=0, fa=P.road_mad_x_thr, f0=med, f1=mad)
    idx, _ = e1.filter(xr, yr, zr, pred, want_points=False)
    assert idx.numel() == road.shape[0]


def test_fullsize_sor_and_ror_sampled(full_frame):
    eng, dl, dd, intr, res = full_frame
    P = FusionParams()
    c = res.counts(0)
    # the cloud the two filters saw: the road after its plane filter (still in the workspace)
    n = c["road_plane"]
    src_plane = eng.stage_src(0, "road_plane", n).to(torch.int64)
    pix = eng.pixel_stage(dl, dd, intr)
    pts = pix["points"][0][src_plane]                                 # [n, 3] fp32
    x, y, z = (pts[:, i].contiguous() for i in range(3))
    e1 = engine_for(H * W)
    avg, stats = e1.knn_mean_distance(x, y, z, P.sor_nb_neighbors, P.sor_std_ratio)
    g = torch.Generator(device="cpu").manual_seed(0)
    sample = torch.randperm(n, generator=g)[:4096].cuda()
    p64 = pts.to(torch.float64)
    ref = brute_knn_avg(p64, p64[sample], P.sor_nb_neighbors)
    assert torch.equal(avg[sample], ref), int((avg[sample] != ref).sum())
    # cloud statistics and the fused path's survivors agree with the per-point values
    a = avg.cpu().numpy()
    mean, std = a.mean(), a.std(ddof=1)
    assert abs(stats[0] - mean) <= 1e-12 * mean and abs(stats[1] - std) <= 1e-9 * std
    thr = float(res.raw["sor_thr"][0])
    assert abs(thr - stats[2]) <= 1e-15 * abs(thr) + 1e-18
    alive = (avg > 0) & (avg < thr)
    assert int(alive.sum()) == c["road_sor"]
    # radius filter on the statistical survivors
    xa, ya, za = x[alive].contiguous(), y[alive].contiguous(), z[alive].contiguous()
    cnt = e1.radius_count(xa, ya, za, P.ror_radius, cap=P.ror_nb_points)
    assert int((cnt > P.ror_nb_points).sum()) == c["road_ror"]
    pa = torch.stack([xa, ya, za], dim=1).to(torch.float64)
    na = pa.shape[0]
    # sparse tail of the cloud (far field) is where counts are below the cap: sample there and everywhere
    order = torch.argsort(za)                                          # most negative z = farthest first
    sample = torch.cat([order[:1024], torch.randperm(na, generator=g)[:1024].cuda()])
    refc = brute_radius_count(pa, pa[sample], P.ror_radius)
    got = cnt[sample].to(torch.int64)
    assert torch.equal(got, torch.clamp(refc, max=P.ror_nb_points + 1)), int((got != torch.clamp(refc, max=81)).sum())


def test_config4_two_million_point_stress(cuda_device):
    """BASELINE.json configs[3]: 2 M-point synthetic road cloud, k = 16 statistical filter."""
    n, k = 2_000_000, 16
    pts = torch.from_numpy(scene.make_road_cloud(n, seed=0)).cuda()
    x, y, z = (pts[:, i].contiguous() for i in range(3))
    eng = engine_for(n)
    avg, stats = eng.knn_mean_distance(x, y, z, k, 0.5)
    assert avg.numel() == n and bool(torch.all(avg > 0))
    g = torch.Generator(device="cpu").manual_seed(1)
    sample = torch.randperm(n, generator=g)[:2048].cuda()
    p64 = pts.to(torch.float64)
    ref = brute_knn_avg(p64, p64[sample], k)
    assert torch.equal(avg[sample], ref), int((avg[sample] != ref).sum())
    a = avg.cpu().numpy()
    assert abs(stats[0] - a.mean()) <= 1e-12 * a.mean()
    kept = int(((avg > 0) & (avg < stats[2])).sum())
    assert 0.5 * n < kept < n                       # the 2 % uniform outliers (and the sparse tail) go


@pytest.mark.parametrize("K", [1024, 2048, 4096, 8192, 16384])
def test_config5_ransac_sweep_bit_exact(full_frame, K):
    """BASELINE.json configs[4]: road + left/right fence clouds of the 1024x2048 frame, 1k-16k seeded hypotheses."""
    eng, dl, dd, intr, res = full_frame
    c = res.counts(0)
    pix = eng.pixel_stage(dl, dd, intr)
    e1 = engine_for(H * W)
    cases = [("road_plane", 1, 5.0), ("left_mad_x", 0, 1.0), ("right_mad_x", 0, 1.0)]
    for stage, axis, thr in cases:
        n = c[stage] if stage != "road_plane" else c["road_plane"]
        src = eng.stage_src(0, stage, n).to(torch.int64)
        pts = pix["points"][0][src]
        x, y, z = (pts[:, i].contiguous() for i in range(3))
        trip = torch.from_numpy(np.random.default_rng(1234).integers(0, n, (K, 3)).astype(np.int32)).cuda()
        counts, best, coeff = e1.ransac_score(x, y, z, axis, thr, trip)
        # restatement of SURVEY row 8-R with torch fp64 ops, on a subset of hypotheses (every 16th) x all points
        sel = torch.arange(0, K, 16, device="cuda")
        p = pts.to(torch.float64)
        iu, iv = {0: (1, 2), 1: (0, 2), 2: (0, 1)}[axis]
        u, v, w = p[:, iu], p[:, iv], p[:, axis]
        t = trip[sel].to(torch.int64)
        p0, p1, p2 = p[t[:, 0]], p[t[:, 1]], p[t[:, 2]]
        a = p1 - p0
        b = p2 - p0
        au, av, aw = a[:, iu], a[:, iv], a[:, axis]
        bu, bv, bw = b[:, iu], b[:, iv], b[:, axis]
        nu = av * bw - aw * bv
        nv = aw * bu - au * bw
        nw = au * bv - av * bu
        valid = (nw != 0) & (t[:, 0] != t[:, 1]) & (t[:, 0] != t[:, 2]) & (t[:, 1] != t[:, 2])
        C0 = -(nu / nw)
        C1 = -(nv / nw)
        C2 = (p0[:, axis] - C0 * p0[:, iu]) - C1 * p0[:, iv]
        ref = torch.zeros(sel.numel(), dtype=torch.int64, device="cuda")
        for s in range(0, sel.numel(), 64):
            r = (((C0[s:s + 64, None] * u[None, :]) + (C1[s:s + 64, None] * v[None, :])) - w[None, :]) + C2[s:s + 64, None]
            ref[s:s + 64] = (r.abs() < thr).sum(dim=1)
        ref = torch.where(valid, ref, torch.zeros_like(ref))
        got = counts[sel].to(torch.int64)
        assert torch.equal(got, ref), (stage, K, int((got != ref).sum()))
        allc = counts.cpu().numpy()
        assert best == int(np.argmax(allc))            # lowest index on ties


def test_config4_against_oracle_golden(cuda_device, golden_dir):
    """BASELINE.json configs[3] against the CPU oracle's answers for ALL 2 M points (tests/golden/make_golden_config4.py):
    the per-point mean distances are bit-identical (SHA-256 of the fp64 array), so are the kept indices."""
    import hashlib
    import os
    g = np.load(os.path.join(golden_dir, "config4_sor.npz"))
    n, k, ratio, seed = int(g["n"]), int(g["k"]), float(g["ratio"]), int(g["seed"])
    pts = torch.from_numpy(scene.make_road_cloud(n, seed=seed)).cuda()
    x, y, z = (pts[:, i].contiguous() for i in range(3))
    avg, stats = engine_for(n).knn_mean_distance(x, y, z, k, ratio)
    a = avg.cpu().numpy()
    assert a.dtype == np.float64 and a.shape == (n,)
    assert np.array_equal(a[g["sample_idx"]], g["sample_avg"])
    assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest() == g["avg_sha256"].tobytes()
    # cloud statistics: order-independent exact accumulation here, sequential fp64 sums in the oracle
    assert abs(stats[0] - float(g["mean"])) <= 1e-12 * float(g["mean"])
    assert abs(stats[2] - float(g["thr"])) <= 1e-12 * float(g["thr"])
    assert float(g["tie_margin_rel"]) > 1e-9                       # no point close enough to the threshold to flip
    keep = np.flatnonzero((a > 0) & (a < stats[2]))
    assert keep.size == int(g["kept"])
    chk = int(np.sum(keep.astype(np.uint64) * (np.arange(keep.size, dtype=np.uint64) % 65521 + 1)) % (1 << 63))
    assert chk == int(g["kept_checksum"])
