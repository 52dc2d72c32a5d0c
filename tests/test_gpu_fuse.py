"""The fused per-frame path against the oracle and the golden fixtures (reference == oracle there)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import frame_ref
from semantic_depth_b200 import scene
from semantic_depth_b200._lib import COUNT_NAMES
from semantic_depth_b200.engine import FusionEngine
from semantic_depth_b200.params import FusionParams
import semantic_depth_lib.pcl as pcl

pytestmark = pytest.mark.gpu

RW_ABS, REL = 1e-3, 1e-4      # north_star: rw / f2f within 1 mm absolute or 1e-4 relative


def close(a, b):
    if a is None or (isinstance(a, float) and np.isnan(a)):
        return b is None or np.isnan(b)
    return abs(a - b) <= max(RW_ABS, REL * abs(b))


def check_against_oracle(res, f, o, eng=None):
    counts = res.counts(f)
    for name, c in o["counts"].items():
        assert counts[name] == c, f"stage {name}: gpu {counts[name]} != oracle {c} (all: {counts} vs {dict(o['counts'])})"
    assert int(res.status[f]) == o["status"], (int(res.status[f]), o["status"])
    assert close(None if np.isnan(res.rw[f]) else float(res.rw[f]), o["rw"]), (res.rw[f], o["rw"])
    if o["rw"] is not None:
        assert float(res.rw[f]) == o["rw"], "rw is a difference of two cloud coordinates: must be bit-exact"
        assert res.raw["xl"][f] == o["xl"] and res.raw["xr"][f] == o["xr"]
    assert close(None if np.isnan(res.f2f[f]) else float(res.f2f[f]), o["f2f"]), (res.f2f[f], o["f2f"])
    for which, key in (("road", "road_coeff"), ("left", "left_coeff"), ("right", "right_coeff")):
        if which in o["coeff"]:
            ref = np.array([o["coeff"][which][k] for k in ("Cx", "Cy", "Cz", "C")])
            np.testing.assert_allclose(res.raw[key][f], ref, rtol=1e-9, atol=1e-10)
    if "fence_mean_x" in o and o["counts"]["fence_abs_z"] > 0:
        assert np.float32(res.raw["fence_mean_x"][f]) == np.float32(o["fence_mean_x"]), "np.mean emulation"
    if eng is not None:
        for which, stage in (("road", "road_ror"), ("left", "left_plane"), ("right", "right_plane")):
            pts, src = eng.final_cloud(f, which)
            assert np.array_equal(src.cpu().numpy(), o["src"][stage]), f"final {which} cloud indices"
        # materialised stages (sd_ws_stage_src) and the filters the fused path keeps as alive bytes (sd_ws_stage_alive)
        for stage in ("road_plane", "fence_abs_z", "fence_mad_y", "left_mad_x", "right_mad_x"):
            src = eng.stage_src(f, stage, counts[stage])
            assert np.array_equal(src.cpu().numpy(), o["src"][stage]), stage


@pytest.mark.parametrize("h,w,seed", [(64, 128, 0), (128, 256, 0), (128, 256, 3), (256, 512, 0), (256, 512, 1)])
def test_fused_matches_oracle(cuda_device, h, w, seed):
    logits, disp, intr = scene.make_frame(h, w, seed)
    P = FusionParams()
    o = frame_ref.fuse_frame(logits, disp, intr.as_q32(), intr.disparity_mult, P)
    eng = FusionEngine(h, w, max_frames=1, device=cuda_device)
    res = eng.fuse_frames(torch.from_numpy(logits[None]).cuda(), torch.from_numpy(disp[None]).cuda(), intr, P)
    check_against_oracle(res, 0, o, eng)


def test_fused_golden_fixtures(cuda_device, golden_dir):
    """Fixtures written by tests/golden/make_golden.py, where the reference's own pcl.py produced them."""
    files = sorted(glob.glob(os.path.join(golden_dir, "frame_*.npz")))
    assert files
    for path in files:
        g = np.load(path)
        h, w, seed = int(g["h"]), int(g["w"]), int(g["seed"])
        logits, disp, intr = scene.make_frame(h, w, seed)
        eng = FusionEngine(h, w, max_frames=1, device=cuda_device)
        res = eng.fuse_frames(torch.from_numpy(logits[None]).cuda(), torch.from_numpy(disp[None]).cuda(), intr)
        counts = res.counts(0)
        for name in COUNT_NAMES:
            assert counts[name] == int(g[f"count/{name}"]), (path, name)
        assert int(res.status[0]) == int(g["status"])
        rw = float(g["rw"])
        assert (np.isnan(rw) and np.isnan(res.rw[0])) or float(res.rw[0]) == rw
        assert abs(float(res.f2f[0]) - float(g["f2f"])) <= max(RW_ABS, REL * float(g["f2f"]))
        for which, stage in (("road", "road_ror"), ("left", "left_plane"), ("right", "right_plane")):
            _, src = eng.final_cloud(0, which)
            src = src.cpu().numpy()
            if f"src/{stage}" in g:
                assert np.array_equal(src, g[f"src/{stage}"]), (path, stage)
            else:
                chk = int(np.sum(src.astype(np.uint64) * (np.arange(src.size, dtype=np.uint64) % 65521 + 1)) % (1 << 63))
                assert chk == int(g[f"srcsum/{stage}"]), (path, stage)
        eng.close()


def test_fused_batch_graph_and_host_path(cuda_device):
    h, w, B = 128, 256, 4
    logits, disp, intr = scene.make_batch(B, h, w, first_seed=20)
    P = FusionParams()
    oracles = [frame_ref.fuse_frame(logits[f], disp[f], intr.as_q32(), intr.disparity_mult, P) for f in range(B)]
    eng = FusionEngine(h, w, max_frames=B, device=cuda_device)
    dl, dd = torch.from_numpy(logits).cuda(), torch.from_numpy(disp).cuda()
    res = eng.fuse_frames(dl, dd, intr, P)
    for f in range(B):
        check_against_oracle(res, f, oracles[f], eng)
    # same call again (workspace self-cleaning), then through a CUDA graph, then through host buffers
    res2 = eng.fuse_frames(dl, dd, intr, P)
    assert res2.raw.tobytes() == res.raw.tobytes(), "fused path must be deterministic run to run"
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        eng.enqueue(dl, dd, intr, P)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            eng.enqueue(dl, dd, intr, P)
        for _ in range(3):
            g.replay()
        res3 = eng.fetch(B)
    assert res3.raw.tobytes() == res.raw.tobytes(), "graph replay must reproduce the eager result"
    res4 = pcl.fuse_frames(logits, disp, intr, P)          # NumPy in: host entry point
    assert res4.raw.tobytes() == res.raw.tobytes()


def test_pipeline_slot_sees_every_camera(cuda_device):
    """A slot's CUDA graph captures the camera by value: batches with different intrinsics through the SAME slot must
    each get their own camera (the reference derives disparity_mult per frame, semantic_depth.py:109,145)."""
    from semantic_depth_b200.params import Intrinsics
    from semantic_depth_b200.stream import FramePipeline
    h, w, B = 128, 256, 2
    logits, disp, intr_a = scene.make_batch(B, h, w, first_seed=31)
    intr_b = Intrinsics(cx=intr_a.cx * 1.02, cy=intr_a.cy * 0.97, f=intr_a.f * 1.1, b=intr_a.b, disparity_mult=intr_a.disparity_mult * 1.25)
    P = FusionParams()
    want = {}
    for tag, intr in (("a", intr_a), ("b", intr_b)):
        want[tag] = [frame_ref.fuse_frame(logits[f], disp[f], intr.as_q32(), intr.disparity_mult, P) for f in range(B)]
    assert want["a"][0]["counts"] != want["b"][0]["counts"]
    pipe = FramePipeline(h, w, B, slots=1, device=cuda_device, params=P, use_graphs=True)
    dl, dd = torch.from_numpy(logits).cuda(), torch.from_numpy(disp).cuda()
    got = []
    for tag, intr in (("a", intr_a), ("b", intr_b), ("a", intr_a), ("b", intr_b)):
        fin = pipe.submit_device(dl, dd, intr, tag=tag)
        if fin:
            got.append(fin)
    got += pipe.drain()
    for tag, intr in (("a", intr_a), ("b", intr_b), ("b", intr_b)):
        fin = pipe.submit_host(logits, disp, intr, tag=tag)
        if fin:
            got.append(fin)
    got += pipe.drain()
    assert [t for t, _ in got] == ["a", "b", "a", "b", "a", "b", "b"]
    for tag, res in got:
        for f in range(B):
            o = want[tag][f]
            assert res.counts(f) == {k: int(v) for k, v in o["counts"].items()}, (tag, f)
            assert (o["rw"] is None and np.isnan(res.rw[f])) or float(res.rw[f]) == o["rw"], (tag, f)
    pipe.close()


def test_pipeline_stream_submission_matches_cached_graph(cuda_device):
    """A stream of batches that each live at their own device address (no per-batch graph capture: eager pixel stage + one
    replayed graph for the rest) gives the answers of the plain cached-graph path, batch for batch, including a short last batch."""
    from semantic_depth_b200.stream import FramePipeline
    h, w, B = 128, 256, 3
    P = FusionParams()
    batches = []
    for k in range(5):
        nfr = B if k < 4 else 2
        lg, dp, intr = scene.make_batch(nfr, h, w, first_seed=50 + k * B)
        batches.append((torch.from_numpy(lg).cuda(), torch.from_numpy(dp).cuda()))
    ref = FramePipeline(h, w, B, slots=1, device=cuda_device, params=P)
    want = []
    for k, (dl, dd) in enumerate(batches):
        ref.submit_device(dl, dd, intr, tag=k)
        want += ref.drain()
    ref.close()
    pipe = FramePipeline(h, w, B, slots=2, device=cuda_device, params=P)
    got = []
    for rep in range(2):                                      # second pass: every slot replays its graph
        for k, (dl, dd) in enumerate(batches):
            fin = pipe.submit_device_stream(dl, dd, intr, tag=k)
            if fin:
                got.append(fin)
        got += pipe.drain()
    pipe.close()
    assert [t for t, _ in got] == list(range(5)) * 2
    for tag, res in got:
        assert res.raw.tobytes() == want[tag][1].raw.tobytes(), tag


@pytest.mark.parametrize("variant", ["rw_only", "no_sor", "no_ror", "no_filters", "depth20", "k16"])
def test_fused_param_variants(cuda_device, variant):
    h, w = 256, 512
    logits, disp, intr = scene.make_frame(h, w, 2)
    P = {"rw_only": FusionParams(approach="rw"), "no_sor": FusionParams(use_sor=False),
         "no_ror": FusionParams(use_ror=False), "no_filters": FusionParams(use_sor=False, use_ror=False),
         "depth20": FusionParams(depth=20.0), "k16": FusionParams(sor_nb_neighbors=16, sor_std_ratio=1.0)}[variant]
    o = frame_ref.fuse_frame(logits, disp, intr.as_q32(), intr.disparity_mult, P)
    eng = FusionEngine(h, w, max_frames=1, device=cuda_device)
    res = eng.fuse_frames(torch.from_numpy(logits[None]).cuda(), torch.from_numpy(disp[None]).cuda(), intr, P)
    counts = res.counts(0)
    for name, c in o["counts"].items():
        assert counts[name] == c, (variant, name, counts, dict(o["counts"]))
    assert int(res.status[0]) == o["status"]
    if o["rw"] is not None:
        assert float(res.rw[0]) == o["rw"]
    if o["f2f"] is not None:
        assert abs(float(res.f2f[0]) - o["f2f"]) <= max(RW_ABS, REL * o["f2f"])
    else:
        assert np.isnan(res.f2f[0])


def test_fused_edge_cases(cuda_device):
    """Empty road, empty fence, constant-disparity (MAD = 0) and empty-slab frames report status bits
    instead of the reference's exceptions, identically to the oracle."""
    h, w = 64, 128
    logits, disp, intr = scene.make_frame(h, w, 0)
    eng = FusionEngine(h, w, max_frames=1, device=cuda_device)
    cases = {}
    lg = logits.copy(); lg[:, 0] = -20.0; cases["no_road"] = (lg, disp)
    lg = logits.copy(); lg[:, 1] = -20.0; cases["no_fence"] = (lg, disp)
    lg = logits.copy(); lg[:, :2] = -20.0; cases["nothing"] = (lg, disp)
    dp = np.full_like(disp, 0.01); cases["flat_disparity"] = (logits, dp)
    dp = disp.copy(); dp[:, : h // 2] = 0.0; cases["zero_disparity_top"] = (logits, dp)
    for name, (lg, dp) in cases.items():
        o = frame_ref.fuse_frame(lg, dp, intr.as_q32(), intr.disparity_mult, FusionParams())
        res = eng.fuse_frames(torch.from_numpy(lg[None]).cuda(), torch.from_numpy(dp[None]).cuda(), intr)
        counts = res.counts(0)
        # "flat_disparity": every point has the same z, the road regression is rank deficient; the reference's lstsq
        # (pcl.py:154) returns the minimum-norm plane and the chain continues -- so does the CUDA path.
        for cname, c in o["counts"].items():
            assert counts[cname] == c, (name, cname, counts, dict(o["counts"]))
        assert int(res.status[0]) == o["status"], (name, int(res.status[0]), o["status"])
        if o["rw"] is None:
            assert np.isnan(res.rw[0]), name
        else:
            assert float(res.rw[0]) == o["rw"], name


def test_fused_randomised_sweep(cuda_device):
    """Randomised shapes / seeds / parameters (seeded): every stage count, index list, rw bit-exact against the oracle.
    Widths that are not multiples of 8 or 32, odd heights, depths where the slab is sparse, k up to 20, and a batch whose
    frames differ -- the cases the fixed-size tests above do not visit."""
    rng = np.random.default_rng(2024)
    shapes = [(72, 132), (100, 260), (136, 300), (190, 404), (250, 500)]
    for case in range(10):
        h, w = shapes[case % len(shapes)]
        seeds = [int(s) for s in rng.integers(0, 10_000, 2)]
        P = FusionParams(depth=float(rng.choice([8.0, 10.0, 12.5, 16.0])),
                         sor_nb_neighbors=int(rng.choice([5, 10, 16, 20])),
                         sor_std_ratio=float(rng.choice([0.3, 0.5, 1.0])),
                         ror_nb_points=int(rng.choice([20, 80])), ror_radius=float(rng.choice([0.3, 0.5])),
                         road_mad_x_thr=float(rng.choice([2.0, 3.0])))
        frames = [scene.make_frame(h, w, s) for s in seeds]
        intr = frames[0][2]
        eng = FusionEngine(h, w, max_frames=2, device=cuda_device)
        lg = torch.from_numpy(np.stack([f[0] for f in frames])).cuda()
        dp = torch.from_numpy(np.stack([f[1] for f in frames])).cuda()
        res = eng.fuse_frames(lg, dp, intr, P)
        for f, (logits, disp, _) in enumerate(frames):
            o = frame_ref.fuse_frame(logits, disp, intr.as_q32(), intr.disparity_mult, P)
            check_against_oracle(res, f, o, eng)
        eng.close()


def test_torch_op_layer_checks_its_arguments(cuda_device):
    """torch.ops.sd_fusion.fuse_frames validates device / dtype / contiguity / shapes in C++ (RuntimeError) and is
    the path FusionEngine.enqueue takes (its answers are the oracle's in every other test of this file)."""
    import ctypes as C
    from semantic_depth_b200 import _lib
    from semantic_depth_b200.engine import camera_struct, params_struct
    h, w = 64, 128
    logits, disp, intr = scene.make_batch(1, h, w, first_seed=0)
    eng = FusionEngine(h, w, max_frames=1)
    ops = _lib.load_ops()
    lg, dp = torch.from_numpy(logits).cuda(), torch.from_numpy(disp).cuda()
    cam, ps = _lib.struct_tensor(camera_struct(intr)), _lib.struct_tensor(params_struct(FusionParams()))
    ws, res = int(eng._ws.value), eng._results

    def call(lg_=lg, dp_=dp, cam_=cam, ps_=ps, ws_=ws, res_=res, hyp=None):
        ops.fuse_frames(lg_, dp_, cam_, ps_, hyp, None, None, ws_, res_)

    call()                                                                        # the good call
    want = eng.fetch(1)
    ref = eng.fuse_frames(lg, dp, intr, FusionParams())
    assert want.raw.tobytes() == ref.raw.tobytes() and want.counts(0)["road_gather"] > 0       # the same answers, byte for byte
    for bad, pattern in [(dict(lg_=lg.double()), "must be Float"), (dict(lg_=lg.cpu()), "CUDA tensor"),
                         (dict(lg_=lg.transpose(1, 2)), "contiguous|must be"), (dict(dp_=dp[:, :, :, ::2]), "contiguous"),
                         (dict(dp_=dp.reshape(1, 2, w, h).contiguous()[:, :1]), r"\[B, 2, H, W\]"),
                         (dict(lg_=lg[:, :-1].contiguous()), "pixels"), (dict(cam_=cam[:-1].clone()), "bytes"),
                         (dict(ps_=ps.cuda()), "CPU uint8"), (dict(ws_=0), "null workspace"), (dict(res_=res[:8]), "bytes"),
                         (dict(res_=res.cpu()), "CUDA tensor"), (dict(hyp=torch.zeros(1, 4, 3, device="cuda")), "must be Int")]:
        with pytest.raises(RuntimeError, match=pattern):
            call(**bad)
    # the library's own refusals surface as RuntimeError with its message (a 2-frame batch into a 1-frame workspace)
    lg2, dp2 = lg.repeat(2, 1, 1), dp.repeat(2, 1, 1, 1)
    res2 = torch.zeros(2 * C.sizeof(_lib.SdFrameResult), dtype=torch.uint8, device="cuda")
    with pytest.raises(RuntimeError, match="sd_fuse_frames failed"):
        ops.fuse_frames(lg2, dp2, cam, ps, None, None, None, ws, res2)
