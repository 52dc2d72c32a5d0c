"""Developer smoke run on a GPU box: staged checks with verbose diagnostics (not a test)."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import frame_ref
from semantic_depth_b200 import scene
from semantic_depth_b200.engine import FusionEngine
from semantic_depth_b200.params import FusionParams

def stage(name):
    print(f"\n=== {name}", flush=True)

def main():
    h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 128)
    print(torch.cuda.get_device_name(0), torch.version.cuda)
    logits, disp, intr = scene.make_frame(h, w, 0)
    eng = FusionEngine(h, w, max_frames=2)
    dl, dd = torch.from_numpy(logits[None]).cuda(), torch.from_numpy(disp[None]).cuda()
    stage("pixel stage")
    out = eng.pixel_stage(dl, dd, intr)
    road, fence = frame_ref.labels_from_logits(logits)
    pp = frame_ref.post_process_disparity(disp)
    pts = frame_ref.reproject_to_3d(pp * np.float32(intr.disparity_mult), intr.as_q32()).reshape(-1, 3)
    lab = out["labels"][0].cpu().numpy()
    print("road mask mismatches", int(((lab & 1) != 0).__ne__(road).sum()), "fence", int(((lab & 2) != 0).__ne__(fence).sum()))
    print("blend mismatches", int((out["disp_pp"][0].cpu().numpy() != pp.reshape(-1)).sum()))
    gp = out["points"][0].cpu().numpy()
    print("points mismatches", int(((gp != pts) & ~(np.isnan(gp) & np.isnan(pts))).sum()), "of", pts.size)
    print("counts gpu", out["counts"][0], "oracle", road.sum(), (pts[road, 2] < -7).sum(), fence.sum())
    stage("fused")
    o = frame_ref.fuse_frame(logits, disp, intr.as_q32(), intr.disparity_mult, FusionParams())
    t = time.time()
    res = eng.fuse_frames(dl, dd, intr, FusionParams())
    print("time", time.time() - t)
    print("gpu   ", res.counts(0))
    print("oracle", dict(o["counts"]))
    print("rw", res.rw[0], o["rw"], "f2f", res.f2f[0], o["f2f"], "status", res.status[0], o["status"])
    print("median", res.raw["median"][0], "mad", res.raw["mad"][0], "mean_x", res.raw["fence_mean_x"][0], o.get("fence_mean_x"))
    print("sor", res.raw["sor_mean"][0], res.raw["sor_std"][0], res.raw["sor_thr"][0], o["sor"]["mean"], o["sor"]["std"], o["sor"]["thr"])
    print("coeff road", res.raw["road_coeff"][0], o["coeff"].get("road"))
    for which, st in (("road", "road_ror"), ("left", "left_plane"), ("right", "right_plane")):
        p, s = eng.final_cloud(0, which)
        print(which, "final src equal:", np.array_equal(s.cpu().numpy(), o["src"][st]))

if __name__ == "__main__":
    try:
        main()
    except Exception:
        traceback.print_exc()
        sys.exit(1)
