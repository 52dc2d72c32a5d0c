"""Extended randomised parity run of the fused path against the oracle (developer tool; the committed tests hold a short
version).  Random shapes, seeds, parameters, degenerate regions (zero / constant disparity bands, missing classes) and both
table layouts (with and without RANSAC hypotheses): every stage count, kept-index list, rw bit-exact.
    python tools/gpu_stress_parity.py [cases] [first_seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from oracle import frame_ref
from semantic_depth_b200 import scene
from semantic_depth_b200.engine import FusionEngine
from semantic_depth_b200.params import FusionParams
from test_gpu_fuse import check_against_oracle


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 777)
    t0 = time.time(); done = 0
    for case in range(cases):
        h = int(rng.integers(8, 40)) * 8; w = int(rng.integers(16, 70)) * 8
        B = int(rng.integers(1, 4))
        P = FusionParams(depth=float(rng.choice([8.0, 10.0, 12.5, 16.0, 25.0])), sor_nb_neighbors=int(rng.choice([3, 5, 10, 16, 20])),
                         sor_std_ratio=float(rng.choice([0.3, 0.5, 1.0])), ror_nb_points=int(rng.choice([10, 20, 80])),
                         ror_radius=float(rng.choice([0.3, 0.5, 0.8])), road_mad_x_thr=float(rng.choice([1.0, 2.0, 3.0])),
                         road_mad_y_thr=float(rng.choice([3.0, 15.0])), fence_mad_y_thr=float(rng.choice([2.0, 5.0])),
                         left_mad_x_thr=float(rng.choice([1.0, 5.0])), right_mad_x_thr=float(rng.choice([1.0, 5.0])))
        frames = []
        for _ in range(B):
            lg, dp, intr = scene.make_frame(h, w, int(rng.integers(0, 100000)))
            kind = rng.integers(0, 8)
            if kind == 0: dp[:, : h // 3] = 0.0                       # -inf depths at the top
            elif kind == 1: dp[:, :, : w // 4] = np.float32(0.02)      # a constant-disparity band
            elif kind == 2: lg[:: 2, 1] = -20.0                        # half of the fence gone
            elif kind == 3: lg[rng.integers(0, h * w, h * w // 50), 0] += 9.0     # salt of road labels anywhere
            frames.append((lg, dp))
        eng = FusionEngine(h, w, max_frames=B, max_hypotheses=64, device="cuda:0")
        dl = torch.from_numpy(np.stack([f[0] for f in frames])).cuda(); dd = torch.from_numpy(np.stack([f[1] for f in frames])).cuda()
        res = eng.fuse_frames(dl, dd, intr, P)
        for f, (lg, dp) in enumerate(frames):
            o = frame_ref.fuse_frame(lg, dp, intr.as_q32(), intr.disparity_mult, P)
            try:
                check_against_oracle(res, f, o, eng)
            except AssertionError as e:
                print(f"case {case} frame {f} shape {h}x{w} FAILED: {str(e)[:400]}"); raise
        eng.close(); done += 1
    print(f"stress parity: {done} cases ({time.time() - t0:.0f} s), all stage counts / index lists / rw equal to the oracle")


if __name__ == "__main__":
    main()
