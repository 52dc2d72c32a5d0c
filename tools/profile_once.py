"""One warm + one measured fused call on a single full-resolution frame batch (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semantic_depth_b200 import scene
from semantic_depth_b200.engine import FusionEngine
from semantic_depth_b200.params import FusionParams

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
H, W = 1024, 2048
lg, dp, intr = scene.make_batch(B, H, W, first_seed=0)
eng = FusionEngine(H, W, max_frames=B)
dl, dd = torch.from_numpy(lg).cuda(), torch.from_numpy(dp).cuda()
for _ in range(2):
    res = eng.fuse_frames(dl, dd, intr, FusionParams())
print(res.rw, res.f2f, res.counts(0))
