"""developer probe: every refused call of the op layer, one after the other, with progress output"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from semantic_depth_b200 import _lib, scene
from semantic_depth_b200.engine import FusionEngine, camera_struct, params_struct
from semantic_depth_b200.params import FusionParams
h, w = 64, 128
logits, disp, intr = scene.make_batch(1, h, w, first_seed=0)
eng = FusionEngine(h, w, max_frames=1)
ops = _lib.load_ops()
lg, dp = torch.from_numpy(logits).cuda(), torch.from_numpy(disp).cuda()
cam, ps = _lib.struct_tensor(camera_struct(intr)), _lib.struct_tensor(params_struct(FusionParams()))
ws, res = int(eng._ws.value), eng._results
def call(lg_=lg, dp_=dp, cam_=cam, ps_=ps, ws_=ws, res_=res, hyp=None):
    ops.fuse_frames(lg_, dp_, cam_, ps_, hyp, None, None, ws_, res_)
call(); torch.cuda.synchronize(); print("good ok", flush=True)
cases = [("double", dict(lg_=lg.double())), ("cpu", dict(lg_=lg.cpu())), ("transpose", dict(lg_=lg.transpose(1, 2))), ("strided disp", dict(dp_=dp[:, :, :, ::2])),
         ("disp dims", dict(dp_=dp.reshape(1, 2, w, h).contiguous()[:, :1])), ("pixels", dict(lg_=lg[:, :-1].contiguous())), ("cam bytes", dict(cam_=cam[:-1].clone())),
         ("ps cuda", dict(ps_=ps.cuda())), ("ws 0", dict(ws_=0)), ("res small", dict(res_=res[:8])), ("res cpu", dict(res_=res.cpu())),
         ("hyp float", dict(hyp=torch.zeros(1, 4, 3, device="cuda")))]
if len(sys.argv) > 1:
    order = [int(v) for v in sys.argv[1].split(",")]
    cases = [cases[i] for i in order]
for name, c in cases:
    print("case", name, flush=True)
    try:
        call(**c); print("  NO ERROR", flush=True)
    except RuntimeError as e:
        print("  ->", str(e).splitlines()[0][:120], flush=True)
lg2, dp2 = lg.repeat(2, 1, 1), dp.repeat(2, 1, 1, 1)
res2 = torch.zeros(2 * C.sizeof(_lib.SdFrameResult), dtype=torch.uint8, device="cuda")
print("case batch 2", flush=True)
try:
    ops.fuse_frames(lg2, dp2, cam, ps, None, None, None, ws, res2); print("  NO ERROR")
except RuntimeError as e:
    print("  ->", str(e).splitlines()[0][:160], flush=True)
