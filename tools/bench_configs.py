"""Secondary measurements (BASELINE.json configs 4 and 5): one JSON line each.

  config 4  statistical filter on the 2 M-point synthetic road cloud, k = 16: points/s
  config 5  RANSAC scoring, K in {1k..16k} hypotheses on a 450k-point road cloud: point-hypothesis tests/s
  resize    cv2.INTER_CUBIC 1024x2048x3 -> 256x512x3 (semantic_depth.py:110-112): frames/s and GB/s
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from semantic_depth_b200 import scene
from semantic_depth_b200.pcl_gpu import engine_for
import semantic_depth_lib.pcl as pcl


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


def main():
    out = []
    n = 2_000_000
    pts = torch.from_numpy(scene.make_road_cloud(n, seed=0)).cuda()
    x, y, z = (pts[:, i].contiguous() for i in range(3))
    eng = engine_for(n)
    ms = timed(lambda: eng.knn_mean_distance(x, y, z, 16, 0.5), 5)
    out.append({"config": "4: 2M-point road cloud, statistical filter k=16 (grid build + k-NN + cloud statistics, one host sync)",
                "ms": ms, "points_per_s": n / (ms * 1e-3)})
    m = 450_000
    road = torch.from_numpy(scene.make_road_cloud(m, seed=1)).cuda()
    rx, ry, rz = (road[:, i].contiguous() for i in range(3))
    eng2 = engine_for(m)
    for K in (1024, 4096, 16384):
        trip = torch.from_numpy(np.random.default_rng(1234).integers(0, m, (K, 3)).astype(np.int32)).cuda()
        ms = timed(lambda: eng2.ransac_score(rx, ry, rz, 1, 5.0, trip), 3)
        out.append({"config": f"5: RANSAC scoring, {K} hypotheses x {m} points (fp64, 5 flop per test)", "ms": ms,
                    "tests_per_s": K * m / (ms * 1e-3), "fp64_gflops": 5.0 * K * m / (ms * 1e-3) / 1e9})
    img = torch.randint(0, 256, (8, 1024, 2048, 3), dtype=torch.uint8, device="cuda")
    ms = timed(lambda: pcl.resize_cubic(img, (512, 256)), 20)
    by = img.numel() + 8 * 256 * 512 * 3
    out.append({"config": "8f-1: cv2.INTER_CUBIC 8 x 1024x2048x3 -> 256x512x3", "ms": ms, "frames_per_s": 8 / (ms * 1e-3),
                "gb_per_s": by / (ms * 1e-3) / 1e9})
    # 8f-3: ASCII PLY rows of a 1 M-point cloud ('%f %f %f %d %d %d'), device formatting only (no file write)
    import ctypes as C
    from semantic_depth_b200 import _lib
    npts = 1_000_000
    cloud = torch.from_numpy(scene.make_road_cloud(npts, seed=2)).cuda()
    px, py, pz = (cloud[:, i].contiguous() for i in range(3))
    rgb = torch.randint(0, 256, (npts, 3), dtype=torch.uint8, device="cuda")
    engp = engine_for(npts)
    buf = torch.empty(48 * npts, dtype=torch.uint8, device="cuda")
    nb = C.c_ulonglong(0)
    lib = _lib.load()
    def fmt():
        lib.sd_ply_rows(px.data_ptr(), py.data_ptr(), pz.data_ptr(), rgb.data_ptr(), npts, buf.data_ptr(), buf.numel(), C.byref(nb),
                        engp._ws, torch.cuda.current_stream().cuda_stream)
    ms = timed(fmt, 10)
    out.append({"config": "8f-3: ASCII PLY rows of 1 M points (np.savetxt '%f %f %f %d %d %d'), device side incl. one host sync",
                "ms": ms, "rows_per_s": npts / (ms * 1e-3), "output_mb": nb.value / 1e6, "gb_per_s_out": nb.value / (ms * 1e-3) / 1e9})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
