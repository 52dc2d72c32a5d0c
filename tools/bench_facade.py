"""Cost of the per-function facade (semantic_depth_lib.pcl called one filter at a time, the way a maintainer who only swaps
`import pcl` would): NumPy arrays in / out (upload + AoS->SoA + download on EVERY call) against CUDA tensors in / out and
against the fused path.  One JSON line per row.  Developer measurement, not a test."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from semantic_depth_b200 import scene
import semantic_depth_lib.pcl as pcl


def wall(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3


def chain(road, colors):
    """The road chain of semantic_depth.py:206-259, one facade call per filter."""
    road, colors = pcl.remove_from_to(road, colors, 2, 0.0, 7.0)
    road, colors = pcl.remove_noise_by_mad(road, colors, 1, 15.0)
    road, colors = pcl.remove_noise_by_mad(road, colors, 0, 2.0)
    road, colors, _, _, _ = pcl.remove_noise_by_fitting_plane(road, colors, axis=1, threshold=5.0, plane_color=[200, 200, 200])
    road, colors = pcl.statistical_outlier_removal(road, colors, 10, 0.5)
    road, colors = pcl.radius_outlier_removal(road, colors, 80, 0.5)
    return pcl.get_end_points_of_road(road, 10.0 - 0.02)


def main():
    n = 600_000
    pts = scene.make_road_cloud(n, seed=3).astype(np.float32)
    cols = np.zeros((n, 3), np.uint8)
    out = []
    ms_np = wall(lambda: chain(pts, cols), 3)
    out.append({"path": "facade, NumPy in / NumPy out (H2D + D2H on every call)", "points": n, "ms_per_frame": ms_np})
    tp, tc = torch.from_numpy(pts).cuda(), torch.from_numpy(cols).cuda()
    ms_t = wall(lambda: chain(tp, tc), 3)
    out.append({"path": "facade, CUDA tensors in / out (no PCIe, one host sync per call)", "points": n, "ms_per_frame": ms_t})
    ms_one = wall(lambda: pcl.remove_noise_by_mad(pts, cols, 1, 15.0), 5)
    out.append({"path": "one remove_noise_by_mad call, NumPy in / out", "points": n, "ms": ms_one,
                "pcie_bytes": n * (12 + 3) * 2, "note": "15 B/point up, survivors down"})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
