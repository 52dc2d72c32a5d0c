"""Golden answers of BASELINE.json configs[3] (2 M-point road cloud, 2 % box outliers, k = 16, ratio 0.5) from the CPU
oracle (cKDTree neighbour sets, Open3D's arithmetic: oracle/frame_ref.py::keep_statistical_outlier_removal).

``python tests/golden/make_golden_config4.py`` -> tests/golden/config4_sor.npz (about 2 minutes of CPU).  The full fp64
array of per-point mean distances is 16 MB, so the fixture keeps its SHA-256, the statistics, the kept count, a checksum
of the kept indices, 4 096 sampled values and the distance of the closest point to the threshold (the tie margin)."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import frame_ref  # noqa: E402
from semantic_depth_b200 import scene  # noqa: E402

N, K, RATIO, SEED = 2_000_000, 16, 0.5, 0


def index_checksum(idx):
    idx = np.asarray(idx)
    return np.uint64(int(np.sum(idx.astype(np.uint64) * (np.arange(idx.size, dtype=np.uint64) % 65521 + 1)) % (1 << 63)))


def main():
    pts = scene.make_road_cloud(N, seed=SEED)
    keep, avg, (thr, mu, sd) = frame_ref.keep_statistical_outlier_removal(pts, K, RATIO, workers=-1)
    rel = np.abs(avg - thr) / thr
    sample = np.random.default_rng(1).choice(N, 4096, replace=False)
    np.savez_compressed(os.path.join(HERE, "config4_sor.npz"),
                        n=N, k=K, ratio=RATIO, seed=SEED, thr=np.float64(thr), mean=np.float64(mu), std=np.float64(sd),
                        kept=np.int64(keep.size), kept_checksum=index_checksum(keep),
                        avg_sha256=np.frombuffer(hashlib.sha256(np.ascontiguousarray(avg).tobytes()).digest(), dtype=np.uint8),
                        sample_idx=sample.astype(np.int64), sample_avg=avg[sample], tie_margin_rel=np.float64(rel.min()))
    print(f"config 4: kept {keep.size} of {N}, thr {thr!r}, mean {mu!r}, std {sd!r}, closest point to the threshold: rel {rel.min():.3e}")


if __name__ == "__main__":
    main()
