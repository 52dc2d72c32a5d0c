"""Pin oracle/ply_ref.py against the reference's own PointCloud2Ply (imported unmodified).  BUILD CONTAINER ONLY.

``python tests/golden/make_golden_ply.py`` writes tests/golden/ply_vectors.npz: seeds + sha256 / length / first rows of the
file the reference class writes, after asserting byte-equality with the oracle."""
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ply_ref  # noqa: E402


def make_cloud(seed, n, dtype, genuine64=False):
    """genuine64: float64 values that are NOT float32-representable (the reference's plane meshes and lines)."""
    rng = np.random.default_rng(seed)
    if genuine64:
        p = rng.standard_normal((n, 3)) * np.array([4.0, 1.0, 30.0]) + np.array([0.0, -1.5, -40.0])
        p[::89, 2] = p[:, 2].min()
        special = np.array([0.0, -0.0, 2.5e-7, -4.9999999e-7, 5.0000001e-7, 0.0078125, 0.0234375, 1.5, 1.5e-6, 1e15 + 0.3,
                            123456789.123456789, -1e-7, 5e-324, 1e22, 2.0 ** 100, 1.5 * 2.0 ** 127, -9.9999995, 0.9999995,
                            0.99999949999999, 1e-300, 4503599627370497.5, 0.1 + 0.2])
        p[1:1 + len(special), 0] = special
        p[1:1 + len(special), 1] = -special
        c = rng.integers(0, 256, (n, 3)).astype(np.uint8)
        return p, c
    p = (rng.standard_normal((n, 3)) * np.array([4.0, 1.0, 30.0]) + np.array([0.0, -1.5, -40.0])).astype(np.float32)
    p[::97, 2] = p[:, 2].min()                      # several rows at the minimum z: the infinity filter drops them all
    special = np.float32([0.0, -0.0, 1e-7, -4.9999999e-7, 5.0000001e-7, 0.5, 2.5e-6, 123456.789, -1e6, 1.0000005, 9.9999995, 3.4e38, 1e-45])
    p[1:1 + len(special), 0] = special
    c = rng.integers(0, 256, (n, 3)).astype(np.uint8)
    return p.astype(dtype), c


def main():
    sys.path.insert(0, "/root/reference")                                    # only here: tests import make_cloud from this module
    from semantic_depth_lib.point_cloud_2_ply import PointCloud2Ply          # the reference's class, unmodified
    out = {}
    for i, (n, dtype, genuine64) in enumerate([(1000, np.float32, False), (50_000, np.float32, False), (3000, np.float64, False),
                                               (20, np.float32, False), (4000, np.float64, True)]):
        p, c = make_cloud(i, n, dtype, genuine64)
        with tempfile.TemporaryDirectory() as d:
            w = PointCloud2Ply(p.copy(), c.copy(), os.path.join(d, "cloud"))
            w.prepare_and_save_point_cloud()
            ref = open(os.path.join(d, "cloud.ply"), "rb").read()
        mine = ply_ref.prepare_and_save_bytes(p, c)
        assert mine == ref, (i, len(mine), len(ref))
        out[f"case{i}"] = np.array([i, n, (2 if genuine64 else 1) if dtype == np.float64 else 0, len(ref)])
        out[f"case{i}_sha256"] = np.frombuffer(hashlib.sha256(ref).digest(), dtype=np.uint8)
        out[f"case{i}_head"] = np.frombuffer(ref[:400], dtype=np.uint8)
        print(f"case {i}: {n} points {dtype.__name__} -> {len(ref)} bytes, oracle == reference")
    path = os.path.join(HERE, "ply_vectors.npz")
    if len(sys.argv) > 1 and sys.argv[1] == "--verify":
        old = np.load(path)
        assert set(old.files) == set(out) and all(np.array_equal(old[k], out[k]) for k in out)
        print("committed fixtures == live reference")
    else:
        np.savez_compressed(path, **out)


if __name__ == "__main__":
    main()
