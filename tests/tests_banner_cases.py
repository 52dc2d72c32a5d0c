"""Inputs of the banner golden cases, shared by tests/golden/make_golden_banner.py (which runs the reference's own
statements on them with cv2) and the tests (which must not need cv2)."""
import numpy as np


def base_frame(h, w, seed):
    """A smooth deterministic BGR image (it compresses well where the banner leaves it visible)."""
    y, x = np.mgrid[0:h, 0:w]
    return np.stack([(x * 3 + y * 5 + seed * 17) % 256, (x + y * 2 + 40 * seed) % 256, (x * 7 + y + 90) % 256], axis=2).astype(np.uint8)


CASES = [
    # name, driver, h, w, kwargs
    ("city_both", "single", 1024, 2048, dict(is_city=True, approach="both", depth=10.0)),
    ("city_rw", "single", 1024, 2048, dict(is_city=True, approach="rw", depth=12.5)),
    ("munich_both", "single", 3024, 4032, dict(is_city=False, approach="both", depth=10.0)),
    ("sequence_found", "sequence", 1024, 2048, dict(line_found=True, depth=10.0)),
    ("sequence_lost", "sequence", 1024, 2048, dict(line_found=False, depth=10.0)),
]
VALUES = dict(left_pt_rw=np.array([[-3.912345, -1.5, -9.98]]), right_pt_rw=np.array([[3.806789, -1.5, -9.98]]), dist_rw=7.719134,
              left_pt_f2f=np.array([[-4.001234, -1.2, -10.0]]), right_pt_f2f=np.array([[3.494857, -1.2, -10.0]]), dist_f2f=7.496091)
