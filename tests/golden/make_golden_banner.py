"""Golden vectors of the result banner, produced by the REFERENCE'S OWN STATEMENTS with the real cv2.
BUILD CONTAINER ONLY (needs opencv and /root/reference): ``python tests/golden/make_golden_banner.py [--verify]``.

The banner code sits in the middle of ``process_frame`` (semantic_depth.py:346-394; live twin
semantic_depth_cityscapes_sequence.py:306-327).  Neither driver can be imported (TensorFlow 1.x, Open3D, scipy.misc), so the
statements are lifted with ``ast`` -- the ``if self.is_city: ... else: ...`` preset block and every ``cv2.rectangle`` /
``cv2.putText`` call of the "Draw letters" section, in source order -- and executed against a stub ``self`` and the local
variables they read.  The script asserts that oracle/banner_ref.py gives the same bytes and stores the banner rows of each
case (the rows below are untouched: asserted)."""
import ast
import os
import sys
import types

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import banner_ref  # noqa: E402
from tests_banner_cases import CASES, VALUES, base_frame  # noqa: E402

REF = "/root/reference"


def _calls_cv2_draw(node) -> bool:
    for n in ast.walk(node):
        if isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute) and n.func.attr in ("rectangle", "putText"):
            return True
    return False


def _has_call(node, names) -> bool:
    for n in ast.walk(node):
        if isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute) and n.func.attr in names:
            return True
    return False


def lift_banner_statements(path: str, single: bool):
    """The statements of the 'Draw letters' section of process_frame, minus resize / imwrite."""
    tree = ast.parse(open(path).read())
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "process_frame")
    if single:
        # semantic_depth.py: everything lives under `if self.save_data:`
        blk = next(n for n in ast.walk(fn) if isinstance(n, ast.If) and isinstance(n.test, ast.Attribute)
                   and n.test.attr == "save_data" and _calls_cv2_draw(n))
        body = blk.body
    else:
        body = fn.body
    start = next(i for i, st in enumerate(body) if _has_call(st, ("resize",)) and not _calls_cv2_draw(st)
                 and any(_calls_cv2_draw(t) for t in body[i + 1:i + 8]))
    last = max(i for i, st in enumerate(body) if _calls_cv2_draw(st))
    keep = []
    for st in body[start + 1:last + 1]:         # from the cubic up-sampling (exclusive) to the last drawing statement
        if _has_call(st, ("imwrite", "print_line_on_image")):
            continue
        plain_assign = isinstance(st, ast.Assign) and all(isinstance(t, ast.Name) for t in st.targets)
        preset_if = isinstance(st, ast.If) and all(isinstance(x, ast.Assign) for x in st.body)
        assert _calls_cv2_draw(st) or plain_assign or preset_if, ast.dump(st)[:200]
        keep.append(st)
    mod = ast.Module(body=keep, type_ignores=[])
    return compile(ast.fix_missing_locations(mod), path, "exec"), len(keep)


def run_reference(name, driver, h, w, kw, seed):
    img = base_frame(h, w, seed)
    path = os.path.join(REF, "semantic_depth.py" if driver == "single" else "semantic_depth_cityscapes_sequence.py")
    code, n = lift_banner_statements(path, driver == "single")
    me = types.SimpleNamespace(segmented_frame=img, depth=kw["depth"], is_city=kw.get("is_city", True), approach=kw.get("approach", "both"),
                               save_data=True, verbose=False)
    ns = dict(self=me, cv2=cv2, original_height=h, original_width=w, h=h, w=w, line_found=kw.get("line_found", True), **VALUES)
    exec(code, ns)
    return me.segmented_frame, n


def product_spec(driver, h, w, kw):
    from semantic_depth_b200 import banner
    if driver == "single":
        return banner.result_banner_spec(h, w, kw["depth"], VALUES["left_pt_rw"], VALUES["right_pt_rw"], VALUES["dist_rw"],
                                         VALUES["left_pt_f2f"], VALUES["right_pt_f2f"], VALUES["dist_f2f"],
                                         is_city=kw["is_city"], approach=kw["approach"])
    return banner.sequence_banner_spec(h, w, kw["depth"], kw["line_found"], VALUES["left_pt_rw"], VALUES["right_pt_rw"], VALUES["dist_rw"])


def main():
    cv2.setNumThreads(1)
    out = {}
    for i, (name, driver, h, w, kw) in enumerate(CASES):
        ref, nst = run_reference(name, driver, h, w, kw, i)
        rects, texts = product_spec(driver, h, w, kw)
        mine = banner_ref.draw(base_frame(h, w, i), [(p1, p2, c) for _, p1, p2, c in rects],
                               [(t, org, s, th, c) for _, t, org, s, th, c in texts])
        rows = int(0.27 * h)
        assert np.array_equal(ref[rows:], base_frame(h, w, i)[rows:]), "the banner reaches below the stored rows"
        diff = int((ref != mine).any(axis=2).sum())
        print(f"case {i} {name}: {nst} reference statements executed, {len(texts)} text lines, oracle vs cv2: {diff} differing pixels")
        assert diff == 0, name
        out[f"case{i}_shape"] = np.array([h, w, rows, i])
        out[f"case{i}_top"] = ref[:rows]
    out["names"] = np.array([c[0] for c in CASES])
    path = os.path.join(HERE, "banner_vectors.npz")
    if "--verify" in sys.argv:
        old = np.load(path)
        for k, v in out.items():
            assert np.array_equal(old[k], v), k
        print("committed fixtures == live reference")
        return
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
