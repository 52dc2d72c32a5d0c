"""CPU oracle of the SemanticDepth fusion hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy restatement of the reference's per-frame fusion stage
(``/root/reference/semantic_depth.py:183-324`` plus ``/root/reference/semantic_depth_lib/pcl.py``),
each function citing the reference lines it follows.  It exists so that the CUDA path can be
checked on a box where ``/root/reference`` does not exist.

Rules:

* Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
  ``--impl reference`` legs may import anything from here -- and only as the checker or as the
  timed CPU baseline, never as (part of) the product path.  Nothing under ``semantic_depth_b200/``
  or ``semantic_depth_lib/`` imports this package; the product fails loudly without its CUDA
  library.
* Pinning status (see DESIGN.md "Oracle"):
    - ``pcl_ref`` (filters, MAD, plane fit, split, slab end points, intersection) is pinned
      against the reference's own ``semantic_depth_lib/pcl.py`` imported unmodified in the build
      container: ``tests/golden/make_golden.py`` runs both on the same inputs, asserts bit-equality
      and commits the vectors under ``tests/golden/``.
    - ``frame_ref.post_process_disparity`` / ``reproject_to_3d`` are pinned against
      ``DepthFrame.post_processing`` / ``DepthFrame.compute_3D_points`` lifted with ``ast`` from
      ``semantic_depth.py:656-664,686-697`` (the latter calling the real
      ``cv2.reprojectImageTo3D``), same script, bit-equal.
    - ``labels_from_logits`` restates two lines (``semantic_depth.py:555-556,563-564``); the
      softmax itself ran inside a TF1 session in the reference and is evaluated here in fp64.
    - ``statistical_outlier_removal`` / ``radius_outlier_removal`` restate Open3D <= 0.7
      (``RemoveStatisticalOutliers`` / ``RemoveRadiusOutliers``), a third-party dependency that is
      neither vendored nor pinned by the reference (absent from requirements.txt; call sites
      ``semantic_depth.py:227-245``) and not installable here: **parity unpinned** for these two
      functions; the neighbour search is ``scipy.spatial.cKDTree``.
    - ``ransac_*`` has no reference counterpart (north_star row 8-R); it is the specification.
"""
