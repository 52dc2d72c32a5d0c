"""``semantic_depth_lib.pcl`` -- the reference's point-cloud call surface, served by CUDA kernels.

Every public name of /root/reference/semantic_depth_lib/pcl.py:30-331 is exported with the same
positional order, defaults and return arity; see ``semantic_depth_b200.pcl_gpu`` for the semantics
and the reference lines each function follows.  Additive entry points cover what the reference's
``FrameProcessor.process_frame`` delegates to other libraries on the same path:

* ``statistical_outlier_removal`` / ``radius_outlier_removal`` -- Open3D, semantic_depth.py:227-245
* ``labels_from_logits``      -- softmax threshold, semantic_depth.py:550-556,563-564
* ``post_process_disparity``  -- DepthFrame.post_processing + cast, semantic_depth.py:656-664,676
* ``reproject_to_3d``         -- DepthFrame.compute_3D_points (cv2.reprojectImageTo3D), :686-697
* ``fuse_frames``             -- the whole fusion section of process_frame, :183-324, batched
* ``resize_cubic``            -- cv2.resize(..., INTER_CUBIC) of the input frame, semantic_depth.py:110-112
* ``overlay_masks`` / ``segment_frame`` -- the masks and the overlaid frame of SegmentFrame.segment_frame, :547-570
* ``draw_banner`` / ``result_banner`` / ``sequence_banner`` -- cv2.rectangle + cv2.putText of the result banner,
  semantic_depth.py:339-394 and semantic_depth_cityscapes_sequence.py:304-327
* ``upsample_scores`` / ``fuse_frames_from_scores`` -- the same path fed by FCN-8s' unexpanded head
  (``second_skip`` + the last transposed convolution, fcn8s/fcn.py:207-213)
"""
from semantic_depth_b200.pcl_gpu import (  # noqa: F401
    remove_from_to, remove_noise_by_mad, mad, remove_noise_by_fitting_plane,
    planes_intersection_at_certain_depth, threshold_complete, extract_pcls, get_end_points_of_road,
    get_end_points_of_segment, compute_distance_in_3D, create_3Dline_from_3Dpoints,
    statistical_outlier_removal, radius_outlier_removal,
)
from semantic_depth_b200.frame_ops import (  # noqa: F401
    labels_from_logits, post_process_disparity, reproject_to_3d, fuse_frames, upsample_scores, fuse_frames_from_scores, resize_cubic,
    overlay_masks, segment_frame, draw_banner, result_banner, sequence_banner,
)
