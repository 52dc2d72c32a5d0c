"""Drop-in replacement of the reference package ``semantic_depth_lib`` for the fusion hot path.

``import semantic_depth_lib.pcl as pcl`` (as /root/reference/semantic_depth.py:70-71 does) now reaches the
B200 CUDA kernels of ``semantic_depth_b200`` instead of NumPy loops.  There is no CPU fallback.
"""
