"""``semantic_depth_lib.point_cloud_2_ply`` -- the reference's ASCII PLY writer, formatted on the GPU.

Same class, constructor and methods as /root/reference/semantic_depth_lib/point_cloud_2_ply.py:33-93; the file written
is byte-identical (header with the reference's indentation, ``'%f %f %f %d %d %d'`` rows).  The "infinity" filter
``z > z.min()`` (:87-89) is a stable compaction, the rows are formatted by ``sd_ply_rows`` (exact integer arithmetic for
``%f``), the host only writes the bytes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from semantic_depth_b200 import _lib
from semantic_depth_b200._lib import PRED_GT, SdPredicate, check
from semantic_depth_b200.pcl_gpu import engine_for


class PointCloud2Ply():

    """3D point cloud tools."""

    #: Header for exporting point cloud to PLY (the reference's literal, point_cloud_2_ply.py:38-49)
    ply_header = (
    '''ply
    format ascii 1.0
    element vertex {vertex_count}
    property float x
    property float y
    property float z
    property uchar red
    property uchar green
    property uchar blue
    end_header
    ''')

    def __init__(self, points3D, colors, output_name):
        self.points3D = np.asarray(points3D).reshape(-1, 3)
        self.colors = np.asarray(colors).reshape(-1, 3)
        self.output_name = output_name

    # -- device helpers ------------------------------------------------------------------------------------------
    @staticmethod
    def _device_cloud(points3D, colors):
        p = np.ascontiguousarray(points3D)
        if p.dtype == np.float64:
            with np.errstate(over="ignore"):
                p32 = p.astype(np.float32)
            if np.array_equal(p32.astype(np.float64), p, equal_nan=True):
                p = p32                                # e.g. a cloud that only made the Open3D round trip (semantic_depth.py:244)
            else:                                      # plane meshes, lines (pcl.py:107-113,321-331): formatted as float64
                fin = np.abs(p[np.isfinite(p)])
                if fin.size and fin.max() >= 2.0 ** 128:
                    raise NotImplementedError("finite float64 coordinates must be below 2^128")
        elif p.dtype != np.float32:
            raise TypeError("points3D must be float32 or float64")
        c = np.asarray(colors)
        if c.dtype != np.uint8:
            ci = c.astype(np.int64)                   # '%d' truncates towards zero
            if ci.size and (ci.min() < 0 or ci.max() > 255):
                raise NotImplementedError("colors outside 0..255")
            c = ci.astype(np.uint8)
        soa = torch.from_numpy(np.ascontiguousarray(p.T)).cuda()
        return soa[0].contiguous(), soa[1].contiguous(), soa[2].contiguous(), torch.from_numpy(np.ascontiguousarray(c)).cuda()

    def _rows(self, points3D, colors) -> bytes:
        n = points3D.shape[0]
        if n == 0:
            return b""
        x, y, z, rgb = self._device_cloud(points3D, colors)
        eng = engine_for(n, x.device)
        lib = _lib.load()
        capacity = 48 * n + 1024
        for _ in range(2):
            out = torch.empty(capacity, dtype=torch.uint8, device=x.device)
            nbytes = C.c_ulonglong(0)
            fn = lib.sd_ply_rows if x.dtype == torch.float32 else lib.sd_ply_rows_f64
            with torch.cuda.device(x.device):
                rc = fn(x.data_ptr(), y.data_ptr(), z.data_ptr(), rgb.data_ptr(), n, out.data_ptr(), capacity,
                        C.byref(nbytes), eng._ws, torch.cuda.current_stream().cuda_stream)
            if rc == 0:
                return out[: nbytes.value].cpu().numpy().tobytes()
            if nbytes.value <= capacity:
                check(rc, "sd_ply_rows")
            capacity = int(nbytes.value)
        check(rc, "sd_ply_rows")

    def write_ply(self, output_file):
        """Export ``PointCloud`` to PLY file for viewing in MeshLab."""
        rows = self._rows(self.points3D, self.colors)
        with open(output_file, 'wb') as f:
            f.write(self.ply_header.format(vertex_count=len(self.points3D)).encode("ascii"))
            f.write(rows)
        print("Point Cloud file generated!")

    def add_extra_point_cloud(self, points3D_extra, colors_extra):
        self.points3D = np.append(self.points3D, points3D_extra, axis=0)
        self.colors = np.append(self.colors, colors_extra, axis=0)

    def prepare_and_save_point_cloud(self):
        """Apply the infinity filter (``z > z.min()``, :87-89) and save the points into ``<output_name>.ply``."""
        n = self.points3D.shape[0]
        if n == 0:
            raise ValueError("zero-size array to reduction operation minimum which has no identity")   # np.min of the reference
        x, y, z, _ = self._device_cloud(self.points3D, self.colors)
        if z.dtype == torch.float64:
            # genuinely float64 rows (visualisation meshes / lines appended to the cloud): the filter kernels work on
            # float32 columns, so this one predicate is evaluated with torch (same semantics: NaN minimum keeps nothing)
            idx = torch.nonzero(z > z.min()).reshape(-1).cpu().numpy().astype(np.int64)
        else:
            eng = engine_for(n, z.device)
            zmin, _, _ = eng.slab_minmax(z, z, 0.0, 0.0, use_f32=2)                # min of z over ALL rows (-inf included)
            if np.isnan(self.points3D[:, 2]).any():
                zmin = np.float32(np.nan)                                           # np.min propagates NaN -> nothing is kept
            keep, _ = eng.filter(x, y, z, SdPredicate(kind=PRED_GT, axis=2, fa=float(zmin)), want_points=False)
            idx = keep.cpu().numpy().astype(np.int64)
        self.points3D = self.points3D[idx]
        self.colors = self.colors[idx]
        self.write_ply('{}.ply'.format(self.output_name))
