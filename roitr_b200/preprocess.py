"""The step before the hot path: per-point surface normals on the GPU.

Mirrors what the reference's datasets do on the CPU with Open3D for every pair (dataset/tdmatch.py:120-127,
dataset/fdmatch.py:83-90): ``estimate_normals(KDTreeSearchParamKNN(knn=33))`` followed by ``normal_redirect`` towards the
view point (dataset/common.py:312-320). One kernel per call: grid kNN (one thread per point, the same search as the kNN
op), covariance from cumulants and the smallest-eigenvalue eigenvector in fp64. No CPU fallback."""
import ctypes

import torch

from . import _lib, ops
from ._lib import c_int, f32, i32, ptr, stream_ptr


def estimate_normals(points, offset=None, knn=33, view_point=(0.0, 0.0, 0.0)):
    """points (n,3) f32 CUDA, offset (b,) int32 cumulative segment ends (default: one cloud) -> normals (n,3) f32, unit
    length, oriented towards ``view_point``."""
    if not points.is_cuda:
        raise _lib.RoitrError("estimate_normals: expected a CUDA tensor (there is no CPU path)")
    points = points.contiguous().float()
    n = points.shape[0]
    if offset is None:
        offset = torch.tensor([n], dtype=torch.int32, device=points.device)
    grid = ops.knn_grid_build(points, offset)
    out = torch.empty(n, 3, dtype=torch.float32, device=points.device)
    vp = (ctypes.c_float * 3)(*[float(v) for v in view_point])
    _lib.call("roitr_estimate_normals", c_int(offset.shape[0]), c_int(n), c_int(knn), f32(points), i32(offset), ptr(grid), vp,
              f32(out), stream_ptr())
    return out
