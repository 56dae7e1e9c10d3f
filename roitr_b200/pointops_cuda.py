"""Drop-in for the reference's pybind11 module ``pointops_cuda`` (cpp_wrappers/pointops/src/pointops_api.cpp:12-23).

``sys.modules['pointops_cuda'] = roitr_b200.pointops_cuda`` lets the unmodified reference
``cpp_wrappers/pointops/functions/pointops.py`` (import at :7, calls at :23,:42) run on libroitr_b200: same names,
same argument order, caller-allocated outputs written in place, no return value, launches on the current stream.
The eight functions RoITr's forward never calls (grouping/interpolation/subtraction/aggregation fwd+bwd) are registered
but raise: they are outside the hot path.
"""
import torch

from . import _lib
from ._lib import c_int, f32, i32, stream_ptr


def knnquery_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2):
    """knnquery_cuda(int m, int nsample, xyz, new_xyz, offset, new_offset, idx, dist2) (knnquery_cuda_kernel.h:7).
    Writes idx (m,nsample) int32 and SQUARED distances dist2 (m,nsample) f32."""
    _lib.call("roitr_knnquery_n", c_int(offset.shape[0]), c_int(int(m)), c_int(int(nsample)), c_int(xyz.shape[0]),
              f32(xyz), f32(new_xyz), i32(offset), i32(new_offset), i32(idx), f32(dist2), stream_ptr())


def furthestsampling_cuda(b, n, xyz, offset, new_offset, tmp, idx):
    """furthestsampling_cuda(int b, int n, xyz, offset, new_offset, tmp, idx) (sampling_cuda_kernel.h:7).
    ``n`` (max segment length; a 0-d tensor in the reference, pointops.py:18-23) fixes the tie order; ``tmp`` is unused."""
    n = int(n)
    _lib.call("roitr_furthestsampling_cfg", c_int(int(b)), c_int(n), c_int(n), f32(xyz), i32(offset), i32(new_offset),
              i32(idx), None, c_int(0), stream_ptr())


def _out_of_scope(name):
    def f(*a, **k):
        raise NotImplementedError("pointops_cuda.%s is not on RoITr's forward path and is not provided by roitr_b200 "
                                  "(SURVEY.md §2.1)" % name)
    f.__name__ = name
    return f


for _n in ("grouping_forward_cuda", "grouping_backward_cuda", "interpolation_forward_cuda",
           "interpolation_backward_cuda", "subtraction_forward_cuda", "subtraction_backward_cuda",
           "aggregation_forward_cuda", "aggregation_backward_cuda"):
    globals()[_n] = _out_of_scope(_n)
