"""Host-side mirror of the reference op API ``cpp_wrappers.pointops.functions.pointops`` for the forward path
(same names, argument order and return conventions), backed by libroitr_b200 (sm_100a CUDA):

    furthestsampling(xyz, offset, new_offset) -> idx int32 (m,)                 pointops.py:10-27
    knnquery(nsample, xyz, new_xyz, offset, new_offset) -> (idx int32, dist)    pointops.py:30-45
    queryandgroup(nsample, xyz, new_xyz, feat, idx, offset, new_offset, return_idx=False, use_xyz=True)  :79-104
    interpolation(xyz, new_xyz, feat, offset, new_offset, k=3) -> (n, c)        pointops.py:168-182

plus the fused op the new backbone uses instead of queryandgroup + gathers + calc_ppf_gpu:

    knn_ppf(nsample, xyz, normals, new_xyz, new_normals, offset, new_offset) -> (idx int32 (m,k), ppf (m,k,4))

grouping / subtraction / aggregation / Interpolation(autograd) are not on RoITr's forward path and are
not provided. Everything here is inference-only (no autograd), like the reference's use of these ops.
"""
import torch

from . import _lib
from ._lib import c_int, f32, i32, stream_ptr


def _check_xyz(t):
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.dim() == 2 and t.shape[1] == 3


def furthestsampling(xyz, offset, new_offset, n_max=None, m_total=None, cluster=0, return_xyz=False):
    """xyz (n,3), offset (b,), new_offset (b,) int32 cumulative ends -> idx int32 (m,).

    n_max / m_total: maximum segment length and total sample count if the caller already knows them on the host
    (avoids the reference's offset[..].item() device reads, pointops.py:18-21). The tie order follows the reference
    block size derived from n_max over the batch (src/cuda_utils.h:11-14)."""
    _check_xyz(xyz)
    b = offset.shape[0]
    if n_max is None:
        ends = offset.tolist()
        n_max = max(e - s for s, e in zip([0] + ends[:-1], ends))
    m = int(new_offset[-1].item()) if m_total is None else int(m_total)
    idx = torch.empty(m, dtype=torch.int32, device=xyz.device)
    new_xyz = torch.empty(m, 3, dtype=torch.float32, device=xyz.device) if return_xyz else None
    _lib.call("roitr_furthestsampling_cfg", c_int(b), c_int(n_max), c_int(n_max), f32(xyz), i32(offset),
              i32(new_offset), i32(idx), f32(new_xyz), c_int(cluster), stream_ptr())
    return (idx, new_xyz) if return_xyz else idx


def knnquery(nsample, xyz, new_xyz, offset, new_offset):
    """-> (idx int32 (m,nsample), dist f32 (m,nsample)) with dist = sqrt(d2), ascending (pointops.py:43)."""
    if new_xyz is None:
        new_xyz = xyz
    _check_xyz(xyz)
    _check_xyz(new_xyz)
    m = new_xyz.shape[0]
    idx = torch.empty(m, nsample, dtype=torch.int32, device=xyz.device)
    dist = torch.empty(m, nsample, dtype=torch.float32, device=xyz.device)
    _lib.call("roitr_knn_ppf_n", c_int(offset.shape[0]), c_int(m), c_int(nsample), c_int(0), c_int(xyz.shape[0]),
              f32(xyz), None, f32(new_xyz), None, i32(offset), i32(new_offset), i32(idx), f32(dist), None, stream_ptr())
    return idx, dist


def knn_ppf(nsample, xyz, normals, new_xyz, new_normals, offset, new_offset, drop_first=1):
    """Fused queryandgroup(return_idx=True) + calc_ppf_gpu: idx int32 (m,nsample), ppf f32 (m,nsample,4)."""
    _check_xyz(xyz)
    _check_xyz(new_xyz)
    _check_xyz(normals)
    _check_xyz(new_normals)
    m = new_xyz.shape[0]
    idx = torch.empty(m, nsample, dtype=torch.int32, device=xyz.device)
    ppf = torch.empty(m, nsample, 4, dtype=torch.float32, device=xyz.device)
    _lib.call("roitr_knn_ppf_n", c_int(offset.shape[0]), c_int(m), c_int(nsample), c_int(drop_first),
              c_int(xyz.shape[0]), f32(xyz), f32(normals), f32(new_xyz), f32(new_normals), i32(offset), i32(new_offset),
              i32(idx), None, f32(ppf), stream_ptr())
    return idx, ppf


def queryandgroup(nsample, xyz, new_xyz, feat, idx, offset, new_offset, return_idx=False, use_xyz=True):
    """pointops.py:79-104. With return_idx=True (the only way RoITr calls it): kNN(nsample+1), drop column 0, int64."""
    if new_xyz is None:
        new_xyz = xyz
    if idx is None:
        m = new_xyz.shape[0]
        idx = torch.empty(m, nsample, dtype=torch.int32, device=xyz.device)
        _lib.call("roitr_knn_ppf_n", c_int(offset.shape[0]), c_int(m), c_int(nsample), c_int(1), c_int(xyz.shape[0]),
                  f32(xyz), None, f32(new_xyz), None, i32(offset), i32(new_offset), i32(idx), None, None, stream_ptr())
        idx = idx.long()
    if return_idx:
        return idx
    m, c = new_xyz.shape[0], feat.shape[1]
    grouped_xyz = gather_rows(xyz, idx.reshape(-1)).view(m, nsample, 3) - new_xyz.unsqueeze(1)
    grouped_feat = gather_rows(feat, idx.reshape(-1)).view(m, nsample, c)
    return torch.cat((grouped_xyz, grouped_feat), -1) if use_xyz else grouped_feat


def interpolation(xyz, new_xyz, feat, offset, new_offset, k=3, base=None):
    """xyz (m,3) coarse, new_xyz (n,3) fine, feat (m,c) -> (n,c); optional ``base`` (n,c) is added (model/model.py:116)."""
    assert feat.is_contiguous()
    idx, dist = knnquery(k, xyz, new_xyz, offset, new_offset)
    n, c = new_xyz.shape[0], feat.shape[1]
    out = torch.empty(n, c, dtype=torch.float32, device=feat.device)
    _lib.call("roitr_interpolate", c_int(n), c_int(c), c_int(k), i32(idx), f32(dist), f32(feat), f32(base), f32(out),
              stream_ptr())
    return out


def gather_rows(src, index, pad_row=-1):
    """src (n,c) f32, index int32/int64 (any shape) -> index.shape + (c,). Rows equal to pad_row read as zeros."""
    import ctypes
    assert src.is_contiguous() and src.dtype == torch.float32 and index.is_contiguous()
    c = src.shape[1]
    rows = index.numel()
    out = torch.empty(*index.shape, c, dtype=torch.float32, device=src.device)
    _lib.call("roitr_gather_rows", ctypes.c_longlong(rows), c_int(c), _lib.ptr(index),
              c_int(1 if index.dtype == torch.int64 else 0), f32(src), f32(out), ctypes.c_longlong(pad_row), stream_ptr())
    return out
