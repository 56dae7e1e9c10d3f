"""The step after the hot path: rigid transform from weighted correspondences on the GPU.

``weighted_procrustes`` keeps the reference's name, argument meaning and return convention (lib/utils.py:159-218); the
computation is one kernel (csrc/procrustes.cu). No CPU fallback."""
import torch

from . import _lib
from ._lib import c_float, c_int, f32, stream_ptr


def weighted_procrustes(src_points, tgt_points, weights=None, weight_thresh=0.0, eps=1e-5, return_transform=False):
    """src_points, tgt_points (B, N, 3) or (N, 3); weights (B, N) or (N,) -> R (B,3,3), t (B,3) [or a (B,4,4) transform]."""
    if not src_points.is_cuda:
        raise _lib.RoitrError("weighted_procrustes: expected CUDA tensors (there is no CPU path)")
    squeeze_first = src_points.ndim == 2
    if squeeze_first:
        src_points, tgt_points = src_points.unsqueeze(0), tgt_points.unsqueeze(0)
        weights = weights.unsqueeze(0) if weights is not None else None
    B, N = src_points.shape[0], src_points.shape[1]
    src, tgt = src_points.contiguous().float(), tgt_points.contiguous().float()
    w = weights.contiguous().float() if weights is not None else None
    R = torch.empty(B, 3, 3, dtype=torch.float32, device=src.device)
    t = torch.empty(B, 3, dtype=torch.float32, device=src.device)
    _lib.call("roitr_weighted_procrustes", c_int(B), c_int(N), f32(src), f32(tgt), f32(w), c_float(weight_thresh), c_float(eps),
              f32(R), f32(t), stream_ptr())
    if return_transform:
        T = torch.eye(4, dtype=torch.float32, device=src.device).unsqueeze(0).repeat(B, 1, 1)
        T[:, :3, :3] = R
        T[:, :3, 3] = t
        return T.squeeze(0) if squeeze_first else T
    return (R.squeeze(0), t.squeeze(0)) if squeeze_first else (R, t)
