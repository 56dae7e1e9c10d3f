"""The step after the hot path: rigid transform from correspondences on the GPU.

``weighted_procrustes`` keeps the reference's name, argument meaning and return convention (lib/utils.py:159-218); the
computation is one kernel (csrc/procrustes.cu). ``ransac_pose_estimation_correspondences`` keeps the name, arguments and
return value of the estimator the reference's evaluator calls (registration/benchmark_utils.py:165-209) and adds a batched
form (csrc/ransac.cu: every pair's 50 000 hypotheses in one launch). No CPU fallback."""
import ctypes

import torch

from . import _lib
from ._lib import c_float, c_int, f32, i32, ptr, stream_ptr

RANSAC_ITERATIONS = 50000      # RANSACConvergenceCriteria(50000, 1000) (benchmark_utils.py:198-199)
RANSAC_EDGE_SIMILARITY = 0.9   # CorrespondenceCheckerBasedOnEdgeLength(0.9) (:194-195)


def ransac_batch(src_corr_list, tgt_corr_list, distance_threshold=0.05, iterations=RANSAC_ITERATIONS, seed=0):
    """Batched correspondence RANSAC. src_corr_list / tgt_corr_list: per pair the matched points (n_i, 3) f32 CUDA (row j of
    both = correspondence j, which is how evaluate_registration_c2f.py:85-88 calls the reference: correspondences =
    arange(n) twice). Returns (transforms (B,4,4) f64, fitness (B,) f64, inlier_rmse (B,) f64, best_iteration (B,) int32),
    all on the device; no host sync."""
    B = len(src_corr_list)
    if B == 0 or not src_corr_list[0].is_cuda:
        raise _lib.RoitrError("ransac_batch: expected a non-empty list of CUDA tensors (there is no CPU path)")
    dev = src_corr_list[0].device
    sizes = [int(s.shape[0]) for s in src_corr_list]
    src = torch.cat([s.reshape(-1, 3).float() for s in src_corr_list]).contiguous()
    tgt = torch.cat([t.reshape(-1, 3).float() for t in tgt_corr_list]).contiguous()
    if src.shape != tgt.shape:
        raise _lib.RoitrError("ransac_batch: src / tgt correspondence lists differ in size")
    ends, acc = [], 0
    for n in sizes:
        acc += n
        ends.append(acc)
    offset = torch.tensor(ends, dtype=torch.int32, device=dev)
    fn = _lib.lib().roitr_ransac_workspace_bytes
    fn.restype = ctypes.c_longlong
    ws = torch.empty(int(fn(c_int(B), c_int(iterations))), dtype=torch.uint8, device=dev)
    T = torch.empty(B, 4, 4, dtype=torch.float64, device=dev)
    fit = torch.empty(B, dtype=torch.float64, device=dev)
    rmse = torch.empty(B, dtype=torch.float64, device=dev)
    itr = torch.empty(B, dtype=torch.int32, device=dev)
    _lib.call("roitr_ransac_correspondences", c_int(B), c_int(iterations), f32(src if src.numel() else None),
              f32(tgt if tgt.numel() else None), i32(offset), c_int(max(sizes)), ctypes.c_double(distance_threshold),
              ctypes.c_double(RANSAC_EDGE_SIMILARITY), ctypes.c_uint(seed & 0xFFFFFFFF), ptr(ws), ptr(T, torch.float64),
              ptr(fit, torch.float64), ptr(rmse, torch.float64), i32(itr), stream_ptr())
    return T, fit, rmse, itr


def ransac_pose_estimation_correspondences(src_pcd, tgt_pcd, correspondences, mutual=False, distance_threshold=0.05,
                                           ransac_n=3, seed=0):
    """registration/benchmark_utils.py:165-209: src_pcd (n,3), tgt_pcd (m,3), correspondences (c,2) integer index pairs ->
    (4,4) float64 numpy transform (``result_ransac.transformation``)."""
    if mutual:
        raise NotImplementedError
    if ransac_n != 3:
        raise _lib.RoitrError("ransac_n must be 3 (what the reference passes)")
    dev = src_pcd.device if torch.is_tensor(src_pcd) and src_pcd.is_cuda else torch.device("cuda", torch.cuda.current_device())
    src = torch.as_tensor(src_pcd, dtype=torch.float32).to(dev)
    tgt = torch.as_tensor(tgt_pcd, dtype=torch.float32).to(dev)
    corr = torch.as_tensor(correspondences).to(dev).long()
    if corr.shape[0] < 3:
        return torch.eye(4, dtype=torch.float64).numpy()        # RegistrationResult(): Open3D returns the identity
    T, _, _, _ = ransac_batch([src[corr[:, 0]]], [tgt[corr[:, 1]]], distance_threshold, seed=seed)
    return T[0].cpu().numpy()


def sample_correspondences(confidence, n_points, generator=None):
    """The evaluator's subsampling before RANSAC (evaluate_registration_c2f.py:78-84): when there are more than n_points
    correspondences, draw n_points of them without replacement with probability proportional to the confidence. Returns
    the selected indices (all of them, in order, when there are at most n_points)."""
    n = confidence.shape[0]
    if n <= n_points:
        return torch.arange(n, device=confidence.device)
    return torch.multinomial(confidence.float() / confidence.float().sum(), n_points, replacement=False, generator=generator)


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
