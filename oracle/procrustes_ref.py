"""TEST INFRASTRUCTURE (oracle) - CPU restatement of weighted_procrustes (lib/utils.py:159-218), plain PyTorch.

Line-by-line the reference's algorithm with the hard-coded ``.cuda()`` calls (lib/utils.py:200,208) dropped; pinned to the
reference function itself (imported through oracle/reference_shim.py) in tests/test_procrustes.py whenever /root/reference
is present, and to the committed fixture tests/golden/procrustes_golden.npz everywhere. Only tests/ may import this file."""
import torch


def weighted_procrustes(src_points, tgt_points, weights=None, weight_thresh=0.0, eps=1e-5):
    """-> R (B,3,3), t (B,3) (or unbatched for 2-D inputs)."""
    squeeze = src_points.ndim == 2
    if squeeze:
        src_points, tgt_points = src_points.unsqueeze(0), tgt_points.unsqueeze(0)
        weights = weights.unsqueeze(0) if weights is not None else None
    B = src_points.shape[0]
    if weights is None:
        weights = torch.ones_like(src_points[:, :, 0])
    weights = torch.where(torch.lt(weights, weight_thresh), torch.zeros_like(weights), weights)      # :188
    weights_norm = weights / (torch.sum(weights, dim=1, keepdim=True) + eps)                         # :189
    src_centroid = torch.sum(src_points * weights_norm.unsqueeze(2), dim=1, keepdim=True)            # :191
    tgt_centroid = torch.sum(tgt_points * weights_norm.unsqueeze(2), dim=1, keepdim=True)
    sc, tc = src_points - src_centroid, tgt_points - tgt_centroid
    H = sc.permute(0, 2, 1) @ torch.diag_embed(weights) @ tc                                         # :196-197
    U, _, V = torch.svd(H)                                                                           # :198
    Ut = U.transpose(1, 2)
    eye = torch.eye(3, dtype=H.dtype).unsqueeze(0).repeat(B, 1, 1)
    eye[:, -1, -1] = torch.sign(torch.det(V @ Ut))                                                   # :201
    R = V @ eye @ Ut
    t = (tgt_centroid.permute(0, 2, 1) - R @ src_centroid.permute(0, 2, 1)).squeeze(2)               # :204-205
    return (R.squeeze(0), t.squeeze(0)) if squeeze else (R, t)
