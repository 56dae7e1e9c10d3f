"""TEST INFRASTRUCTURE (oracle) - CPU restatement of the normal-estimation step that precedes the hot path.

Follows dataset/tdmatch.py:120-127 (Open3D ``estimate_normals(KDTreeSearchParamKNN(knn=33))`` on each cloud, then
``normal_redirect``) and dataset/common.py:312-320 (normal_redirect). Only tests/ may import this file.

PARITY UNPINNED for the Open3D part: Open3D (third-party, pinned ``open3d==0.13.0`` in requirements.txt) is not under
/root/reference and not installable here, so its published algorithm is restated: the knn nearest points of the query
(itself included), population covariance from first/second cumulants (geometry/EstimateNormals.cpp ComputeCovariance),
eigenvector of the smallest eigenvalue (ComputeNormal). The eigenvector is taken from numpy.linalg.eigh in float64, which
agrees with Open3D's closed-form solver wherever the two smallest eigenvalues are separated; the sign Open3D leaves
arbitrary is fixed by normal_redirect, restated verbatim.
"""
import numpy as np
from scipy.spatial import cKDTree


def normal_redirect(points, normals, view_point):
    """dataset/common.py:312-320."""
    vec_dot = np.sum((view_point - points) * normals, axis=-1)
    mask = vec_dot < 0.0
    out = normals.copy()
    out[mask] *= -1.0
    return out


def estimate_normals(points, knn=33, view_point=(0.0, 0.0, 0.0)):
    """-> (normals float32 (n,3), gap (n,)): gap = (l1 - l0) / l2 of the covariance spectrum (0 = direction undefined)."""
    p = np.asarray(points, dtype=np.float64)
    n = p.shape[0]
    k = min(knn, n)
    _, idx = cKDTree(p).query(p, k=k)
    idx = idx.reshape(n, k)
    nb = p[idx]                                            # (n, k, 3)
    mean = nb.mean(1)
    cov = np.einsum("nki,nkj->nij", nb, nb) / k - mean[:, :, None] * mean[:, None, :]
    w, v = np.linalg.eigh(cov)
    normals = v[:, :, 0]
    if k < 3:
        normals = np.tile(np.array([0.0, 0.0, 1.0]), (n, 1))
    gap = (w[:, 1] - w[:, 0]) / np.maximum(w[:, 2], 1e-300)
    return normal_redirect(p, normals, np.asarray(view_point, dtype=np.float64)).astype(np.float32), gap
