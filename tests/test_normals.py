"""Normal estimation: oracle sanity on CPU, CUDA kernel against the oracle on the GPU."""
import numpy as np
import pytest
import torch

from oracle import normals_ref
from roitr_b200.synthetic import synthetic_pair


def test_oracle_plane_and_orientation():
    g = np.random.default_rng(0)
    xy = g.uniform(-1, 1, size=(2000, 2))
    pts = np.stack([xy[:, 0], xy[:, 1], 0.3 * xy[:, 0] + 2.0 + 1e-4 * g.standard_normal(2000)], 1)
    n, gap = normals_ref.estimate_normals(pts, 33, (0.0, 0.0, 0.0))
    ref = np.array([0.3, 0.0, -1.0]) / np.linalg.norm([0.3, 0.0, -1.0])          # plane normal facing the origin (z = 2 plane above it)
    assert np.abs(n @ ref).min() > 0.999
    assert (np.sum((0.0 - pts) * n, 1) >= 0).all()                                # normal_redirect: towards the view point
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-6)
    assert (gap > 0.1).all()
    # redirect flips exactly the normals that look away
    flipped = normals_ref.normal_redirect(pts, -n.astype(np.float64), np.zeros(3))
    assert np.allclose(flipped, n, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("n_pts,knn", [(2048, 33), (20000, 33), (5000, 17), (20, 33)])
def test_gpu_normals_match_oracle(n_pts, knn):
    from roitr_b200 import preprocess
    pair = synthetic_pair(3, max(n_pts, 64))
    pts = pair["tgt_pcd"][:n_pts].contiguous()
    vp = (0.3, -0.2, 4.0)
    ref, gap = normals_ref.estimate_normals(pts.numpy(), knn, vp)
    out = preprocess.estimate_normals(pts.cuda(), knn=knn, view_point=vp).cpu().numpy()
    assert np.allclose(np.linalg.norm(out, axis=1), 1.0, atol=1e-5)
    dot = np.sum(out * ref, 1)
    ok = gap > 1e-3                                      # direction well defined
    facing = np.abs(np.sum((np.asarray(vp) - pts.numpy()) * ref, 1)) > 1e-4     # sign well defined
    assert ok.mean() > 0.95
    assert (np.abs(dot[ok]) > 1 - 1e-5).all(), float(np.abs(dot[ok]).min())
    assert (dot[ok & facing] > 0).all()


@pytest.mark.gpu
def test_gpu_normals_segmented_batch_equals_single_clouds():
    from roitr_b200 import preprocess
    a, b = synthetic_pair(1, 4096)["src_pcd"], synthetic_pair(2, 3000)["tgt_pcd"][:3000]
    both = torch.cat([a, b]).cuda()
    off = torch.tensor([a.shape[0], a.shape[0] + b.shape[0]], dtype=torch.int32, device="cuda")
    n_both = preprocess.estimate_normals(both, off)
    n_a, n_b = preprocess.estimate_normals(a.cuda()), preprocess.estimate_normals(b.cuda())
    assert torch.equal(n_both[:a.shape[0]], n_a) and torch.equal(n_both[a.shape[0]:], n_b)
