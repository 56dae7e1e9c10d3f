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


# ---- analytic known-answer tests (the Open3D part of the oracle is unpinned: these pin oracle AND kernel to geometry) ----
def _kat_sphere(n=6000, r=1.5, centre=(0.2, -0.1, 3.0), seed=1):
    g = np.random.default_rng(seed)
    v = g.standard_normal((n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return (np.asarray(centre) + r * v).astype(np.float32), v       # points, outward unit normals


def _kat_edge(n=4000, seed=2):
    """two half planes meeting at the x axis: z = 0 for y < 0 and z = y for y >= 0 (a 45 degree ridge)"""
    g = np.random.default_rng(seed)
    x, y = g.uniform(-1, 1, n), g.uniform(-1, 1, n)
    z = np.where(y < 0, 0.0, y)
    pts = np.stack([x, y, z + 2.0], 1).astype(np.float32)
    nrm = np.where((y < 0)[:, None], np.array([0.0, 0.0, 1.0]), np.array([0.0, -1.0, 1.0]) / np.sqrt(2.0))
    return pts, nrm, np.abs(y)                                       # distance from the ridge


def _check_kats(estimate):
    # sphere: normals are radial; redirected towards a view point at the centre they point INWARD
    pts, radial = _kat_sphere()
    n = estimate(pts, 33, (0.2, -0.1, 3.0))
    assert (np.sum(n * radial, 1) < -0.995).all()
    # ... and OUTWARD for a far view point on the side that sees the point
    far = np.array([0.2, -0.1, 50.0])
    n2 = estimate(pts, 33, tuple(far))
    to_vp = far - pts
    sees = np.sum(to_vp * radial, 1) / np.linalg.norm(to_vp, axis=1) > 0.1           # well inside the visible cap
    assert (np.sum(n2 * radial, 1)[sees] > 0.995).all()
    # ridge: away from the edge each face keeps its own plane normal; oriented towards the origin (below both faces)
    pts, face, dist = _kat_edge()
    n3 = estimate(pts, 33, (0.0, 0.0, 0.0))
    away = dist > 0.25
    assert (np.abs(np.sum(n3 * face, 1))[away] > 0.9999).all()
    assert (np.sum((0.0 - pts) * n3, 1) >= 0).all()
    # k nearest neighbours that are exactly collinear: covariance rank 1, any unit vector orthogonal to the line is valid
    line = np.stack([np.linspace(0, 1, 50), np.zeros(50), np.ones(50)], 1).astype(np.float32)
    n4 = estimate(line, 9, (0.0, 0.0, 0.0))
    assert np.allclose(np.linalg.norm(n4, axis=1), 1.0, atol=1e-5) and (np.abs(n4[:, 0]) < 1e-4).all()


def test_oracle_analytic_kats():
    _check_kats(lambda p, k, vp: normals_ref.estimate_normals(p, k, vp)[0])


@pytest.mark.gpu
def test_gpu_analytic_kats():
    from roitr_b200 import preprocess
    _check_kats(lambda p, k, vp: preprocess.estimate_normals(torch.from_numpy(p).cuda(), knn=k, view_point=vp).cpu().numpy())
