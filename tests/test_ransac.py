"""Correspondence RANSAC (registration/benchmark_utils.py:165-209): the numpy oracle's known-answer
behaviour on the CPU; the CUDA kernel against the oracle (same seed = same 50 000 hypotheses) and against the ground-truth
transform of synthetic correspondences on the GPU."""
import numpy as np
import pytest
import torch

from oracle import ransac_ref


def _problem(seed, n, inlier_frac, noise=0.01):
    g = np.random.default_rng(seed)
    q = g.normal(size=4); q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    t = g.uniform(-0.5, 0.5, size=3)
    src = g.uniform(-1.5, 1.5, size=(n, 3))
    tgt = src @ R.T + t + noise * g.normal(size=(n, 3))
    out = g.random(n) >= inlier_frac
    tgt[out] = g.uniform(-1.5, 1.5, size=(int(out.sum()), 3))
    return src.astype(np.float32), tgt.astype(np.float32), R, t, ~out


def _rot_err_deg(Ra, Rb):
    return float(np.degrees(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1))))


def test_hash_is_uniform_and_deterministic():
    a = ransac_ref.sample_indices(7, 0, 50000, 1000)
    b = ransac_ref.sample_indices(7, 0, 50000, 1000)
    assert np.array_equal(a, b) and a.min() >= 0 and a.max() < 1000
    assert not np.array_equal(a, ransac_ref.sample_indices(8, 0, 50000, 1000))
    assert not np.array_equal(a, ransac_ref.sample_indices(7, 1, 50000, 1000))
    hist = np.bincount(a.ravel(), minlength=1000)
    assert hist.min() > 80 and hist.max() < 230                   # 150 expected per bin


def test_umeyama_recovers_exact_transform():
    src, tgt, R, t, _ = _problem(0, 12, 1.0, noise=0.0)
    Rh, th = ransac_ref.umeyama_rigid(src[None].astype(np.float64), tgt[None].astype(np.float64))
    assert np.abs(Rh[0] - R).max() < 1e-5 and np.abs(th[0] - t).max() < 1e-5 and abs(np.linalg.det(Rh[0]) - 1) < 1e-9


def test_oracle_recovers_ground_truth_under_outliers():
    src, tgt, R, t, inl = _problem(1, 1000, 0.3)
    r = ransac_ref.ransac_correspondences(src, tgt, 0.05, 50000, seed=3)
    T = r["transformation"]
    assert _rot_err_deg(T[:3, :3], R) < 1.5 and np.abs(T[:3, 3] - t).max() < 0.03
    assert r["fitness"] >= 0.9 * inl.mean() and 0 < r["inlier_rmse"] < 0.05 and 0 <= r["best_itr"] < 50000


def test_oracle_degenerate_inputs():
    assert np.array_equal(ransac_ref.ransac_correspondences(np.zeros((2, 3)), np.zeros((2, 3)))["transformation"], np.eye(4))
    src, tgt, *_ = _problem(2, 200, 0.0)                            # no inliers at all: nothing passes the checkers (or a tiny fit)
    r = ransac_ref.ransac_correspondences(src, tgt, 0.05, 20000)
    assert r["fitness"] < 0.05


@pytest.mark.gpu
@pytest.mark.parametrize("n,frac", [(1000, 0.3), (250, 0.5), (5000, 0.15), (37, 0.8)])
def test_gpu_ransac_equals_oracle_hypothesis_for_hypothesis(n, frac):
    from roitr_b200 import registration
    src, tgt, R, t, _ = _problem(10 + n, n, frac)
    T, fit, rmse, itr = registration.ransac_batch([torch.from_numpy(src).cuda()], [torch.from_numpy(tgt).cuda()], 0.05, seed=5)
    r = ransac_ref.ransac_correspondences(src, tgt, 0.05, 50000, seed=5)
    assert int(itr[0]) == r["best_itr"], (int(itr[0]), r["best_itr"], float(fit[0]), r["fitness"])
    assert abs(float(fit[0]) - r["fitness"]) < 1e-12 and abs(float(rmse[0]) - r["inlier_rmse"]) < 1e-9
    assert np.abs(T[0].cpu().numpy() - r["transformation"]).max() < 1e-9
    assert _rot_err_deg(T[0, :3, :3].cpu().numpy(), R) < 3.0 and np.abs(T[0, :3, 3].cpu().numpy() - t).max() < 0.05


@pytest.mark.gpu
def test_gpu_ransac_batched_pairs_and_reference_signature():
    from roitr_b200 import registration
    probs = [_problem(100 + i, n, 1.0 if n == 3 else 0.4) for i, n in enumerate((1000, 640, 3, 2000))]
    T, fit, rmse, itr = registration.ransac_batch([torch.from_numpy(p[0]).cuda() for p in probs],
                                                  [torch.from_numpy(p[1]).cuda() for p in probs], 0.05, seed=9)
    for i, p in enumerate(probs):
        r = ransac_ref.ransac_correspondences(p[0], p[1], 0.05, 50000, seed=9, pair=i)
        if p[0].shape[0] == 3:
            # n = 3: every hypothesis that draws the three distinct rows fits them exactly; which of those ties wins is decided
            # by 1e-16-level rounding of a zero rmse (Jacobi vs SVD), so compare the result, not the iteration
            assert abs(float(fit[i]) - r["fitness"]) < 1e-12 and np.abs(T[i].cpu().numpy() - r["transformation"]).max() < 1e-5
            continue
        assert int(itr[i]) == r["best_itr"] and np.abs(T[i].cpu().numpy() - r["transformation"]).max() < 1e-9
    # the reference's call: (src_pcd, tgt_pcd, correspondences (c,2)) -> (4,4) float64 numpy
    src, tgt, R, t, _ = probs[0]
    corr = torch.arange(src.shape[0])[:, None].expand(-1, 2)
    Tn = registration.ransac_pose_estimation_correspondences(torch.from_numpy(src), torch.from_numpy(tgt), corr)
    assert isinstance(Tn, np.ndarray) and Tn.shape == (4, 4) and Tn.dtype == np.float64
    assert _rot_err_deg(Tn[:3, :3], R) < 1.5 and np.abs(Tn[:3, 3] - t).max() < 0.03
    with pytest.raises(NotImplementedError):
        registration.ransac_pose_estimation_correspondences(src, tgt, corr, mutual=True)


@pytest.mark.gpu
def test_gpu_ransac_on_forward_output():
    """End of the path: forward -> subsample by confidence -> RANSAC, as evaluate_registration_c2f.py:76-88 chains them."""
    from roitr_b200 import model, registration
    from roitr_b200.synthetic import forward_args, synthetic_pair
    from tests.helpers import CONFIG_3D, weights
    m = model.create_model(CONFIG_3D)
    m.load_state_dict(weights(1))
    m = m.cuda().eval()
    out = m(*forward_args(synthetic_pair(0, 4096), "cuda:0"))
    g = torch.Generator(device="cuda").manual_seed(0)
    sel = registration.sample_correspondences(out["corr_scores"], 1000, generator=g)
    assert sel.shape[0] == min(1000, out["corr_scores"].shape[0]) and sel.unique().shape[0] == sel.shape[0]
    s, t = out["src_corr_points"][sel].contiguous(), out["tgt_corr_points"][sel].contiguous()
    T, fit, rmse, itr = registration.ransac_batch([s], [t], 0.05, seed=1)
    r = ransac_ref.ransac_correspondences(s.cpu().numpy(), t.cpu().numpy(), 0.05, 50000, seed=1)
    assert int(itr[0]) == r["best_itr"] and np.abs(T[0].cpu().numpy() - r["transformation"]).max() < 1e-9
    assert torch.isfinite(T).all() and abs(float(torch.det(T[0, :3, :3])) - 1.0) < 1e-9
