"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md). Never imported by roitr_b200/.

numpy (float64) restatement of the correspondence RANSAC the reference's evaluator runs on the forward's output:
``ransac_pose_estimation_correspondences`` (registration/benchmark_utils.py:165-209, called from
registration/evaluate_registration_c2f.py:88), i.e. Open3D's ``registration_ransac_based_on_correspondence`` with
``TransformationEstimationPointToPoint(False)``, ``ransac_n = 3``, checkers ``EdgeLength(0.9)`` + ``Distance(thr)`` and
``RANSACConvergenceCriteria(50000, 1000)``.

PARITY UNPINNED for the Open3D part: the algorithm lives in the third-party dependency ``open3d==0.13.0``
(requirements.txt:64), which is not in /root/reference and not installable here, so its published algorithm
(open3d/pipelines/registration/Registration.cpp, RegistrationRANSACBasedOnCorrespondence / EvaluateRANSACBasedOnCorrespondence;
CorrespondenceChecker.cpp; Eigen::umeyama) is restated:

  for itr in range(max_iteration):                      # confidence 1000 clamps to 1.0: log(1 - 1) = -inf, no early exit
      sample ransac_n correspondences uniformly WITH replacement
      T = umeyama(src[sample] -> tgt[sample], scaling off)
      reject unless all sampled pairs (i, j): |s_i - s_j| >= 0.9 |t_i - t_j| and |t_i - t_j| >= 0.9 |s_i - s_j|
      reject unless all sampled i: |T s_i - t_i| <= thr
      inliers = {i : |T s_i - t_i| < thr};  fitness = |inliers| / n;  rmse = sqrt(sum d^2 / |inliers|)
      keep T if fitness > best.fitness or (fitness == best.fitness and rmse < best.rmse)

Open3D draws from a process-global Mersenne twister inside an OpenMP loop, so its output is not reproducible even against
itself. Here the samples are a counter-based hash of (seed, pair, 3 itr + j) shared with csrc/ransac.cu, and ties go to
the lower iteration: with the same seed the CUDA kernel and this file evaluate the SAME 50 000 hypotheses.
"""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def _hash(seed, counter):
    """PCG-RXS-M-XS-32 output function over a Weyl-style state; counter: uint64 array holding 32-bit values."""
    h = (counter * np.uint64(747796405) + np.uint64(seed) * np.uint64(2891336453) + np.uint64(1)) & M32
    h = (((h >> ((h >> np.uint64(28)) + np.uint64(4))) ^ h) * np.uint64(277803737)) & M32
    return ((h >> np.uint64(22)) ^ h) & M32


def sample_indices(seed, pair, iterations, n):
    """(iterations, 3) int64: the rows hypothesis itr draws (csrc/ransac.cu ransac_kernel)."""
    pair_seed = (np.uint64(seed) + np.uint64(0x9E3779B9) * np.uint64(pair)) & M32
    c = (np.arange(iterations, dtype=np.uint64)[:, None] * np.uint64(3) + np.arange(3, dtype=np.uint64)[None, :]) & M32
    return ((_hash(pair_seed, c) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


def umeyama_rigid(s, t):
    """Batched Eigen::umeyama without scaling: s, t (H, k, 3) -> R (H,3,3), tr (H,3) with t ~ R s + tr."""
    ms, mt = s.mean(1, keepdims=True), t.mean(1, keepdims=True)
    sigma = np.einsum("hki,hkj->hij", t - mt, s - ms) / s.shape[1]
    U, _, Vt = np.linalg.svd(sigma)
    S = np.ones((s.shape[0], 3))
    S[:, 2] = np.sign(np.linalg.det(U) * np.linalg.det(Vt))
    S[S[:, 2] == 0, 2] = 1.0
    R = np.einsum("hij,hj,hjk->hik", U, S, Vt)
    return R, mt[:, 0] - np.einsum("hij,hj->hi", R, ms[:, 0])


def ransac_correspondences(src, tgt, distance_threshold=0.05, iterations=50000, seed=0, pair=0, edge_similarity=0.9,
                           chunk=2048):
    """src, tgt (n,3): correspondence i = row i of both. -> dict(transformation (4,4) f64, fitness, inlier_rmse, best_itr,
    inliers)."""
    src, tgt = np.asarray(src, dtype=np.float64), np.asarray(tgt, dtype=np.float64)
    n = src.shape[0]
    out = dict(transformation=np.eye(4), fitness=0.0, inlier_rmse=0.0, best_itr=-1, inliers=0)
    if n < 3:
        return out
    idx = sample_indices(seed, pair, iterations, n)
    s, t = src[idx], tgt[idx]                                       # (I,3,3)
    ok = np.ones(iterations, dtype=bool)
    for i, j in ((0, 1), (0, 2), (1, 2)):
        ds = np.sqrt(((s[:, i] - s[:, j]) ** 2).sum(-1))
        dt = np.sqrt(((t[:, i] - t[:, j]) ** 2).sum(-1))
        ok &= (ds >= dt * edge_similarity) & (dt >= ds * edge_similarity)
    cand = np.nonzero(ok)[0]
    if cand.size == 0:
        return out
    R, tr = umeyama_rigid(s[cand], t[cand])
    d = np.sqrt(((np.einsum("hij,hkj->hki", R, s[cand]) + tr[:, None] - t[cand]) ** 2).sum(-1))
    keep = (d <= distance_threshold).all(1)
    cand, R, tr = cand[keep], R[keep], tr[keep]
    best = (-1, 0.0, -1, None, None)
    for c0 in range(0, cand.size, chunk):
        Rc, tc = R[c0:c0 + chunk], tr[c0:c0 + chunk]
        d2 = ((np.einsum("hij,nj->hni", Rc, src) + tc[:, None] - tgt[None]) ** 2).sum(-1)      # (h, n)
        inl = np.sqrt(d2) < distance_threshold
        cnt = inl.sum(1)
        e2 = (d2 * inl).sum(1)
        for h in np.nonzero(cnt >= max(best[0], 1))[0]:
            key = (int(cnt[h]), float(e2[h]), int(cand[c0 + h]))
            if key[0] > best[0] or (key[0] == best[0] and (key[1] < best[1] or (key[1] == best[1] and key[2] < best[2]))):
                best = (key[0], key[1], key[2], Rc[h], tc[h])
    if best[0] <= 0:
        return out
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = best[3], best[4]
    return dict(transformation=T, fitness=best[0] / n, inlier_rmse=float(np.sqrt(best[1] / best[0])), best_itr=best[2], inliers=best[0])
