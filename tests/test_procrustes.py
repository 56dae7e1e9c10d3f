"""Weighted Procrustes: oracle pinned to the reference's own outputs; CUDA kernel against both."""
import os

import numpy as np
import pytest
import torch

from oracle import procrustes_ref
from oracle import reference_shim as rs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "procrustes_golden.npz")


def _cases():
    z = np.load(GOLD)
    for i in range(int(z["n"])):
        yield (torch.from_numpy(z["src%d" % i]), torch.from_numpy(z["tgt%d" % i]), torch.from_numpy(z["w%d" % i]), float(z["thr%d" % i]),
               torch.from_numpy(z["R%d" % i]), torch.from_numpy(z["t%d" % i]))


def test_oracle_matches_reference_golden():
    for src, tgt, w, thr, R, t in _cases():
        R2, t2 = procrustes_ref.weighted_procrustes(src, tgt, w, thr)
        assert torch.allclose(R2, R, atol=2e-6) and torch.allclose(t2, t, atol=5e-6)
        assert torch.allclose(torch.det(R2), torch.ones(R2.shape[0]), atol=1e-5)


@pytest.mark.skipif(not rs.available(), reason="/root/reference not present (GPU box)")
def test_oracle_matches_live_reference():
    rs.install()
    from lib.utils import weighted_procrustes
    for src, tgt, w, thr, _, _ in _cases():
        R, t = weighted_procrustes(src, tgt, w, weight_thresh=thr)
        R2, t2 = procrustes_ref.weighted_procrustes(src, tgt, w, thr)
        assert torch.equal(R, R2) and torch.equal(t, t2)
        T = weighted_procrustes(src[0], tgt[0], w[0], weight_thresh=thr, return_transform=True)
        assert T.shape == (4, 4)


def test_oracle_reflection_case_gives_proper_rotation():
    g = torch.Generator().manual_seed(1)
    src = torch.randn(1, 50, 3, generator=g)
    tgt = src * torch.tensor([1.0, 1.0, -1.0])          # a mirror image: the best ROTATION is wanted, not the reflection
    R, _ = procrustes_ref.weighted_procrustes(src, tgt)
    assert torch.allclose(torch.det(R), torch.ones(1), atol=1e-5)


@pytest.mark.gpu
def test_gpu_matches_reference_golden_and_oracle():
    from roitr_b200 import registration
    for src, tgt, w, thr, R, t in _cases():
        R2, t2 = registration.weighted_procrustes(src.cuda(), tgt.cuda(), w.cuda(), weight_thresh=thr)
        assert (R2.cpu() - R).abs().max().item() <= 5e-6, (R2.cpu() - R).abs().max().item()
        assert (t2.cpu() - t).abs().max().item() <= 2e-5
    # unbatched call, no weights, transform form (lib/utils.py:176-183,207-213)
    src, tgt, w, thr, R, t = next(_cases())
    T = registration.weighted_procrustes(src[0].cuda(), tgt[0].cuda(), return_transform=True).cpu()
    Ro, to = procrustes_ref.weighted_procrustes(src[0], tgt[0])
    assert T.shape == (4, 4) and (T[:3, :3] - Ro).abs().max().item() <= 5e-6 and (T[:3, 3] - to).abs().max().item() <= 2e-5


@pytest.mark.gpu
def test_gpu_registration_from_forward_correspondences():
    """End of the pipeline: correspondences of a forward pass -> pose; recovers the synthetic ground-truth transform."""
    from roitr_b200 import model, registration
    from roitr_b200.synthetic import forward_args, synthetic_pair
    from tests.helpers import CONFIG_3D, weights
    pair = synthetic_pair(0, 4096)
    m = model.create_model(CONFIG_3D)
    m.load_state_dict(weights(1))
    m = m.cuda().eval()
    out = m(*forward_args(pair, "cuda:0"))
    R, t = registration.weighted_procrustes(out["src_corr_points"], out["tgt_corr_points"], out["corr_scores"])
    Ro, to = procrustes_ref.weighted_procrustes(out["src_corr_points"].cpu(), out["tgt_corr_points"].cpu(), out["corr_scores"].cpu())
    assert (R.cpu() - Ro).abs().max().item() <= 1e-5 and (t.cpu() - to).abs().max().item() <= 5e-5
