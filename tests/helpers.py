"""Shared test helpers: golden loading, seeded weights, comparison with flip budgets."""
import json
import os

import numpy as np
import torch

from roitr_b200.synthetic import forward_args, seeded_state_dict, synthetic_pair

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CONFIG_3D = dict(with_cross_pos_embed=True, benchmark="3DLoMatch", num_est_coarse_corr=256,
                 transformer_architecture=["self", "cross", "self", "cross", "self", "cross"], mode="test",
                 point_per_patch=64, matching_radius=0.05, num_gt_coarse_corr=128, coarse_overlap_threshold=0.1,
                 fine_matching_topk=3, fine_matching_mutual=True, fine_matching_confidence_threshold=0.05,
                 fine_matching_use_dustbin=False, fine_matching_use_global_score=False,
                 fine_matching_correspondence_threshold=3)
CONFIG_4D = dict(CONFIG_3D, benchmark="4DLoMatch", num_est_coarse_corr=128, fine_matching_topk=2)


def schema(factor=1):
    return json.load(open(os.path.join(GOLDEN, "state_dict_schema_f%d.json" % factor)))


def weights(factor=1, seed=42, fine_scale=8.0):
    return seeded_state_dict(schema(factor), seed, fine_scale=fine_scale)


def baseline_pair(name):
    """The two BASELINE.json configurations that are not the headline one, with UNEQUAL cloud sizes:
    "3dmatch_30k" = config 3 (2 x ~30k points -> 468 / 437 superpoints), "4dmatch_8k" = config 5 (2 x ~8k points, non-rigid
    source, factor-2 backbone, adaptive head)."""
    n_src, n_tgt, four_d, index = {"3dmatch_30k": (30000, 28000, False, 30), "4dmatch_8k": (8000, 7000, True, 31)}[name]
    p = dict(synthetic_pair(index, n_src, deform=four_d))
    for k in ("tgt_pcd", "tgt_feats", "tgt_normals"):
        p[k] = p[k][:n_tgt].contiguous()
    return p


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta


def golden_case(name):
    z, meta = load_golden(name)
    pair = synthetic_pair(meta["pair_index"], meta["n"], deform=meta["deform"])
    cfg = CONFIG_4D if meta["factor"] == 2 else CONFIG_3D
    return z, meta, pair, cfg, weights(meta["factor"], meta["seed"])


def knn_equal_up_to_ties(idx_a, d_a, idx_b, d_b):
    """Bit-exact kNN comparison that tolerates only permutations among exactly-equal distances
    (the reference's heap order among equal d2 is unspecified)."""
    idx_a, idx_b = np.asarray(idx_a), np.asarray(idx_b)
    d_a, d_b = np.asarray(d_a), np.asarray(d_b)
    if not np.array_equal(d_a, d_b):
        return False
    bad = idx_a != idx_b
    if not bad.any():
        return True
    # every mismatching slot must sit in a run of equal distances holding the same index multiset
    for r in np.unique(np.nonzero(bad)[0]):
        for dv in np.unique(d_a[r][bad[r]]):
            sel = d_a[r] == dv
            if sorted(idx_a[r][sel].tolist()) != sorted(idx_b[r][sel].tolist()):
                return False
    return True
