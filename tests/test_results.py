"""Result record: keys / order / dtypes of lib/tester.py:56-69, round trip through torch.save, and the
fields the unchanged evaluators read. Runs on the CPU with the oracle's forward output (same 22-key dict)."""
import torch

from oracle import forward_ref as fr
from roitr_b200 import results
from roitr_b200.synthetic import forward_args, synthetic_pair
from tests.helpers import CONFIG_3D, weights

REFERENCE_KEYS = ["src_raw_pcd", "src_pcd", "tgt_pcd", "src_nodes", "tgt_nodes", "src_node_desc", "tgt_node_desc", "src_point_desc",
                  "tgt_point_desc", "src_corr_pts", "tgt_corr_pts", "confidence", "gt_tgt_node_occ", "gt_src_node_occ", "rot", "trans"]


def _forward():
    pair = synthetic_pair(0, 1024)
    with torch.no_grad():
        out = fr.riga_forward(weights(1), CONFIG_3D, *forward_args(pair))
    inputs = dict(src_pcd=pair["src_pcd"], tgt_pcd=pair["tgt_pcd"], src_raw_pcd=pair["src_raw_pcd"], rot=pair["rot"], trans=pair["trans"])
    return inputs, out


def test_record_has_the_reference_layout(tmp_path):
    inputs, out = _forward()
    rec = results.tester_record(inputs, out, "3DMatch")
    assert list(rec) == REFERENCE_KEYS                                   # same keys, same order (lib/tester.py:57-66)
    assert all(torch.is_tensor(v) and not v.is_cuda for v in rec.values())
    assert torch.equal(rec["confidence"], out["corr_scores"]) and torch.equal(rec["src_corr_pts"], out["src_corr_points"])
    assert torch.equal(rec["src_point_desc"], out["src_point_feats"]) and rec["src_point_desc"].shape == (1024, 256)
    assert rec["rot"].shape == (3, 3) and rec["gt_tgt_node_occ"].dtype == out["gt_tgt_node_occ"].dtype
    # 4DMatch records carry the metric index list through untouched
    rec4 = results.tester_record(inputs, out, "4DLoMatch", metric_index=[torch.arange(5)])
    assert list(rec4) == REFERENCE_KEYS + ["metric_index_list"] and torch.equal(rec4["metric_index_list"][0], torch.arange(5))
    # file name = global pair index; round trip; the evaluator's fields are all there
    path = results.save_record(rec, str(tmp_path), "3DMatch", 1337)
    assert path.endswith("3DMatch/1337.pth")
    back = results.load_for_registration(path)
    for k in ("src_pcd", "tgt_pcd", "src_nodes", "tgt_nodes", "src_node_desc", "tgt_node_desc", "rot", "trans", "src_corr_pts",
              "tgt_corr_pts", "confidence"):
        assert torch.equal(back[k], rec[k]), k
    # the evaluator's first step (registration/evaluate_registration_c2f.py:78): sampling probabilities from the confidences
    prob = back["confidence"] / torch.sum(back["confidence"])
    assert abs(float(prob.sum()) - 1.0) < 1e-5


import pytest


@pytest.mark.gpu
def test_record_from_cuda_forward():
    """CUDA tensors travel to the host through the per-dtype staging buffer; the record equals per-tensor .cpu() copies."""
    from roitr_b200 import model
    pair = synthetic_pair(0, 1024)
    m = model.create_model(CONFIG_3D)
    m.load_state_dict(weights(1))
    m = m.cuda().eval()
    out = m(*forward_args(pair, "cuda:0"))
    inputs = {k: pair[k].cuda() for k in ("src_pcd", "tgt_pcd", "src_raw_pcd", "rot", "trans")}
    rec = results.tester_record(inputs, out, "3DMatch")
    assert list(rec) == REFERENCE_KEYS and all(not v.is_cuda for v in rec.values())
    assert torch.equal(rec["confidence"], out["corr_scores"].cpu()) and torch.equal(rec["tgt_point_desc"], out["tgt_point_feats"].cpu())
    assert torch.equal(rec["rot"], pair["rot"]) and torch.equal(rec["src_nodes"], out["src_nodes"].cpu())
