"""Scaling-form vs pure log-domain Sinkhorn on the same patches: where do they differ? (GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roitr_b200 import _lib, model
from roitr_b200.synthetic import forward_args, synthetic_pair
from tests.helpers import CONFIG_3D, weights
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = model.create_model(CONFIG_3D); m.load_state_dict(weights(1)); m = m.cuda().eval()
args = forward_args(synthetic_pair(0, N), "cuda:0")
outs = {}
for w in (1000, 20, 40):
    _lib.lib().roitr_debug_fine_warmup(w)
    outs[w] = m(*args)
ref = outs[1000]["matching_scores"]
km_t, km_s = outs[1000]["tgt_node_corr_knn_masks"], outs[1000]["src_node_corr_knn_masks"]
for w in (20, 40):
    ms = outs[w]["matching_scores"]
    live = ref > -1e5
    d = ((ms - ref).abs() / (1 + ref.abs()))
    d = torch.where(live, d, torch.zeros_like(d))
    print("warmup", w, "max rel diff", d.max().item(), "n>1e-3:", int((d > 1e-3).sum()), "corr", outs[w]["corr_scores"].shape[0], "vs", outs[1000]["corr_scores"].shape[0])
    i = int(d.argmax()); p, r, c = i // (65 * 65), (i // 65) % 65, i % 65
    print("  worst at patch", p, "row", r, "col", c, "ref", ref[p, r, c].item(), "got", ms[p, r, c].item(), "valid rows/cols", int(km_t[p].sum()), int(km_s[p].sum()))
    bad = (d > 1e-3)
    print("  patches with diffs:", sorted(set((bad.nonzero()[:, 0]).tolist()))[:10], " rows:", sorted(set(bad.nonzero()[:, 1].tolist()))[:12], "cols:", sorted(set(bad.nonzero()[:, 2].tolist()))[:12])
