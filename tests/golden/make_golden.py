"""Generates the golden vectors in this directory by executing the UNMODIFIED reference (/root/reference) on CPU
under oracle/reference_shim.py. Run from the repo root, in the build container only:

    python tests/golden/make_golden.py

The reference has no tests or fixtures of its own, so these are the known-answer vectors for the
path: seeded synthetic pair (roitr_b200.synthetic.synthetic_pair) + seeded weights (seeded_state_dict over the
reference's own state_dict schema, loaded with strict=True into the reference model).

Outputs (kept small; large tensors are strided samples plus full-tensor float64 checksums):
    state_dict_schema_f1.json / _f2.json    (name, shape) list of RIGA_v2(...).state_dict(), factor 1 / 2
    golden_3dmatch_n1024.npz                config 1 of BASELINE.json (2x1024 pts, 3DMatch head)
    golden_3dmatch_n4096.npz                larger case (64 superpoints -> top-256 of 4096 actually selects)
    golden_4dmatch_n1024.npz                4DMatch head (factor 2, adaptive superpoint matching)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_shim as rs  # noqa: E402
from roitr_b200.synthetic import forward_args, seeded_state_dict, synthetic_pair  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROW_STRIDE = 8      # point-feature rows kept
PATCH_KEEP = 8      # matching_scores patches kept in full


def dump_schema():
    for factor, bench in ((1, "3DLoMatch"), (2, "4DLoMatch")):
        m = rs.create_reference_model(rs.default_config(bench))
        schema = [(k, list(v.shape)) for k, v in m.state_dict().items()]
        with open(os.path.join(HERE, "state_dict_schema_f%d.json" % factor), "w") as f:
            json.dump(schema, f)


def run_case(name, bench, factor, pair_index, n, deform):
    schema = json.load(open(os.path.join(HERE, "state_dict_schema_f%d.json" % factor)))
    cfg = rs.default_config(bench)
    model = rs.create_reference_model(cfg)
    model.load_state_dict(seeded_state_dict(schema, 42), strict=True)
    pair = synthetic_pair(pair_index, n, deform=deform)
    with torch.no_grad(), rs.Trace() as tr:
        out = model(*forward_args(pair))
    g = dict(meta=np.array(json.dumps(dict(bench=bench, factor=factor, pair_index=pair_index, n=n, deform=deform,
                                             seed=42, row_stride=ROW_STRIDE, patch_keep=PATCH_KEEP))))
    for i, t in enumerate(tr.fps):
        g["fps_%d" % i] = t.numpy()
    for i, (ns, idx, dist) in enumerate(tr.knn):
        if idx.shape[0] <= 1100:             # keep every call at n=1024; only small ones for bigger clouds
            g["knn_%d_idx" % i] = idx.numpy()
            g["knn_%d_dist" % i] = dist.numpy()
        g["knn_%d_idxsum" % i] = np.array([ns, idx.shape[0], int(idx.long().sum())], dtype=np.int64)
    for k, v in out.items():
        a = v.detach().numpy()
        if a.dtype.kind == "f":
            g["sum64_" + k] = np.array(a.astype(np.float64).sum())
        if k in ("src_point_feats", "tgt_point_feats"):
            a = a[::ROW_STRIDE]
        elif k == "matching_scores":
            g["rowsum64_" + k] = a.astype(np.float64).reshape(a.shape[0], -1).sum(1)
            a = a[:PATCH_KEEP]
        elif k.endswith("knn_points") or k.endswith("knn_masks"):
            a = a[:PATCH_KEEP]
        elif k in ("src_points", "tgt_points"):
            continue                        # inputs, regenerated from the seed
        g[k] = a
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **g)
    print(name, "%.2f MB" % (os.path.getsize(path) / 1e6), "corr", out["corr_scores"].shape[0])


if __name__ == "__main__":
    dump_schema()
    run_case("golden_3dmatch_n1024", "3DLoMatch", 1, 0, 1024, False)
    run_case("golden_3dmatch_n4096", "3DLoMatch", 1, 1, 4096, False)
    run_case("golden_4dmatch_n1024", "4DLoMatch", 2, 2, 1024, True)
