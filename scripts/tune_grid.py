"""Grid cell size vs query type at the bench step's shapes (32 clouds per launch): ms per launch for every kNN query of the
forward with the reference set's grid built at different average points per cell. usage: tune_grid.py (GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roitr_b200 import ops
from roitr_b200.synthetic import synthetic_pair
DEV = "cuda:0"
B = 16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
def timeit(fn, iters=5):
    for _ in range(2): fn()
    ts = []
    for _ in range(iters):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
pairs = [synthetic_pair(i, 20000) for i in range(B)]
pts = torch.cat([p["src_raw_pcd"] for p in pairs] + [p["tgt_pcd"] for p in pairs]).to(DEV)
nrm = torch.cat([p["src_normals"] for p in pairs] + [p["tgt_normals"] for p in pairs]).to(DEV)
def offs(n): return torch.tensor([n * (i + 1) for i in range(2 * B)], dtype=torch.int32, device=DEV)
L = [dict(p=pts, n=nrm, o=offs(20000), sz=20000)]
for li in range(1, 4):
    prev = L[-1]; sz = prev["sz"] // 4
    idx, p_ = ops.fps(prev["p"], prev["o"], offs(sz), prev["sz"], sz * 2 * B)
    L.append(dict(p=p_, n=ops.gather_rows(prev["n"], idx), o=offs(sz), sz=sz))
targets = (0.2, 0.3, 0.4, 0.5, 0.7, 1.0, 1.5)
print("query (refs <- queries, k)            " + "  ".join("t=%.1f" % t for t in targets))
def row(name, ref, qry, k, drop, ppf):
    out = []
    for t in targets:
        g = ops.knn_grid_build(ref["p"], ref["o"], t)
        qg = g if qry is ref else ops.knn_grid_build(qry["p"], qry["o"])
        out.append(timeit(lambda: ops.knn_ppf(k, ref["p"], ref["n"] if ppf else None, qry["p"], qry["n"] if ppf else None, ref["o"], qry["o"],
                                              drop_first=drop, want_ppf=ppf, want_dist=not ppf, grid=g, qgrid=qg)))
    print("%-38s" % name + "  ".join("%5.3f" % v for v in out))
row("L1 self k=8 (9 slots)", L[0], L[0], 8, 1, True)
row("L2 <- L1 refs, k=16 (17 slots)", L[0], L[1], 16, 1, True)
row("L2 self k=16", L[1], L[1], 16, 1, True)
row("L3 <- L2 refs, k=16", L[1], L[2], 16, 1, True)
row("L3 self k=16", L[2], L[2], 16, 1, True)
row("L1 queries in L2 refs, 3-NN", L[1], L[0], 3, 0, False)
row("L2 queries in L3 refs, 3-NN", L[2], L[1], 3, 0, False)
row("L1 self 1-NN (occlusion-like)", L[0], L[0], 1, 0, False)
for t in targets:
    print("grid build L1 t=%.1f: %.3f ms" % (t, timeit(lambda: ops.knn_grid_build(L[0]["p"], L[0]["o"], t))))
row("L3 queries in L4 refs, 3-NN", L[3], L[2], 3, 0, False)
row("L4 <- L3 refs, k=16", L[2], L[3], 16, 1, True)
