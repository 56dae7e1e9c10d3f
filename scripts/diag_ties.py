import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roitr_b200 import _lib, ops
from roitr_b200.synthetic import synthetic_pair
DEV = "cuda:0"
i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
pair = synthetic_pair(0, 20000)
xyz, nrm = pair["tgt_pcd"].to(DEV), pair["tgt_normals"].to(DEV)
levels = [(xyz, nrm)]
for n in (20000, 5000, 1250):
    idx, p = ops.fps(levels[-1][0], i32([n]), i32([n // 4]), n, n // 4)
    levels.append((p, ops.gather_rows(levels[-1][1], idx)))
_lib.lib().roitr_debug_skip_knn_fixup(1)
for (li, lq, k) in [(0, 0, 8), (0, 1, 16), (1, 1, 16), (1, 2, 16), (2, 2, 16), (2, 3, 16), (3, 3, 16)]:
    (x, xn), (q, qn) = levels[li], levels[lq]
    idx, _, _ = ops.knn_ppf(k, x, xn, q, qn, i32([x.shape[0]]), i32([q.shape[0]]))
    torch.cuda.synchronize()
    print("n=%d m=%d k=%d flagged queries: %d" % (x.shape[0], q.shape[0], k, int((idx[:, 0] < 0).sum())))
