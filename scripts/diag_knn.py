import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import native
from roitr_b200 import pointops, pointops_cuda
DEV = "cuda:0"
def cloud(n, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(n, 3, generator=g) * 2 - 1).contiguous()
for (n, m, ns) in [(1024, 1024, 9), (2048, 2048, 9), (5000, 1250, 17), (37, 37, 17), (20000, 20000, 9)]:
    xyz = cloud(n, 7); q = xyz[:m].contiguous()
    off = torch.tensor([n], dtype=torch.int32); noff = torch.tensor([m], dtype=torch.int32)
    idx_o, d2_o = native.knn(ns, xyz, q, off, noff)
    idx2 = torch.zeros(m, ns, dtype=torch.int32, device=DEV); d2 = torch.zeros(m, ns, device=DEV)
    pointops_cuda.knnquery_cuda(m, ns, xyz.to(DEV), q.to(DEV), off.to(DEV), noff.to(DEV), idx2, d2)
    torch.cuda.synchronize()
    d2c, ic = d2.cpu(), idx2.cpu()
    bad = (d2c != d2_o)
    print(n, m, ns, "d2 mismatches", int(bad.sum()), "idx mismatches", int((ic != idx_o).sum()))
    if bad.any():
        r, c = [int(v[0]) for v in torch.nonzero(bad, as_tuple=True)]
        print("  first bad row", r, "col", c, "gpu", d2c[r].tolist(), ic[r].tolist(), "\n   oracle", d2_o[r].tolist(), idx_o[r].tolist())
    s_g = torch.sqrt(d2).cpu(); s_c = torch.sqrt(d2c)
    _, dist = pointops.knnquery(ns, xyz.to(DEV), q.to(DEV), off.to(DEV), noff.to(DEV))
    print("  sqrt: kernel vs torch-cuda", int((dist.cpu() != s_g).sum()), " kernel vs torch-cpu", int((dist.cpu() != s_c).sum()), " torch-cuda vs torch-cpu", int((s_g != s_c).sum()))
