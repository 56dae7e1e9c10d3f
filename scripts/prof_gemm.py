"""One launch per ablation mask of the streaming dense-layer kernel, for ncu. usage: prof_gemm.py M N K mask [mask ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roitr_b200 import _lib, ops, engine
DEV = "cuda:0"
M, N, K = (int(x) for x in sys.argv[1:4])
masks = [int(x) for x in sys.argv[4:]] or [0]
a = torch.randn(M, K, device=DEV); w = torch.randn(N, K, device=DEV); b = torch.randn(N, device=DEV); out = torch.empty(M, N, device=DEV)
wp = engine.pack_linear_tc(w)
for m in masks:
    _lib.lib().roitr_debug_linear_ablate(m)
    ops.linear(a, w, b, out=out, wpack=wp)
torch.cuda.synchronize()
