"""One launch of the dense-layer kernel at a given shape, for ncu. usage: prof_gemm.py M N K"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roitr_b200 import ops, engine
DEV = "cuda:0"
M, N, K = (int(x) for x in sys.argv[1:4])
a = torch.randn(M, K, device=DEV); w = torch.randn(N, K, device=DEV); b = torch.randn(N, device=DEV); out = torch.empty(M, N, device=DEV)
wp = engine.pack_linear_tc(w)
ops.linear(a, w, b, out=out, wpack=wp)
torch.cuda.synchronize()
