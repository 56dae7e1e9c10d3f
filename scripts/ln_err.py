import sys; sys.path.insert(0, '/root/repo')
import torch
from roitr_b200 import ops, engine
DEV='cuda:0'
for (M,N,K) in [(20000,64,64),(20000,128,128)]:
    g = torch.Generator().manual_seed(1)
    a = (3*torch.randn(M, K, generator=g)).to(DEV); w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV); b = torch.randn(N, generator=g).to(DEV)
    gamma, beta = (1 + 0.1 * torch.randn(N, generator=g)).to(DEV), (0.1 * torch.randn(N, generator=g)).to(DEV)
    res = (3*torch.randn(M, N, generator=g)).to(DEV)
    wp = engine.pack_linear_tc(w)
    y = ops.linear_ln(a, w, b, wp, gamma, beta, res_pre=res)
    t32 = ops.linear(a, w, b, wpack=wp)
    two = ops.row_epilogue(t32, res_pre=res, gamma=gamma, beta=beta, mode=ops.MODE_LN)
    t = a.double() @ w.double().t() + b.double() + res.double()
    ref = torch.nn.functional.layer_norm(t, (N,), gamma.double(), beta.double(), 1e-5)
    print(M,N,K, "fused err %.3e  two-kernel err %.3e  fused-two %.3e" % ((y.double()-ref).abs().max().item(), (two.double()-ref).abs().max().item(), (y-two).abs().max().item()))
