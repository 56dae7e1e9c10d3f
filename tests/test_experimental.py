"""Opt-in checks of code paths that are compiled but NOT enabled by default (ROITR_EXPERIMENTAL=1 pytest tests/test_experimental.py
on a GPU box). They are excluded from `-m gpu` and skipped otherwise; enable a path in the engine only after its check is green."""
import pytest
import torch

pytestmark = pytest.mark.experimental


@pytest.mark.parametrize("M,K,pre,post,relu", [(9984, 256, True, False, False), (4992, 512, True, False, False), (40000, 256, False, True, True)])
def test_fused_layernorm_256_column_tile(M, K, pre, post, relu):
    """engine.LN256: LayerNorm of a 256-channel layer in the epilogue of a 256-column tile == linear + row_epilogue."""
    from roitr_b200 import engine, ops
    g = torch.Generator().manual_seed(M + K)
    a = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(256, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(256, generator=g).cuda()
    gamma, beta = (1 + 0.1 * torch.randn(256, generator=g)).cuda(), (0.1 * torch.randn(256, generator=g)).cuda()
    res = torch.randn(M, 256, generator=g).cuda()
    y = ops.linear_ln(a, w, b, engine.pack_linear_tc(w), gamma, beta, res_pre=res if pre else None, res_post=res if post else None,
                      relu=relu, wpack_wide=engine.pack_linear_tc(w, 256))
    t = ops.linear(a, w, b, wpack=engine.pack_linear_tc(w))
    two = ops.row_epilogue(t, res_pre=res if pre else None, gamma=gamma, beta=beta, res_post=res if post else None,
                           mode=ops.MODE_LN | (ops.MODE_RELU if relu else 0))
    assert (y - two).abs().max().item() <= 1e-5 * max(1.0, two.abs().max().item())
