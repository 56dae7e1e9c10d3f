"""The exact algebraic folds done at weight-packing time (engine.pack_weights) reproduce the layer-by-layer computation:
checked in fp64 on the CPU against the unfolded formulas of the reference (no GPU, no kernels involved)."""
import torch

from roitr_b200 import engine
from tests.helpers import CONFIG_3D, weights


def _packed():
    sd = weights(1)
    return engine.pack_weights(sd, torch.device("cpu"), CONFIG_3D["transformer_architecture"]), sd


def test_local_layer_folds():
    W, sd = _packed()
    g = torch.Generator().manual_seed(0)
    for p in ("backbone.enc1.1.transformer.transformer", "backbone.enc2.0.transformer", "backbone.enc3.0.transformer"):
        a = p + ".transformer.attention"
        Win, b_in = sd[p + ".in_proj.weight"].double(), sd[p + ".in_proj.bias"].double()
        x = torch.randn(50, Win.shape[1], generator=g, dtype=torch.float64)
        f = x @ Win.t() + b_in                                                   # ppftransformer.py:246
        q, k, v = (f @ sd[a + ".proj_%s.weight" % t].double().t() + sd[a + ".proj_%s.bias" % t].double() for t in "qkv")
        C = Win.shape[0]
        if (p + "#W4") in W:
            y = x @ W[p + "#W4"].double().t() + W[p + "#b4"].double()
            for got, ref in zip((y[:, :C], y[:, C:2 * C], y[:, 2 * C:3 * C], y[:, 3 * C:]), (f, q, k, v)):
                assert (got - ref).abs().max() <= 2e-6 * max(1.0, ref.abs().max())
        kv = x @ W[p + "#Wkv"].double().t() + W[p + "#bkv"].double()
        fq = x @ W[p + "#Wfq"].double().t() + W[p + "#bfq"].double()
        for got, ref in zip((kv[:, :C], kv[:, C:], fq[:, :C], fq[:, C:]), (k, v, f, q)):
            assert (got - ref).abs().max() <= 2e-6 * max(1.0, ref.abs().max())
        # positional fold: p_ij = W_p (W_e ppf + b_e) + b_p  ==  (W_p W_e) ppf + (W_p b_e + b_p)    (attention.py:176-183)
        ppf = torch.rand(20, 4, generator=g, dtype=torch.float64)
        We, be = sd[p + ".embedding.proj.weight"].double(), sd[p + ".embedding.proj.bias"].double()
        for nm in ("p", "vp"):
            ref = (ppf @ We.t() + be) @ sd[a + ".proj_%s.weight" % nm].double().t() + sd[a + ".proj_%s.bias" % nm].double()
            got = ppf @ W[p + "#A" + nm].double().t() + W[p + "#c" + nm].double()
            assert (got - ref).abs().max() <= 2e-6 * max(1.0, ref.abs().max())


def test_global_self_layer_folds():
    W, sd = _packed()
    g = torch.Generator().manual_seed(1)
    lp = "backbone.global_transformer.transformer.layers.0"
    a = lp + ".attention.attention"
    C, H = 256, engine.HEADS
    c = C // H
    x = torch.randn(30, C, generator=g, dtype=torch.float64)
    q = x @ sd[a + ".proj_q.weight"].double().t() + sd[a + ".proj_q.bias"].double()
    y = x @ W[a + "#Wqkvg"].double().t() + W[a + "#bqkvg"].double()
    assert (y[:, :C] - q).abs().max() <= 2e-6 * q.abs().max()
    # gq[n, h, :] = W_p,h^T q_h: the per-head query mapped through proj_p's rows of that head (geoattention.py:105-108)
    Wp = sd[a + ".proj_p.weight"].double()
    gq = y[:, 3 * C:].view(30, H, C)
    for h in range(H):
        ref = q[:, h * c:(h + 1) * c] @ Wp[h * c:(h + 1) * c, :]
        assert (gq[:, h] - ref).abs().max() <= 2e-6 * max(1.0, ref.abs().max())
    # position branch: pos_linear(sum_h blocks of proj_vp applied to G) folded into one matrix (geoattention.py:133, :214)
    G = torch.randn(30, H, C, generator=g, dtype=torch.float64)
    Wvp, bvp = sd[a + ".proj_vp.weight"].double(), sd[a + ".proj_vp.bias"].double()
    pos = torch.cat([G[:, h] @ Wvp[h * c:(h + 1) * c, :].t() for h in range(H)], 1) + bvp
    ref = pos @ sd[lp + ".attention.pos_linear.weight"].double().t() + sd[lp + ".attention.pos_linear.bias"].double()
    got = G.reshape(30, H * C) @ W[a + "#Wposf"].double().t() + W[a + "#bposf"].double()
    assert (got - ref).abs().max() <= 2e-6 * max(1.0, ref.abs().max())
