"""The dense layers (tcgen05 3xTF32 packed kernels, fp32 FFMA for K < 16), the geometric embedding and the global attention
against float64 PyTorch references."""
import pytest
import torch

from roitr_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

SHAPES = [(128, 64, 32), (256, 64, 64), (20000, 64, 64), (5000, 384, 128), (1250, 768, 256), (312, 256, 512),
          (20000, 192, 64), (777, 100, 96), (130, 24, 40), (64, 256, 256), (312, 312, 256), (4100, 256, 72)]


def _ref(a, w, b, relu):
    y = a.double() @ w.double().t() + (b.double() if b is not None else 0)
    return torch.relu(y) if relu else y


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_linear_matches_fp64(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    y = ops.linear(a, w, b, relu=(M % 2 == 0))
    ref = _ref(a, w, b, M % 2 == 0)
    err = (y.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 6e-6 * scale * max(1.0, (K / 256) ** 0.5), (err, scale)      # fp32-grade (plain TF32 would be ~5e-4)


def test_linear_strided_gather_add():
    g = torch.Generator().manual_seed(5)
    big = torch.randn(3000, 768, generator=g).to(DEV)
    pos = torch.randn(3000, 768, generator=g).to(DEV)
    w = (torch.randn(256, 256, generator=g) / 16).to(DEV)
    idx = torch.randint(0, 3000, (1111,), generator=g).int().to(DEV)
    out = torch.zeros(1111, 1024, device=DEV)
    a, a2 = big[:, 256:512], pos[:, 256:512]                      # column slices: lda = 768
    ops.linear(a, w, None, a_index=idx, a_add=a2, out=out[:, 512:768], M=1111, K=256)
    ref = (a[idx.long()] + a2[idx.long()]).double() @ w.double().t()
    assert (out[:, 512:768].double() - ref).abs().max().item() <= 4e-6 * ref.abs().max().item()
    assert out[:, :512].abs().max().item() == 0 and out[:, 768:].abs().max().item() == 0   # nothing written outside


def test_tc_and_ffma_agree_on_model_shapes():
    from roitr_b200 import engine
    g = torch.Generator().manual_seed(9)
    for (M, N, K) in [(20000, 192, 64), (5000, 128, 128), (312, 1024, 64), (312, 64, 256)]:
        a = torch.randn(M, K, generator=g).to(DEV)
        w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
        y1, y2 = ops.linear(a, w, None, wpack=engine.pack_linear_tc(w)), ops.linear(a, w, None)
        assert (y1 - y2).abs().max().item() <= 5e-6 * y2.abs().max().item()


@pytest.mark.parametrize("N", [16, 64, 312])
def test_geo_embedding_tensor_core_matches_oracle(N):
    from oracle import forward_ref as fr
    from roitr_b200 import engine
    from tests.helpers import weights
    sd = weights(1)
    e = "backbone.global_transformer.embedding"
    g = torch.Generator().manual_seed(N)
    pts = (torch.rand(N, 3, generator=g) * 3 - 1.5)
    ref = fr.geometric_embedding(sd, e, pts[None])[0]
    W = {k: sd[k].to(DEV) for k in sd if k.startswith(e)}
    wpack = torch.stack([engine.pack_tf32_sw128(W[e + ".proj_d.weight"]), engine.pack_tf32_sw128(W[e + ".proj_a.weight"])], 0).contiguous()
    p = pts.to(DEV)
    nn3 = ops.geo_knn(p, 3)
    b = ops.geo_embedding_tc(p, nn3, wpack, W[e + ".proj_d.bias"], W[e + ".proj_a.bias"], W[e + ".embedding.div_term"], 0.2, 15.0)
    torch.cuda.synchronize()
    off = ~torch.eye(N, dtype=torch.bool)     # the diagonal distance is rounding noise (x2 - 2xy + y2), compare off-diagonal
    assert (b.cpu() - ref)[off].abs().max().item() < 2e-5


PACKED_SHAPES = SHAPES + [(640000, 192, 64), (40000, 768, 256), (10000, 512, 512), (129, 65, 33), (4992, 1024, 64),
                          # row-group kernel (N > 128, >= 74 row tiles): odd number of weight tiles (last group holds ONE tile),
                          # N not a multiple of 32 inside the second tile, a ragged last row tile
                          (20000, 520, 96), (12000, 200, 64), (9473, 640, 32), (9472, 258, 128)]


@pytest.mark.parametrize("M,N,K", PACKED_SHAPES)
def test_linear_tc_packed_matches_fp64(M, N, K):
    from roitr_b200 import engine
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    y = ops.linear(a, w, b, relu=(M % 2 == 0), wpack=engine.pack_linear_tc(w))
    ref = _ref(a, w, b, M % 2 == 0)
    err = (y.double() - ref).abs().max().item()
    assert err <= 6e-6 * ref.abs().max().item() * max(1.0, (K / 256) ** 0.5), err


def test_linear_tc_packed_strided_gather_add():
    from roitr_b200 import engine
    g = torch.Generator().manual_seed(6)
    big = torch.randn(3000, 768, generator=g).to(DEV)
    pos = torch.randn(3000, 768, generator=g).to(DEV)
    w = (torch.randn(256, 256, generator=g) / 16).to(DEV)
    idx = torch.randint(0, 3000, (1111,), generator=g).int().to(DEV)
    out = torch.zeros(1111, 1024, device=DEV)
    a, a2 = big[:, 256:512], pos[:, 256:512]
    ops.linear(a, w, None, a_index=idx, a_add=a2, out=out[:, 512:768], M=1111, K=256, wpack=engine.pack_linear_tc(w))
    ref = (a[idx.long()] + a2[idx.long()]).double() @ w.double().t()
    assert (out[:, 512:768].double() - ref).abs().max().item() <= 6e-6 * ref.abs().max().item()
    assert out[:, :512].abs().max().item() == 0 and out[:, 768:].abs().max().item() == 0


@pytest.mark.parametrize("M,R,N,K", [(160000, 640000, 128, 64), (40000, 160000, 512, 128), (9984, 40000, 1024, 256), (333, 1000, 64, 64),
                                     (5000, 5000, 192, 96)])
def test_linear_tc_packed_gathered_rows(M, R, N, K):
    """Rows gathered through a_index alone take the GATHER instantiation of the streaming kernel (the down-sampling [f|q]
    layers): strided source, repeated and out-of-order indices, bias + ReLU, against fp64."""
    from roitr_b200 import engine
    g = torch.Generator().manual_seed(M + N + K)
    big = torch.randn(R, K + 64, generator=g).to(DEV)
    a = big[:, 32:32 + K]                                   # column slice of a wider buffer: lda = K + 64, 16-byte aligned
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    idx = torch.randint(0, R, (M,), generator=g).int().to(DEV)
    y = ops.linear(a, w, b, relu=True, a_index=idx, wpack=engine.pack_linear_tc(w))
    ref = torch.relu(a[idx.long()].double() @ w.double().t() + b.double())
    assert (y.double() - ref).abs().max().item() <= 6e-6 * ref.abs().max().item() * max(1.0, (K / 256) ** 0.5)
    # the same function as the plain path on pre-gathered rows, bit for bit
    y2 = ops.linear(a[idx.long()].contiguous(), w, b, relu=True, wpack=engine.pack_linear_tc(w))
    assert torch.equal(y, y2)


@pytest.mark.parametrize("N,scale", [(16, 3.0), (64, 3.0), (312, 3.0), (97, 40.0), (64, 400.0)])
def test_geo_embedding_table_matches_oracle_and_tensor_core(N, scale):
    """Tabulated embedding (csrc/geo_table.cu): shared-memory tables (scale 3 m), the global-table path (distances up to
    ~70 m at scale 40) and the direct evaluation beyond every table (scale 400), against the oracle and the GEMM kernel."""
    from oracle import forward_ref as fr
    from roitr_b200 import engine
    from tests.helpers import weights
    sd = weights(1)
    e = "backbone.global_transformer.embedding"
    g = torch.Generator().manual_seed(N)
    B = 3
    pts = (torch.rand(B * N, 3, generator=g) - 0.5) * scale
    W = {k: sd[k].to(DEV) for k in sd if k.startswith(e)}
    tables = engine.build_geo_tables(W[e + ".proj_d.weight"], W[e + ".proj_d.bias"], W[e + ".proj_a.weight"],
                                     W[e + ".proj_a.bias"], W[e + ".embedding.div_term"])
    assert tables["bound"] <= engine.GEO_TABLE_TOL
    wpack = torch.stack([engine.pack_tf32_sw128(W[e + ".proj_d.weight"]), engine.pack_tf32_sw128(W[e + ".proj_a.weight"])], 0).contiguous()
    p = pts.to(DEV)
    nn3 = ops.geo_knn_batched(B, N, p, 3)
    a = ops.geo_embedding_table(B, N, p, nn3, tables, W[e + ".proj_d.weight"], W[e + ".proj_d.bias"], W[e + ".proj_a.weight"],
                                W[e + ".proj_a.bias"], W[e + ".embedding.div_term"], 0.2, 15.0)
    b = ops.geo_embedding_tc_batched(B, N, p, nn3, wpack, W[e + ".proj_d.bias"], W[e + ".proj_a.bias"],
                                     W[e + ".embedding.div_term"], 0.2, 15.0)
    torch.cuda.synchronize()
    off = ~torch.eye(N, dtype=torch.bool)
    # the fp32 evaluations (oracle, GEMM kernel) round the sinusoid argument t * div_term: their own error grows with t
    tol = 2e-5 * max(1.0, scale / 3.0)
    for c in range(B):
        ref = fr.geometric_embedding(sd, e, pts[None, c * N:(c + 1) * N])[0]
        assert (a[c].cpu() - ref)[off].abs().max().item() < tol
        assert (a[c] - b[c]).abs().max().item() < tol


@pytest.mark.parametrize("N,M,C", [(312, 312, 256), (16, 16, 256), (15, 15, 256), (125, 125, 512), (40, 57, 256)])
def test_attention_matches_fp64(N, M, C):
    """Q K^T / P V on tcgen05 + the streaming E pass (csrc/geo_attn2.cu: the barrier-free kernel at C = 256, the
    chunk-synchronous one at C = 512) against a float64 evaluation of geoattention.py:43-66,101-136."""
    g = torch.Generator().manual_seed(N * 1000 + M)
    B, H = 3, 4
    c = C // H
    qkv = torch.randn(B * N, 3 * C, generator=g).to(DEV)
    kv = torch.randn(B * M, 2 * C, generator=g).to(DEV)
    # cross attention (no E): q from one set of clouds, k / v from another
    q, k, v = qkv[:, :C], kv[:, :C], kv[:, C:]
    h_new = ops.attention_tc(B, N, M, C, q, k, v)
    q64, k64, v64 = (t.double().view(B, -1, H, c).permute(0, 2, 1, 3) for t in (q, k, v))
    ref = (torch.softmax(q64 @ k64.transpose(-1, -2) / c ** 0.5, -1) @ v64).permute(0, 2, 1, 3).reshape(B * N, C)
    assert (h_new.double() - ref).abs().max().item() < 2e-5
    if N != M:
        return
    # RPE self attention
    E = torch.randn(B, N, N, C, generator=g).to(DEV) * 0.5
    gq = torch.randn(B * N, H, C, generator=g).to(DEV) * 0.1
    bp = torch.randn(C, generator=g).to(DEV) * 0.1
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    h_new, G_new = ops.attention_tc(B, N, N, C, q, k, v, E=E, gq=gq, bp=bp)
    q64, k64, v64 = (t.double().view(B, N, H, c).permute(0, 2, 1, 3) for t in (q, k, v))
    sp = torch.einsum("bnhc,bnmc->bhnm", gq.double().view(B, N, H, C), E.double())
    qb = (q.double().view(B, N, H, c) * bp.double().view(1, 1, H, c)).sum(-1).permute(0, 2, 1)[..., None]
    S = (q64 @ k64.transpose(-1, -2) + sp + qb) / c ** 0.5
    ref_h = (torch.softmax(S, -1) @ v64).permute(0, 2, 1, 3).reshape(B * N, C)
    Sm = S.masked_fill(torch.eye(N, dtype=torch.bool, device=DEV)[None, None], float("-inf"))
    ref_G = torch.einsum("bhnm,bnmc->bnhc", torch.softmax(Sm, -1), E.double()).reshape(B * N, H, C)
    assert (h_new.double() - ref_h).abs().max().item() < 3e-5
    assert (G_new.double() - ref_G).abs().max().item() < 3e-5



@pytest.mark.parametrize("M,N,K,pre,gather,post,relu", [(20000, 64, 64, True, False, False, False), (5000, 128, 128, True, True, False, False),
                                                       (4100, 64, 64, False, False, True, True), (1250, 128, 256, False, False, False, True),
                                                       (777, 64, 128, True, True, True, True), (130, 128, 64, False, False, True, False),
                                                       # two weight tiles per row (N > 128): both halves of the row in TMEM, one LayerNorm
                                                       (9984, 256, 256, True, False, False, False), (40000, 256, 512, False, False, True, True),
                                                       (4992, 256, 1024, True, True, False, False), (1000, 192, 64, True, False, True, False),
                                                       (300, 160, 96, False, False, False, True), (60000, 32, 64, True, False, False, False),
                                                       (333, 96, 64, False, False, True, False)])
def test_linear_ln_fused_matches_fp64(M, N, K, pre, gather, post, relu):
    """roitr_linear_ln_tc_packed (LayerNorm / residuals / ReLU in the dense layer's epilogue) against fp64."""
    from roitr_b200 import engine
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    gamma, beta = (1 + 0.1 * torch.randn(N, generator=g)).to(DEV), (0.1 * torch.randn(N, generator=g)).to(DEV)
    R = 3 * M if gather else M
    res_pre = torch.randn(R, N, generator=g).to(DEV) if pre else None
    idx = torch.randint(0, R, (M,), generator=g).int().to(DEV) if gather else None
    res_post = torch.randn(M, N, generator=g).to(DEV) if post else None
    y = ops.linear_ln(a, w, b, engine.pack_linear_tc(w), gamma, beta, res_pre=res_pre, res_pre_index=idx, res_post=res_post, relu=relu)
    t = a.double() @ w.double().t() + b.double()
    if pre:
        t = t + (res_pre[idx.long()] if gather else res_pre).double()
    ref = torch.nn.functional.layer_norm(t, (N,), gamma.double(), beta.double(), 1e-5)
    if post:
        ref = ref + res_post.double()
    if relu:
        ref = torch.relu(ref)
    assert (y.double() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
    # and it is the same function as the two-kernel path
    t32 = ops.linear(a, w, b, wpack=engine.pack_linear_tc(w))
    two = ops.row_epilogue(t32, res_pre=res_pre, res_pre_index=idx, gamma=gamma, beta=beta, res_post=res_post,
                           mode=ops.MODE_LN | (ops.MODE_RELU if relu else 0))
    assert (y - two).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())


def test_geo_embedding_falls_back_to_the_gemm_when_the_table_bound_fails():
    """engine.pack_weights: weights whose 4th-derivative bound defeats the table (here proj_d scaled x 1e6) take the tcgen05
    GEMM (csrc/geo_tc.cu) without any flag; the forward still agrees with the oracle on the embedding."""
    from oracle import forward_ref as fr
    from roitr_b200 import engine
    from tests.helpers import CONFIG_3D, weights
    sd = dict(weights(1))
    e = "backbone.global_transformer.embedding"
    sd[e + ".proj_d.weight"] = sd[e + ".proj_d.weight"] * 1.0e6
    W = engine.pack_weights(sd, torch.device(DEV), CONFIG_3D["transformer_architecture"])
    assert (e + "#tables") not in W and (e + "#wpack") in W
    W_ok = engine.pack_weights(weights(1), torch.device(DEV), CONFIG_3D["transformer_architecture"])
    assert (e + "#tables") in W_ok
    g = torch.Generator().manual_seed(3)
    N, B = 96, 1
    pts = (torch.rand(2 * B * N, 3, generator=g) * 3 - 1.5)
    E_all, _ = engine.geometric_embedding_batch(W, B, pts.to(DEV), B * N)
    ref = fr.geometric_embedding(sd, e, pts[None, :N])[0]
    off = ~torch.eye(N, dtype=torch.bool)
    # fp32-grade against the scale of the (x 1e6) embedding: entries are sums of terms of that size
    assert (E_all[0].cpu() - ref)[off].abs().max().item() < 2e-5 * ref[off].abs().max().item()


def test_geo_embedding_table_propagates_nan_instead_of_reading_out_of_bounds():
    """A NaN / Inf superpoint coordinate must give NaN rows like the reference, not an out-of-table read."""
    from roitr_b200 import engine
    from tests.helpers import weights
    sd = weights(1)
    e = "backbone.global_transformer.embedding"
    W = {k: sd[k].to(DEV) for k in sd if k.startswith(e)}
    tables = engine.build_geo_tables(W[e + ".proj_d.weight"], W[e + ".proj_d.bias"], W[e + ".proj_a.weight"],
                                     W[e + ".proj_a.bias"], W[e + ".embedding.div_term"])
    N = 40
    pts = torch.rand(N, 3, generator=torch.Generator().manual_seed(1)).to(DEV)
    pts[7, 1] = float("nan")
    pts[11, 0] = float("inf")
    nn3 = ops.geo_knn_batched(1, N, pts, 3)
    E = ops.geo_embedding_table(1, N, pts, nn3.clamp(0, N - 1), tables, W[e + ".proj_d.weight"], W[e + ".proj_d.bias"],
                                W[e + ".proj_a.weight"], W[e + ".proj_a.bias"], W[e + ".embedding.div_term"], 0.2, 15.0)
    torch.cuda.synchronize()
    assert torch.isnan(E[0, 7, 3]).any() and torch.isnan(E[0, 3, 7]).any()
    clean = [i for i in range(N) if i not in (7, 11) and 7 not in nn3[i].tolist() and 11 not in nn3[i].tolist()]
    assert torch.isfinite(E[0][clean][:, clean]).all()


@pytest.mark.parametrize("m,n,C,K,gather,ordered", [(1000, 1000, 64, 8, False, False), (1001, 4000, 64, 16, True, False),
                                                    (777, 777, 128, 16, False, True), (500, 2000, 128, 8, True, False),
                                                    (313, 313, 256, 16, False, False), (125, 500, 512, 16, True, False), (1, 40, 64, 8, True, False)])
def test_local_attention_matches_fp64(m, n, C, K, gather, ordered):
    """csrc/local_attn.cu against a float64 evaluation of the folded form
    of attention.py:166-200: S = (q.k_j + (Ap^T q).ppf_j + q.cp) / sqrt(c), A = softmax_j S, out = sum_j A_j v_j + Avp (sum_j A_j ppf_j) + cvp."""
    g = torch.Generator().manual_seed(m + C + K)
    H, c = 4, C // 4
    qkv = torch.randn(n, 3 * C, generator=g).to(DEV)
    node_idx = torch.randint(0, n, (m,), generator=g).int().to(DEV) if gather else None
    group = torch.randint(0, n, (m, K), generator=g).int().to(DEV)
    ppf = torch.rand(m, K, 4, generator=g).to(DEV)
    Ap, Avp = (torch.randn(C, 4, generator=g) * 0.3).to(DEV), (torch.randn(C, 4, generator=g) * 0.3).to(DEV)
    cp, cvp = (torch.randn(C, generator=g) * 0.1).to(DEV), (torch.randn(C, generator=g) * 0.1).to(DEV)
    order = None
    if ordered:      # the queries' own grid as visiting order (any permutation must give the same rows)
        pts = torch.rand(m, 3, generator=g).to(DEV)
        off = torch.tensor([m], dtype=torch.int32, device=DEV)
        order = (ops.knn_grid_build(pts, off), 1)
    out = ops.local_attention(qkv, C, node_idx, group, ppf, Ap, cp, Avp, cvp, order=order)
    q = qkv[:, :C].double()[node_idx.long() if gather else torch.arange(m, device=DEV)].view(m, H, c)
    k = qkv[:, C:2 * C].double()[group.long()].view(m, K, H, c)
    v = qkv[:, 2 * C:].double()[group.long()].view(m, K, H, c)
    qa = torch.einsum("mhc,hcf->mhf", q, Ap.double().view(H, c, 4))
    qb = (q * cp.double().view(1, H, c)).sum(-1)
    S = (torch.einsum("mhc,mkhc->mhk", q, k) + torch.einsum("mhf,mkf->mhk", qa, ppf.double()) + qb[..., None]) / c ** 0.5
    A = torch.softmax(S, -1)
    w = torch.einsum("mhk,mkf->mhf", A, ppf.double())
    ref = torch.einsum("mhk,mkhc->mhc", A, v) + torch.einsum("hcf,mhf->mhc", Avp.double().view(H, c, 4), w) + cvp.double().view(1, H, c)
    assert (out.double() - ref.reshape(m, C)).abs().max().item() < 2e-5
