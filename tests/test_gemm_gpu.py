"""roitr_linear_tc (tcgen05, 3xTF32 split precision) and roitr_linear (fp32 FFMA) against an fp64 reference."""
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


@pytest.mark.parametrize("tc", [True, False])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_linear_matches_fp64(M, N, K, tc):
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    y = ops.linear(a, w, b, relu=(M % 2 == 0), tc=tc)
    ref = _ref(a, w, b, M % 2 == 0)
    err = (y.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 6e-6 * scale * max(1.0, (K / 256) ** 0.5), (err, scale)      # fp32-grade (plain TF32 would be ~5e-4)


@pytest.mark.parametrize("tc", [True, False])
def test_linear_strided_gather_add(tc):
    g = torch.Generator().manual_seed(5)
    big = torch.randn(3000, 768, generator=g).to(DEV)
    pos = torch.randn(3000, 768, generator=g).to(DEV)
    w = (torch.randn(256, 256, generator=g) / 16).to(DEV)
    idx = torch.randint(0, 3000, (1111,), generator=g).int().to(DEV)
    out = torch.zeros(1111, 1024, device=DEV)
    a, a2 = big[:, 256:512], pos[:, 256:512]                      # column slices: lda = 768
    ops.linear(a, w, None, a_index=idx, a_add=a2, out=out[:, 512:768], M=1111, K=256, tc=tc)
    ref = (a[idx.long()] + a2[idx.long()]).double() @ w.double().t()
    assert (out[:, 512:768].double() - ref).abs().max().item() <= 4e-6 * ref.abs().max().item()
    assert out[:, :512].abs().max().item() == 0 and out[:, 768:].abs().max().item() == 0   # nothing written outside


def test_tc_and_ffma_agree_on_model_shapes():
    g = torch.Generator().manual_seed(9)
    for (M, N, K) in [(20000, 192, 64), (5000, 128, 128), (312, 1024, 64), (312, 64, 256)]:
        a = torch.randn(M, K, generator=g).to(DEV)
        w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
        y1, y2 = ops.linear(a, w, None, tc=True), ops.linear(a, w, None, tc=False)
        assert (y1 - y2).abs().max().item() <= 5e-6 * y2.abs().max().item()


@pytest.mark.parametrize("N", [16, 64, 312])
def test_geo_embedding_tensor_core_matches_ffma_and_oracle(N):
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
    a = ops.geo_embedding(p, nn3, W[e + ".proj_d.weight"], W[e + ".proj_d.bias"], W[e + ".proj_a.weight"], W[e + ".proj_a.bias"],
                          W[e + ".embedding.div_term"], 0.2, 15.0)
    b = ops.geo_embedding_tc(p, nn3, wpack, W[e + ".proj_d.bias"], W[e + ".proj_a.bias"], W[e + ".embedding.div_term"], 0.2, 15.0)
    torch.cuda.synchronize()
    off = ~torch.eye(N, dtype=torch.bool)     # the diagonal distance is rounding noise (x2 - 2xy + y2), compare off-diagonal
    assert (a.cpu() - ref)[off].abs().max().item() < 2e-5
    assert (b.cpu() - ref)[off].abs().max().item() < 2e-5
    assert (a - b).abs().max().item() < 2e-5


PACKED_SHAPES = SHAPES + [(640000, 192, 64), (40000, 768, 256), (10000, 512, 512), (129, 65, 33), (4992, 1024, 64)]


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
