"""GPU parity of the native point ops (through the C ABI) against the oracle (oracle/pointops_ref.c,
oracle/forward_ref.py) — bit-exact for indices, 1e-6 for PPF."""
import numpy as np
import pytest
import torch

from oracle import forward_ref as fr
from oracle import native
from roitr_b200 import pointops, pointops_cuda
from roitr_b200.synthetic import synthetic_pair
from tests.helpers import knn_equal_up_to_ties

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _i32(v):
    return torch.tensor(v, dtype=torch.int32)


def _cloud(n, seed):
    g = torch.Generator().manual_seed(seed)
    p = torch.rand(n, 3, generator=g) * 2 - 1
    nr = torch.randn(n, 3, generator=g)
    return p.contiguous(), (nr / nr.norm(dim=1, keepdim=True)).contiguous()


def _boundary_ok(idx_g, d2_g, idx_o, d2_o, xyz, q):
    """Bit-exact except permutations among exactly equal distances, including ties at the k-th boundary."""
    if knn_equal_up_to_ties(idx_g, d2_g, idx_o, d2_o):
        return True
    if not np.array_equal(d2_g, d2_o):
        return False
    bad_rows = np.unique(np.nonzero(idx_g != idx_o)[0])
    for r in bad_rows:
        for c in np.nonzero(idx_g[r] != idx_o[r])[0]:
            for ix in (idx_g[r, c], idx_o[r, c]):   # both candidates must really sit at that distance
                d = q[r].astype(np.float32) - xyz[ix].astype(np.float32)
                t = np.float32(d[1] * d[1])
                t = np.float32(np.float64(d[0]) * d[0] + t)
                t = np.float32(np.float64(d[2]) * d[2] + t)
                if t != d2_o[r, c]:
                    return False
    return True


@pytest.mark.parametrize("n,m,ns,segs", [
    (1024, 1024, 9, None), (5000, 1250, 17, None), (300, 90, 5, ([100, 300], [40, 90])),
    (37, 37, 17, None), (10, 10, 17, None), (4099, 33, 1, None), (20000, 20000, 9, None),
    (2600, 700, 17, ([1300, 2600], [350, 700])), (6, 3, 3, ([2, 6], [1, 3])),
])
def test_knnquery_bit_exact(n, m, ns, segs):
    xyz, _ = _cloud(n, 7)
    q = xyz[:m].clone() if m <= n else _cloud(m, 8)[0]
    if n == m:
        q = xyz.clone()
    off, noff = (_i32([n]), _i32([m])) if segs is None else (_i32(segs[0]), _i32(segs[1]))
    idx_o, d2_o = native.knn(ns, xyz, q, off, noff)
    idx_g, dist_g = pointops.knnquery(ns, xyz.to(DEV), q.to(DEV), off.to(DEV), noff.to(DEV))
    torch.cuda.synchronize()
    # returned distance = IEEE sqrt of the squared distance (== torch.sqrt on CUDA, what the reference runs;
    # torch's CPU sqrt is not correctly rounded: ~0.5% of entries differ by 1 ulp, measured on the B200 box)
    np.testing.assert_allclose(dist_g.cpu().numpy(), torch.sqrt(d2_o).numpy(), rtol=2e-7, atol=0)
    # drop-in module form: caller-allocated outputs, squared distances
    idx2 = torch.zeros(m, ns, dtype=torch.int32, device=DEV)
    d2 = torch.zeros(m, ns, dtype=torch.float32, device=DEV)
    pointops_cuda.knnquery_cuda(m, ns, xyz.to(DEV), q.to(DEV), off.to(DEV), noff.to(DEV), idx2, d2)
    assert np.array_equal(d2.cpu().numpy(), d2_o.numpy())
    assert torch.equal(idx2, idx_g)
    assert torch.equal(idx_g.cpu(), idx_o)   # bit-exact, exact-distance ties included (knn_tie_fixup_kernel)


def test_knn_duplicates_and_unfilled():
    xyz = torch.tensor([[0., 0, 0], [0, 0, 0], [1, 0, 0], [0, 0, 0]])
    off = _i32([4])
    idx_o, d2_o = native.knn(6, xyz, xyz, off, off)
    idx_g, dist_g = pointops.knnquery(6, xyz.to(DEV), xyz.to(DEV), off.to(DEV), off.to(DEV))
    np.testing.assert_allclose(dist_g.cpu().numpy(), torch.sqrt(d2_o).numpy(), rtol=2e-7, atol=0)
    assert (idx_g[:, 4:] == 0).all()          # unfilled slots keep the segment start (knnquery_cuda_kernel.cu:88-91)
    assert torch.equal(idx_g.cpu(), idx_o)    # exact ties: the reference's heap order, replayed exactly


@pytest.mark.parametrize("n,m,k", [(1024, 1024, 8), (5000, 1250, 16), (20000, 5000, 16), (16, 16, 16)])
def test_knn_ppf_fused(n, m, k):
    xyz, nrm = _cloud(n, 3)
    sel = torch.randperm(n, generator=torch.Generator().manual_seed(1))[:m].sort()[0] if m < n else torch.arange(n)
    q, qn = xyz[sel].contiguous(), nrm[sel].contiguous()
    off, noff = _i32([n]), _i32([m])
    g = fr.group_indices(k, xyz, q, off, noff)
    ppf_o = fr.ppf(q, qn, xyz[g], nrm[g])
    idx_g, ppf_g = pointops.knn_ppf(k, xyz.to(DEV), nrm.to(DEV), q.to(DEV), qn.to(DEV), off.to(DEV), noff.to(DEV))
    same = (idx_g.cpu().long() == g)
    assert bool(same.all())
    np.testing.assert_allclose(ppf_g.cpu().numpy()[same.numpy()], ppf_o.numpy()[same.numpy()], rtol=0, atol=2e-6)
    qg = pointops.queryandgroup(k, xyz.to(DEV), q.to(DEV), xyz.to(DEV), None, off.to(DEV), noff.to(DEV), return_idx=True)
    assert qg.dtype == torch.int64 and torch.equal(qg.int(), idx_g)


@pytest.mark.parametrize("n,cluster", [(16, 0), (312, 1), (1250, 0), (1250, 2), (5000, 0), (5000, 1), (5000, 8),
                                       (20000, 0), (20000, 4), (30000, 0), (1024, 0), (8191, 2)])
def test_fps_bit_exact(n, cluster):
    xyz, _ = _cloud(n, 11)
    off, noff = _i32([n]), _i32([n // 4])
    ref = native.fps(xyz, off, noff)
    got = pointops.furthestsampling(xyz.to(DEV), off.to(DEV), noff.to(DEV), cluster=cluster)
    torch.cuda.synchronize()
    assert torch.equal(got.cpu(), ref)


def test_fps_batched_segments_and_dropin():
    xyz, _ = _cloud(9000, 5)
    off, noff = _i32([5000, 6000, 9000]), _i32([1250, 1500, 2250])
    ref = native.fps(xyz, off, noff)   # reference semantics: block size from the batch-wide n_max
    idx = torch.zeros(2250, dtype=torch.int32, device=DEV)
    tmp = torch.full((9000,), 1e10, device=DEV)
    pointops_cuda.furthestsampling_cuda(3, 5000, xyz.to(DEV), off.to(DEV), noff.to(DEV), tmp, idx)
    assert torch.equal(idx.cpu(), ref)
    got, nx = pointops.furthestsampling(xyz.to(DEV), off.to(DEV), noff.to(DEV), return_xyz=True)
    assert torch.equal(got.cpu(), ref) and torch.equal(nx.cpu(), xyz[ref.long()])


def test_knn_lattice_many_exact_ties():
    # a lattice makes almost every query tie-affected: exercises the exact heap replay path at scale
    g = torch.stack(torch.meshgrid(*[torch.arange(13.)] * 3, indexing="ij"), -1).reshape(-1, 3).contiguous() * 0.1
    off = _i32([g.shape[0]])
    idx_o, d2_o = native.knn(17, g, g, off, off)
    idx_g, dist_g = pointops.knnquery(17, g.to(DEV), g.to(DEV), off.to(DEV), off.to(DEV))
    assert torch.equal(idx_g.cpu(), idx_o)
    np.testing.assert_allclose(dist_g.cpu().numpy(), torch.sqrt(d2_o).numpy(), rtol=2e-7, atol=0)


def test_fps_exact_ties_follow_reference_tree():
    # a regular lattice produces many exactly equal maxima: the reference's winner depends on its block tree
    g = torch.stack(torch.meshgrid(*[torch.arange(12.)] * 3, indexing="ij"), -1).reshape(-1, 3).contiguous()
    off, noff = _i32([g.shape[0]]), _i32([g.shape[0] // 4])
    ref = native.fps(g, off, noff)
    for cl in (1, 2, 4, 8):
        got = pointops.furthestsampling(g.to(DEV), off.to(DEV), noff.to(DEV), cluster=cl)
        assert torch.equal(got.cpu(), ref), cl


def test_interpolation_matches_oracle():
    pair = synthetic_pair(0, 4096)
    fine = pair["tgt_pcd"]
    oc, of = _i32([1024]), _i32([4096])
    coarse = fine[native.fps(fine, of, oc).long()].contiguous()
    feat = torch.randn(1024, 128, generator=torch.Generator().manual_seed(0))
    base = torch.randn(4096, 128, generator=torch.Generator().manual_seed(1))
    ref = base + fr.interpolate3(coarse, fine, feat, oc, of)
    got = pointops.interpolation(coarse.to(DEV), fine.to(DEV), feat.to(DEV), oc.to(DEV), of.to(DEV), base=base.to(DEV))
    np.testing.assert_allclose(got.cpu().numpy(), ref.numpy(), rtol=0, atol=1e-6)


@pytest.mark.parametrize("n,m,k,drop,segs", [
    (20000, 20000, 8, 1, None), (20000, 5000, 16, 1, None), (5000, 20000, 3, 0, None), (4096, 4096, 16, 1, None),
    (9000, 2250, 16, 1, ([5000, 6000, 9000], [1250, 1500, 2250])), (3000, 3000, 31, 1, None), (2500, 40, 1, 0, None),
])
def test_grid_knn_equals_brute_force(n, m, k, drop, segs):
    """The grid-accelerated kNN must return exactly what the brute-force kernel returns (indices, distances, PPF)."""
    from roitr_b200 import ops
    from roitr_b200.synthetic import synthetic_pair
    pair = synthetic_pair(3, max(n, m))
    xyz, nrm = pair["tgt_pcd"][:n].contiguous().to(DEV), pair["tgt_normals"][:n].contiguous().to(DEV)
    if m == n:
        q, qn = xyz, nrm
    elif m < n:
        sel = torch.arange(0, n, n // m)[:m]
        q, qn = xyz[sel].contiguous(), nrm[sel].contiguous()
    else:   # queries outside / around the reference set's bounding box too
        q = (pair["src_pcd"][:m] * 1.3).contiguous().to(DEV)
        qn = pair["src_normals"][:m].contiguous().to(DEV)
    off, noff = (_i32([n]), _i32([m])) if segs is None else (_i32(segs[0]), _i32(segs[1]))
    off, noff = off.to(DEV), noff.to(DEV)
    want_ppf = drop == 1
    a = ops.knn_ppf(k, xyz, nrm, q, qn, off, noff, drop_first=drop, want_ppf=want_ppf, want_dist=True)
    grid = ops.knn_grid_build(xyz, off)
    b = ops.knn_ppf(k, xyz, nrm, q, qn, off, noff, drop_first=drop, want_ppf=want_ppf, want_dist=True, grid=grid)
    torch.cuda.synchronize()
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2])
    if want_ppf:
        assert torch.equal(a[1], b[1])


def test_grid_knn_lattice_ties_and_degenerate_clouds():
    from roitr_b200 import ops
    g = (torch.stack(torch.meshgrid(*[torch.arange(14.)] * 3, indexing="ij"), -1).reshape(-1, 3) * 0.1).contiguous().to(DEV)
    off = _i32([g.shape[0]]).to(DEV)
    idx_o, d2_o = native.knn(17, g.cpu(), g.cpu(), off.cpu(), off.cpu())
    grid = ops.knn_grid_build(g, off)
    idx, _, dist = ops.knn_ppf(17, g, None, g, None, off, off, drop_first=0, want_ppf=False, want_dist=True, grid=grid)
    assert torch.equal(idx.cpu(), idx_o)
    flat = torch.zeros(3000, 3); flat[:, 0] = torch.linspace(0, 1, 3000)      # all points on a line: ny = nz = 1
    flat = flat.contiguous().to(DEV)
    off = _i32([3000]).to(DEV)
    a = ops.knn_ppf(5, flat, None, flat, None, off, off, drop_first=0, want_ppf=False, want_dist=True)
    b = ops.knn_ppf(5, flat, None, flat, None, off, off, drop_first=0, want_ppf=False, want_dist=True, grid=ops.knn_grid_build(flat, off))
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2])


def test_dropin_is_called_the_way_the_reference_op_layer_calls_it():
    """INTEGRATION.md level 1: ``sys.modules['pointops_cuda'] = roitr_b200.pointops_cuda`` under the UNMODIFIED reference
    op layer. /root/reference does not exist on the GPU box, so the two call sequences the forward path uses are restated
    here statement by statement from cpp_wrappers/pointops/functions/pointops.py (FurthestSampling.forward :17-25,
    KNNQuery.forward :37-43): legacy ``torch.cuda.IntTensor`` / ``FloatTensor`` constructors for the outputs, ``n_max``
    arriving as a 0-d CUDA tensor produced by ``max()`` over offset differences, ``m`` / ``nsample`` as Python ints."""
    import sys
    import types
    saved = sys.modules.get("pointops_cuda")
    sys.modules["pointops_cuda"] = pointops_cuda
    try:
        import pointops_cuda as ext                                  # what `import pointops_cuda` (pointops.py:7) resolves to
        assert isinstance(ext, types.ModuleType) and ext is pointops_cuda
        xyz_h, _ = _cloud(9000, 6)
        xyz = xyz_h.to(DEV)
        offset, new_offset = _i32([5000, 6000, 9000]).to(DEV), _i32([1250, 1500, 2250]).to(DEV)
        # ---- FurthestSampling.forward, pointops.py:17-25 ----
        assert xyz.is_contiguous()
        n, b, n_max = xyz.shape[0], offset.shape[0], offset[0]
        for i in range(1, b):
            n_max = max(offset[i] - offset[i - 1], n_max)
        assert torch.is_tensor(n_max) and n_max.dim() == 0 and n_max.is_cuda        # the 0-d tensor the binding must accept
        idx = torch.cuda.IntTensor(new_offset[b - 1].item()).zero_()
        tmp = torch.cuda.FloatTensor(n).fill_(1e10)
        ext.furthestsampling_cuda(b, n_max, xyz, offset, new_offset, tmp, idx)
        del tmp
        assert torch.equal(idx.cpu(), native.fps(xyz_h, offset.cpu(), new_offset.cpu()))
        # ---- KNNQuery.forward, pointops.py:37-43 (new_xyz = the sampled points) ----
        nsample = 17
        new_xyz = xyz[idx.long()].contiguous()
        assert xyz.is_contiguous() and new_xyz.is_contiguous()
        m = new_xyz.shape[0]
        kidx = torch.cuda.IntTensor(m, nsample).zero_()
        dist2 = torch.cuda.FloatTensor(m, nsample).zero_()
        ext.knnquery_cuda(m, nsample, xyz, new_xyz, offset, new_offset, kidx, dist2)
        idx_o, d2_o = native.knn(nsample, xyz_h, new_xyz.cpu(), offset.cpu(), new_offset.cpu())
        assert torch.equal(kidx.cpu(), idx_o) and torch.equal(dist2.cpu(), d2_o)
        assert kidx.dtype == torch.int32 and torch.sqrt(dist2).dtype == torch.float32   # what :43 returns
        # the out-of-path entry points exist (pointops_api.cpp:15-22) and say so when called
        with pytest.raises(NotImplementedError):
            ext.grouping_forward_cuda(1, 1, 1, 1, None, None, None)
    finally:
        if saved is None:
            del sys.modules["pointops_cuda"]
        else:
            sys.modules["pointops_cuda"] = saved
