"""oracle/pointops_ref.c against independent brute-force numpy, including the reference's edge semantics:
unfilled slots (n < nsample) stay (idx=start, d2=1e10) (knnquery_cuda_kernel.cu:88-91), several segments,
first FPS sample = first point of the segment (sampling_cuda_kernel.cu:39), golden kNN/FPS traces."""
import numpy as np
import torch

from oracle import native
from tests.helpers import knn_equal_up_to_ties, load_golden
from roitr_b200.synthetic import synthetic_pair


def _d2(q, r):
    dx, dy, dz = (q[:, None, i].astype(np.float32) - r[None, :, i].astype(np.float32) for i in range(3))
    # same association as the kernels: fma(dz,dz,fma(dx,dx,dy*dy)), emulated in float64 then rounded stepwise
    t = (dy * dy).astype(np.float32)
    t = (dx.astype(np.float64) * dx + t).astype(np.float32)
    return (dz.astype(np.float64) * dz + t).astype(np.float32)


def test_knn_matches_bruteforce_multi_segment():
    g = torch.Generator().manual_seed(1)
    xyz = torch.rand(300, 3, generator=g)
    new = torch.rand(90, 3, generator=g)
    off = torch.tensor([100, 300], dtype=torch.int32)
    noff = torch.tensor([40, 90], dtype=torch.int32)
    idx, d2 = native.knn(5, xyz, new, off, noff)
    full = _d2(new.numpy(), xyz.numpy())
    for q in range(90):
        s, e = (0, 100) if q < 40 else (100, 300)
        order = np.argsort(full[q, s:e], kind="stable")[:5] + s
        assert np.array_equal(np.sort(full[q, order]), d2[q].numpy())
        assert knn_equal_up_to_ties(idx[q:q + 1].numpy(), d2[q:q + 1].numpy(), order[None], full[q, order][None])


def test_knn_unfilled_slots():
    xyz = torch.rand(4, 3)
    idx, d2 = native.knn(6, xyz, xyz, torch.tensor([4], dtype=torch.int32), torch.tensor([4], dtype=torch.int32))
    assert (d2[:, 4:] == 1e10).all() and (idx[:, 4:] == 0).all()
    assert (idx[:, 0] == torch.arange(4)).all() and (d2[:, 0] == 0).all()


def test_fps_first_point_and_bruteforce():
    g = torch.Generator().manual_seed(2)
    xyz = torch.rand(200, 3, generator=g)
    off = torch.tensor([80, 200], dtype=torch.int32)
    noff = torch.tensor([20, 50], dtype=torch.int32)
    idx = native.fps(xyz, off, noff).numpy()
    assert idx[0] == 0 and idx[20] == 80
    x = xyz.numpy()
    for (s, e, ms, me) in ((0, 80, 0, 20), (80, 200, 20, 50)):
        tmp = np.full(e - s, 1e10, np.float32)
        last = s
        for j in range(ms + 1, me):
            tmp = np.minimum(tmp, _d2(x[s:e], x[last:last + 1])[:, 0])
            assert tmp[idx[j] - s] == tmp.max()
            last = idx[j]


def test_native_reproduces_golden_traces():
    z, meta = load_golden("golden_3dmatch_n1024")
    pair = synthetic_pair(meta["pair_index"], meta["n"])
    o = torch.tensor([1024], dtype=torch.int32)
    idx, d2 = native.knn(9, pair["src_raw_pcd"], pair["src_raw_pcd"], o, o)
    assert np.array_equal(idx.numpy(), z["knn_0_idx"])
    assert np.array_equal(torch.sqrt(d2).numpy(), z["knn_0_dist"])
    f = native.fps(pair["src_raw_pcd"], o, torch.tensor([256], dtype=torch.int32))
    assert np.array_equal(f.numpy(), z["fps_0"])


def test_reference_block_size_rule_is_integer_log2():
    # the CUDA FPS kernel derives the reference's block size (src/cuda_utils.h:11-14) with integer arithmetic
    for n in list(range(1, 5000)) + [2 ** p + d for p in range(12, 17) for d in (-1, 0, 1)]:
        assert native.fps_block_size(n) == min(1 << (n.bit_length() - 1), 1024), n
