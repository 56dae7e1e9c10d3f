"""The reference's OWN kernels (oracle/_ref: knnquery_cuda_kernel.cu / sampling_cuda_kernel.cu compiled unmodified for
sm_100a) against the C oracle and against libroitr_b200, on the GPU, at the benchmark size. This pins the oracle's
arithmetic (FMA association, tie order) to the reference binary rather than to a reading of its source."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import native
from roitr_b200 import pointops
from tests.test_pointops_gpu import _boundary_ok, _cloud, _i32

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REF_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref",
                      "libpointops_ref_cuda.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    L = ctypes.CDLL(REF_SO)
    P, I = ctypes.c_void_p, ctypes.c_int
    L.knnquery_cuda_launcher.argtypes = [I, I, P, P, P, P, P, P]
    L.furthestsampling_cuda_launcher.argtypes = [I, I, P, P, P, P, P]
    return L


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


@pytest.mark.parametrize("n,m,ns", [(20000, 20000, 9), (20000, 5000, 17), (1250, 312, 17), (10, 10, 17)])
def test_reference_knn_kernel_vs_oracle_and_ours(ref, n, m, ns):
    xyz, _ = _cloud(n, 21)
    q = xyz[:m].contiguous()
    off, noff = _i32([n]), _i32([m])
    xd, qd, od, nd = xyz.to(DEV), q.to(DEV), off.to(DEV), noff.to(DEV)
    idx_r = torch.zeros(m, ns, dtype=torch.int32, device=DEV)
    d2_r = torch.zeros(m, ns, dtype=torch.float32, device=DEV)
    torch.cuda.synchronize()
    ref.knnquery_cuda_launcher(m, ns, _p(xd), _p(qd), _p(od), _p(nd), _p(idx_r), _p(d2_r))  # legacy default stream
    torch.cuda.synchronize()
    idx_o, d2_o = native.knn(ns, xyz, q, off, noff)
    assert torch.equal(d2_r.cpu(), d2_o)          # C oracle == reference kernel, bit for bit
    assert torch.equal(idx_r.cpu(), idx_o)        # including the heap's order among exact ties
    idx_g, dist_g = pointops.knnquery(ns, xd, qd, od, nd)
    assert torch.equal(dist_g.cpu(), torch.sqrt(d2_o.to(DEV)).cpu())   # IEEE sqrt, as torch.sqrt on CUDA
    assert torch.equal(idx_g.cpu(), idx_r.cpu())  # ours == the reference binary, bit for bit


@pytest.mark.parametrize("n", [20000, 5000, 1250, 312])
def test_reference_fps_kernel_vs_oracle_and_ours(ref, n):
    xyz, _ = _cloud(n, 22)
    off, noff = _i32([n]), _i32([n // 4])
    xd, od, nd = xyz.to(DEV), off.to(DEV), noff.to(DEV)
    idx_r = torch.zeros(n // 4, dtype=torch.int32, device=DEV)
    tmp = torch.full((n,), 1e10, device=DEV)
    torch.cuda.synchronize()
    ref.furthestsampling_cuda_launcher(1, n, _p(xd), _p(od), _p(nd), _p(tmp), _p(idx_r))
    torch.cuda.synchronize()
    assert torch.equal(idx_r.cpu(), native.fps(xyz, off, noff))
    assert torch.equal(pointops.furthestsampling(xd, od, nd).cpu(), idx_r.cpu())
