"""Micro-benchmark of the native point ops vs the reference's own kernels (oracle/_ref) on one GPU. CUDA-event timing,
warm-up, L2 flush between iterations. Not the headline bench (that is bench.py)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from roitr_b200 import pointops  # noqa: E402
from roitr_b200.synthetic import synthetic_pair  # noqa: E402

DEV = "cuda:0"
flush = None


def timeit(fn, iters=10, warm=3):
    global flush
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ref = None
    so = os.path.join(ROOT, "oracle", "_ref", "libpointops_ref_cuda.so")
    if os.path.exists(so):
        ref = ctypes.CDLL(so)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
    pair = synthetic_pair(0, 20000)
    xyz, nrm = pair["tgt_pcd"].to(DEV), pair["tgt_normals"].to(DEV)
    levels = [(xyz, nrm)]
    for n in (20000, 5000, 1250):
        o, no = i32([n]), i32([n // 4])
        idx = pointops.furthestsampling(levels[-1][0], o, no).long()
        levels.append((levels[-1][0][idx].contiguous(), levels[-1][1][idx].contiguous()))
    print("# ours = libroitr_b200 (kNN FUSED with PPF where the op is knn_ppf); reference_kernel = the reference's own .cu compiled")
    print("# unmodified for sm_100a (oracle/_ref: knnquery_cuda_kernel.cu:65-108 - kNN only, no PPF; sampling_cuda_kernel.cu:14-129).")
    print("# One cloud per call (batch size 1, as the reference runs); median of 10, 256 MiB L2 flush between iterations.")
    print("op, n_ref, m_query, k, ours_ms, reference_kernel_ms")
    for (li, lq, k) in [(0, 0, 8), (0, 1, 16), (1, 1, 16), (1, 2, 16), (2, 2, 16), (2, 3, 16), (3, 3, 16)]:
        (x, xn), (q, qn) = levels[li], levels[lq]
        n, m = x.shape[0], q.shape[0]
        o, no = i32([n]), i32([m])
        t = timeit(lambda: pointops.knn_ppf(k, x, xn, q, qn, o, no))
        tr = float("nan")
        if ref is not None:
            idx = torch.zeros(m, k + 1, dtype=torch.int32, device=DEV)
            d2 = torch.zeros(m, k + 1, device=DEV)
            tr = timeit(lambda: ref.knnquery_cuda_launcher(m, k + 1, p(x), p(q), p(o), p(no), p(idx), p(d2)))
        print("knn_ppf, %d, %d, %d, %.4f, %.4f" % (n, m, k, t, tr))
    for (lc, lf) in [(3, 2), (2, 1), (1, 0)]:
        (x, _), (q, _) = levels[lc], levels[lf]
        o, no = i32([x.shape[0]]), i32([q.shape[0]])
        t = timeit(lambda: pointops.knnquery(3, x, q, o, no))
        tr = float("nan")
        if ref is not None:
            m = q.shape[0]
            idx = torch.zeros(m, 3, dtype=torch.int32, device=DEV)
            d2 = torch.zeros(m, 3, device=DEV)
            tr = timeit(lambda: ref.knnquery_cuda_launcher(m, 3, p(x), p(q), p(o), p(no), p(idx), p(d2)))
        print("knn3, %d, %d, 3, %.4f, %.4f" % (x.shape[0], q.shape[0], t, tr))
    for li, n in enumerate((20000, 5000, 1250)):
        x = levels[li][0]
        o, no = i32([n]), i32([n // 4])
        for cl in (1, 2, 4, 8):
            try:
                t = timeit(lambda: pointops.furthestsampling(x, o, no, n_max=n, m_total=n // 4, cluster=cl), iters=5, warm=1)
            except Exception as e:  # capacity
                t = float("nan")
            print("fps_cluster%d, %d, %d, 0, %.4f, nan" % (cl, n, n // 4, t))
        if ref is not None:
            idx = torch.zeros(n // 4, dtype=torch.int32, device=DEV)
            def run():
                tmp = torch.full((n,), 1e10, device=DEV)
                ref.furthestsampling_cuda_launcher(1, n, p(x), p(o), p(no), p(tmp), p(idx))
            print("fps_reference_kernel, %d, %d, 0, nan, %.4f" % (n, n // 4, timeit(run, iters=5, warm=1)))
    # batched FPS throughput: 64 clouds of 20000 in one launch
    B = 64
    xb = torch.cat([synthetic_pair(i, 20000)["tgt_pcd"] for i in range(B)]).to(DEV)
    o = i32([20000 * (i + 1) for i in range(B)])
    no = i32([5000 * (i + 1) for i in range(B)])
    for cl in (1, 2):
        t = timeit(lambda: pointops.furthestsampling(xb, o, no, n_max=20000, m_total=5000 * B, cluster=cl), iters=3, warm=1)
        print("fps_batch%d_cluster%d, 20000, 5000, 0, %.4f (%.4f ms/cloud), nan" % (B, cl, t, t / B))


if __name__ == "__main__":
    main()
