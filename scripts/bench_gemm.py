"""Dense-layer micro-benchmark at the bench step's shapes (B = 16 pairs): the streaming kernel (linear_tc3) and, with a row
gather (the same kernel, GATHER instantiation: 16-byte cp.async pieces of the indexed rows), with a 256 MiB L2 flush before every timed launch.
usage: bench_gemm.py   (run on a GPU box; writes gpurun_out/bench_gemm.txt)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from roitr_b200 import _lib, engine, ops
DEV = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
def timeit(fn, iters=7):
    for _ in range(2): fn()
    ts = []
    for _ in range(iters):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
out = open(os.path.join(ROOT, "gpurun_out", "bench_gemm.txt"), "w")
def log(s):
    print(s); out.write(s + "\n"); out.flush()
log("M,N,K, streaming_ms (linear_tc3), gathered_rows_ms (linear_tc3, a_index), streaming GB/s(A+C+W), frac of measured HBM 6532 GB/s")
SHAPES = [(640000, 64, 64), (640000, 192, 64), (640000, 128, 64), (640000, 256, 64), (640000, 384, 128), (160000, 128, 128),
          (160000, 384, 128), (160000, 256, 128), (160000, 768, 256), (40000, 256, 256), (40000, 768, 256), (9984, 256, 256),
          (9984, 768, 256), (4992, 256, 256), (4992, 512, 256), (4992, 256, 512), (4992, 768, 256), (4992, 1024, 256), (4992, 256, 1024)]
for (M, N, K) in SHAPES:
    a = torch.randn(M, K, device=DEV); w = torch.randn(N, K, device=DEV) / K ** 0.5; b = torch.randn(N, device=DEV)
    o3 = torch.empty(M, N, device=DEV)
    wp = engine.pack_linear_tc(w)
    t3 = timeit(lambda: ops.linear(a, w, b, out=o3, wpack=wp))
    idx = torch.randperm(M, device=DEV).int()
    t2 = timeit(lambda: ops.linear(a, w, b, out=o3, wpack=wp, a_index=idx))      # gathered rows
    gbs = (M * K + M * N + N * K) * 4 / t3 / 1e6
    log("%d,%d,%d, %.4f, %.4f, %.0f, %.3f" % (M, N, K, t3, t2, gbs, gbs / 6532.5))

log("fused LayerNorm epilogue: M,N,K, plain_ms, row_epilogue_ms, fused_ms (res_pre), fused_ms (res_post+relu)")
for (M, N, K) in [(640000, 64, 64), (160000, 128, 128)]:
    a = torch.randn(M, K, device=DEV); w = torch.randn(N, K, device=DEV) / K ** 0.5; b = torch.randn(N, device=DEV)
    g_, be = torch.ones(N, device=DEV), torch.zeros(N, device=DEV)
    r = torch.randn(M, N, device=DEV)
    wp = engine.pack_linear_tc(w)
    o = torch.empty(M, N, device=DEV)
    t0 = timeit(lambda: ops.linear(a, w, b, out=o, wpack=wp))
    t1 = timeit(lambda: ops.row_epilogue(o, res_pre=r, gamma=g_, beta=be, mode=ops.MODE_LN))
    t2 = timeit(lambda: ops.linear_ln(a, w, b, wp, g_, be, res_pre=r))
    t3 = timeit(lambda: ops.linear_ln(a, w, b, wp, g_, be, res_post=r, relu=True))
    log("%d,%d,%d, %.4f, %.4f, %.4f, %.4f" % (M, N, K, t0, t1, t2, t3))
