import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roitr_b200 import ops
DEV = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
def timeit(fn, iters=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
print("M,N,K, ffma_ms, tc_ms, tc_TFLOPs(fp32-equivalent), GB/s(A+C)")
for (M, N, K) in [(320000, 64, 64), (320000, 192, 64), (80000, 384, 128), (80000, 128, 128), (20000, 768, 256), (20000, 256, 256), (5000, 256, 256), (320000, 256, 64), (4992, 768, 256), (312, 768, 256), (312, 512, 256)]:
    a = torch.randn(M, K, device=DEV); w = torch.randn(N, K, device=DEV); b = torch.randn(N, device=DEV); out = torch.empty(M, N, device=DEV)
    from roitr_b200 import engine; wp = engine.pack_linear_tc(w)
    t1 = timeit(lambda: ops.linear(a, w, b, out=out, tc=False)); t2 = timeit(lambda: ops.linear(a, w, b, out=out, wpack=wp))
    print("%d,%d,%d, %.4f, %.4f, %.1f, %.0f" % (M, N, K, t1, t2, 2.0 * M * N * K / t2 / 1e9, (M * K + M * N) * 4 / t2 / 1e6))
