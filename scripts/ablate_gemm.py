"""Where does the streaming dense-layer kernel's time go? Times linear_tc3 with parts switched off (debug flags; results
are wrong by construction, only the timing is of interest). usage: ablate_gemm.py (GPU box; gpurun_out/ablate_gemm.txt)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from roitr_b200 import _lib, engine, ops
DEV = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
def timeit(fn, iters=5):
    for _ in range(2): fn()
    ts = []
    for _ in range(iters):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
out = open(os.path.join(ROOT, "gpurun_out", "ablate_gemm.txt"), "w")
def log(s):
    print(s); out.write(s + "\n"); out.flush()
MASKS = [(0, "full"), (1, "no C stores"), (2, "W once"), (4, "no MMA"), (8, "no split"), (16, "no A loads"), (3, "no C, W once"),
         (1 | 2 | 4 | 8, "A loads only"), (2 | 4 | 8 | 16, "C stores only"), (31, "sync skeleton")]
log("M,N,K | " + " | ".join(n for _, n in MASKS))
for (M, N, K) in [(640000, 64, 64), (640000, 192, 64), (160000, 128, 128), (640000, 384, 128), (160000, 768, 256), (4992, 256, 256)]:
    a = torch.randn(M, K, device=DEV); w = torch.randn(N, K, device=DEV) / K ** 0.5; b = torch.randn(N, device=DEV)
    o = torch.empty(M, N, device=DEV)
    wp = engine.pack_linear_tc(w)
    ts = []
    for mask, _ in MASKS:
        _lib.lib().roitr_debug_linear_ablate(mask)
        ts.append(timeit(lambda: ops.linear(a, w, b, out=o, wpack=wp)))
    _lib.lib().roitr_debug_linear_ablate(0)
    log("%d,%d,%d | " % (M, N, K) + " | ".join("%.4f" % t for t in ts))
    if N <= 64:
        vs = []
        for v in (0, 3):
            _lib.lib().roitr_debug_linear_variant(v)
            vs.append(timeit(lambda: ops.linear(a, w, b, out=o, wpack=wp)))
        _lib.lib().roitr_debug_linear_variant(0)
        log("   configurations (deep rings, light footprint): " + " | ".join("%.4f" % t for t in vs))
