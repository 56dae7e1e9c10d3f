"""Micro-benchmarks used to pick launch parameters on a B200: grid-kNN cell occupancy target and FPS cluster size for the
bench batch (32 clouds x 20000 points). usage: tune_ops.py  (writes gpurun_out/tune_ops.txt)"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from roitr_b200 import _lib, ops
from roitr_b200.synthetic import synthetic_pair
dev = torch.device("cuda", 0)
B, N = 16, 20000
pairs = [synthetic_pair(g, N) for g in range(B)]
pts = torch.cat([p["src_pcd"] for p in pairs] + [p["tgt_pcd"] for p in pairs]).to(dev)
nrm = torch.cat([p["src_normals"] for p in pairs] + [p["tgt_normals"] for p in pairs]).to(dev)
o = torch.tensor([N * (i + 1) for i in range(2 * B)], dtype=torch.int32, device=dev)
M = N // 4
qo = torch.tensor([M * (i + 1) for i in range(2 * B)], dtype=torch.int32, device=dev)
sel = torch.cat([torch.arange(M, device=dev) * 4 + i * N for i in range(2 * B)])
q, qn = pts[sel].contiguous(), nrm[sel].contiguous()
out = open(os.path.join(ROOT, "gpurun_out", "tune_ops.txt"), "w")

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]

def log(s):
    print(s); out.write(s + "\n"); out.flush()

ref = None
for thread_q in (0, 1):
    _lib.lib().roitr_debug_knn_thread_per_query(thread_q)
    for target in ((1.0, 1.5, 2.0) if not thread_q else (0.5, 1.0, 1.5, 2.0, 3.0, 4.0)):
        _lib.lib().roitr_debug_set_knn_grid_target(ctypes.c_float(target))
        t_build = timeit(lambda: ops.knn_grid_build(pts, o))
        grid = ops.knn_grid_build(pts, o)
        gq = ops.knn_grid_build(q, qo)
        t_self = timeit(lambda: ops.knn_ppf(8, pts, nrm, pts, nrm, o, o, grid=grid))
        t_down = timeit(lambda: ops.knn_ppf(16, pts, nrm, q, qn, o, qo, grid=grid, qgrid=gq))
        t_down_n = timeit(lambda: ops.knn_ppf(16, pts, nrm, q, qn, o, qo, grid=grid))
        t_one = timeit(lambda: ops.knn_ppf(1, pts, None, pts, None, o, o, drop_first=0, want_ppf=False, want_dist=True, grid=grid))
        t_up = timeit(lambda: ops.knn_ppf(3, q, None, pts, None, qo, o, drop_first=0, want_ppf=False, want_dist=True, grid=gq, qgrid=grid))
        t_up_n = timeit(lambda: ops.knn_ppf(3, q, None, pts, None, qo, o, drop_first=0, want_ppf=False, want_dist=True, grid=gq))
        idx = ops.knn_ppf(8, pts, nrm, pts, nrm, o, o, grid=grid)[0]
        idx2 = ops.knn_ppf(16, pts, nrm, q, qn, o, qo, grid=grid, qgrid=gq)[0]
        if ref is None:
            ref = (idx.clone(), idx2.clone())
        same = bool(torch.equal(idx, ref[0]) and torch.equal(idx2, ref[1]))
        log("thread-per-query %d grid target %5.1f: build %.3f ms | self k=9 (640k q) %.3f ms | down k=17 (160k q) %.3f (natural order %.3f) ms | "
            "k=1 self (640k q) %.3f ms | interp k=3 (640k q, 160k refs) %.3f (natural order %.3f) ms | identical to warp kernel: %s"
            % (thread_q, target, t_build, t_self, t_down, t_down_n, t_one, t_up, t_up_n, same))
_lib.lib().roitr_debug_set_knn_grid_target(ctypes.c_float(1.0))
# brute force vs grid on the 1250-point level
M2 = M // 4
q2o = torch.tensor([M2 * (i + 1) for i in range(2 * B)], dtype=torch.int32, device=dev)
sel2 = torch.cat([torch.arange(M2, device=dev) * 4 + i * M for i in range(2 * B)])
p2, n2 = q[sel2].contiguous(), qn[sel2].contiguous()
g2 = ops.knn_grid_build(p2, q2o)
for name, g in (("brute", None), ("grid", g2)):
    t_a = timeit(lambda: ops.knn_ppf(16, p2, n2, p2, n2, q2o, q2o, grid=g))
    t_b = timeit(lambda: ops.knn_ppf(3, p2, None, q, None, q2o, qo, drop_first=0, want_ppf=False, want_dist=True, grid=g))
    log("1250-pt level %s: self k=17 (40k q) %.3f ms | interp k=3 (160k q) %.3f ms" % (name, t_a, t_b))
a = ops.knn_ppf(16, p2, n2, p2, n2, q2o, q2o, grid=None)[0]; b = ops.knn_ppf(16, p2, n2, p2, n2, q2o, q2o, grid=g2)[0]
log("1250-pt level grid == brute: %s" % bool(torch.equal(a, b)))
for cl in (1, 2, 4, 8):
    t = timeit(lambda: ops.fps(pts, o, qo, N, 2 * B * M, cluster=cl), reps=3)
    log("fps 32 clouds 20000->5000 cluster=%d: %.3f ms" % (cl, t))
for nb in (8, 16, 24, 32, 36):
    oo = o[:nb]; qq = qo[:nb]
    t = timeit(lambda: ops.fps(pts[:nb * N], oo, qq, N, nb * M, cluster=0), reps=3)
    log("fps %d clouds auto cluster: %.3f ms" % (nb, t))
