"""Timing of the two steps either side of the hot path at the bench shapes, GPU kernel vs the CPU oracle.
usage: bench_pre_post.py  (GPU box; writes gpurun_out/bench_pre_post.txt)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import normals_ref, procrustes_ref
from roitr_b200 import preprocess, registration
from roitr_b200.synthetic import synthetic_pair
dev = torch.device("cuda", 0)
out = open(os.path.join(ROOT, "gpurun_out", "bench_pre_post.txt"), "w")
def log(s):
    print(s); out.write(s + "\n"); out.flush()
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
B, N = 16, 20000
pairs = [synthetic_pair(g, N) for g in range(B)]
pts = torch.cat([p["src_pcd"] for p in pairs] + [p["tgt_pcd"] for p in pairs]).to(dev)
off = torch.tensor([N * (i + 1) for i in range(2 * B)], dtype=torch.int32, device=dev)
t_gpu = timeit(lambda: preprocess.estimate_normals(pts, off, knn=33))
t0 = time.perf_counter(); normals_ref.estimate_normals(pairs[0]["src_pcd"].numpy(), 33); t_cpu = time.perf_counter() - t0
log("normals knn=33: %d clouds x %d points: GPU %.3f ms (%.1f us per cloud pair... %.1f M points/s) | CPU oracle (scipy cKDTree + eigh, 1 cloud) %.1f ms -> %.0f x per cloud"
    % (2 * B, N, t_gpu, 1000 * t_gpu / B, 2 * B * N / t_gpu / 1e3, 1000 * t_cpu, 1000 * t_cpu / (t_gpu / (2 * B))))
g = torch.Generator().manual_seed(0)
src = torch.randn(B, 3200, 3, generator=g); tgt = torch.randn(B, 3200, 3, generator=g); w = torch.rand(B, 3200, generator=g)
sc, tc, wc = src.to(dev), tgt.to(dev), w.to(dev)
t_gpu = timeit(lambda: registration.weighted_procrustes(sc, tc, wc))
t0 = time.perf_counter(); procrustes_ref.weighted_procrustes(src, tgt, w); t_cpu = time.perf_counter() - t0
log("weighted_procrustes: %d pairs x 3200 correspondences: GPU %.3f ms | CPU oracle (torch.svd) %.2f ms" % (B, t_gpu, 1000 * t_cpu))
