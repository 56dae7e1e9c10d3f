"""Per-call breakdown of one eager batched step: entry point + leading integer arguments -> CUDA-event time.
usage: step_breakdown.py [B] [N]   (run on a GPU box; writes gpurun_out/step_breakdown.txt)"""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from roitr_b200 import _lib, model
from roitr_b200.synthetic import synthetic_pair
from tests.helpers import CONFIG_3D, weights
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
dev = torch.device("cuda", 0)
m = model.create_model(dict(CONFIG_3D)); m.load_state_dict(weights(1), strict=True); m = m.to(dev).eval()
pairs = [{k: v.to(dev) for k, v in synthetic_pair(g, N).items()} for g in range(B)]
r = m.batch_runner(B, N, N, graph=False)
r.load(pairs); r.run(); torch.cuda.synchronize()
for k in _lib.KERNELS_PER_CALL:
    _lib.TIMED[k] = []
_lib.reset_stats(); _lib.RECORD_ARGS = True
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); r.load(pairs); r.run(); b.record(); torch.cuda.synchronize()
agg = collections.OrderedDict()
for name, evs in _lib.TIMED.items():
    for (e0, e1), ints in zip(evs, _lib.ARGS.get(name, [])):
        key = (name, tuple(ints[:6]))
        t = agg.setdefault(key, [0, 0.0]); t[0] += 1; t[1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "step_breakdown.txt"), "w") as o:
    o.write("# B=%d N=%d eager step %.3f ms wall(dev), sum of calls %.3f ms\n" % (B, N, a.elapsed_time(b), tot))
    for (name, ints), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        o.write("%-30s %-40s n=%4d %9.3f ms %5.1f%%\n" % (name, ints, n, ms, 100 * ms / tot))
print(open(os.path.join(ROOT, "gpurun_out", "step_breakdown.txt")).read()[:6000])
