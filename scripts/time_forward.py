"""Eager single-pair forward latency + per-kernel breakdown (torch profiler) at the benchmark size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roitr_b200 import model
from roitr_b200.synthetic import synthetic_pair, forward_args
from tests.helpers import CONFIG_3D, weights
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
m = model.create_model(CONFIG_3D); m.load_state_dict(weights(1)); m = m.cuda().eval()
args = forward_args(synthetic_pair(0, N), "cuda:0")
for _ in range(3): out = m(*args)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = m(*args); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ts.sort(); print("N=%d forward median %.3f ms  min %.3f ms  (corr %d)" % (N, ts[5], ts[0], out["corr_scores"].shape[0]))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): m(*args)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
