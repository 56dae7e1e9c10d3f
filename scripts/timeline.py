"""Timeline of one step INSIDE its CUDA graph: start/end (ms from the step start) of every C-ABI call with its stream.
Events are recorded into the capture with cudaEventRecordWithFlags(..., cudaEventRecordExternal), so after a replay their
timestamps are those of the real multi-stream graph execution (not of an eager, CPU-launch-bound run).
usage: timeline.py [B] [N]   (GPU box; writes gpurun_out/timeline.txt and gpurun_out/timeline.json (chrome://tracing))"""
import collections, ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from roitr_b200 import _lib, model, ops
from roitr_b200.synthetic import synthetic_pair
from tests.helpers import CONFIG_3D, weights
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
dev = torch.device("cuda", 0)
torch.cuda.init()
rt = None
for name in ("libcudart.so.12", "libcudart.so"):
    try:
        rt = ctypes.CDLL(name); break
    except OSError:
        pass
if rt is None:
    import glob
    rt = ctypes.CDLL(sorted(glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*")))[0])
def ev_create():
    e = ctypes.c_void_p()
    assert rt.cudaEventCreate(ctypes.byref(e)) == 0
    return e
def ev_record(e, stream):
    rc = rt.cudaEventRecordWithFlags(e, ctypes.c_void_p(stream), ctypes.c_uint(1))      # cudaEventRecordExternal
    assert rc == 0, rc
def ev_ms(a, b):
    ms = ctypes.c_float()
    rc = rt.cudaEventElapsedTime(ctypes.byref(ms), a, b)
    assert rc == 0, rc
    return ms.value
m = model.create_model(dict(CONFIG_3D)); m.load_state_dict(weights(1), strict=True); m = m.to(dev).eval()
pairs = [{k: v.to(dev) for k, v in synthetic_pair(g, N).items()} for g in range(B)]
REC = []
ON = [False]
orig_call = _lib.call
def call(name, *args):
    if not ON[0]:
        return orig_call(name, *args)
    st = torch.cuda.current_stream().cuda_stream
    e0, e1 = ev_create(), ev_create()
    ev_record(e0, st)
    orig_call(name, *args)
    ev_record(e1, st)
    ints = []
    for a in args:
        if isinstance(a, _lib.c_int): ints.append(a.value)
        else: break
    REC.append((name, st, e0, e1, tuple(ints[:5])))
_lib.call = call
ops._lib.call = call
r = m.batch_runner(B, N, N, graph=True)
r.load(pairs)
# BatchRunner.run(): one eager pass, then capture; switch recording on only for the capture pass
r.outs, r.counts = r._body(); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
T0, T1 = ev_create(), ev_create()
with torch.cuda.graph(g):
    ev_record(T0, torch.cuda.current_stream().cuda_stream)
    ON[0] = True
    r.outs, r.counts = r._body()
    ON[0] = False
    ev_record(T1, torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.zero_(); g.replay(); torch.cuda.synchronize()
total = ev_ms(T0, T1)
streams, rows = {}, []
for name, sid, e0, e1, ints in REC:
    s = streams.setdefault(sid, len(streams))
    rows.append((ev_ms(T0, e0), ev_ms(T0, e1), s, name.replace("roitr_", ""), ints))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "timeline.txt"), "w") as o:
    o.write("# B=%d N=%d step inside its CUDA graph: %.3f ms, %d calls, %d streams (s0 = first stream seen)\n" % (B, N, total, len(rows), len(streams)))
    busy = collections.defaultdict(float)
    for a, b, s, n, i in rows: busy[s] += b - a
    o.write("# sum of call durations per stream (ms): " + ", ".join("s%d=%.2f" % (s, busy[s]) for s in sorted(busy)) + "\n")
    for a, b, s, n, i in sorted(rows):
        o.write("%8.3f %8.3f %7.3f  s%-2d %-28s %s\n" % (a, b, b - a, s, n, i))
json.dump([{"name": n, "ph": "X", "ts": a * 1000, "dur": (b - a) * 1000, "pid": 0, "tid": s, "args": {"ints": list(i)}} for a, b, s, n, i in rows],
          open(os.path.join(ROOT, "gpurun_out", "timeline.json"), "w"))
print(open(os.path.join(ROOT, "gpurun_out", "timeline.txt")).read()[:400])
