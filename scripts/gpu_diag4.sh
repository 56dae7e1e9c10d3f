#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
for cfg in "1 1 16" "2 0 16"; do
set -- $cfg
ROITR_PIPELINE=$1 ROITR_MID_LEVEL=$2 timeout 600 python bench.py --steps 12 --warmup 4 --batch $3 --no-cpu-baseline > gpurun_out/bench_p$1_m$2_b$3.json 2> gpurun_out/bench_p$1_m$2_b$3.err
echo "pipeline=$1 mid=$2 B=$3 rc=$? $(python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_p$1_m$2_b$3.json"))
    print("value %.1f e2e %.1f ms/step %.2f serial %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["serial_replica_ms"]))
    print(json.dumps(d["kernel_shares_ms_per_step"]))
except Exception as e:
    print("ERR", e)
PY
)"
tail -2 gpurun_out/bench_p$1_m$2_b$3.err
done
python scripts/timeline.py 16 20000 > gpurun_out/timeline.log 2>&1; echo "timeline rc=$?"; tail -3 gpurun_out/timeline.log
