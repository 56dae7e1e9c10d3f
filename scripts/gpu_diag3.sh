#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/test_batch_gpu.py tests/test_ransac.py tests/test_forward_gpu.py -q -m gpu --timeout 600 --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
for cfg in "1 1 16" "2 1 16" "2 0 16" "2 2 16" "2 1 8" "3 1 8" "2 1 12"; do
set -- $cfg
ROITR_PIPELINE=$1 ROITR_MID_LEVEL=$2 timeout 600 python bench.py --steps 12 --warmup 4 --batch $3 --no-cpu-baseline > gpurun_out/bench_p$1_m$2_b$3.json 2> gpurun_out/bench_p$1_m$2_b$3.err
echo "pipeline=$1 mid=$2 B=$3 rc=$? $(python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_p$1_m$2_b$3.json"))
    print("value %.1f e2e %.1f ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
except Exception as e:
    print("ERR", e)
PY
)"
tail -2 gpurun_out/bench_p$1_m$2_b$3.err
done
