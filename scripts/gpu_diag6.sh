#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout -k 10 900 python -m pytest tests/test_gemm_gpu.py -q -m gpu --timeout 600 --tb=short -x > gpurun_out/pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"
tail -12 gpurun_out/pytest_gemm.log | cut -c1-300
timeout -k 10 600 python scripts/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; echo "bench_gemm rc=$?"; cat gpurun_out/bench_gemm.txt
timeout -k 10 1500 python -m pytest tests -q -m gpu --timeout 900 --tb=short -x --deselect tests/test_gemm_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
for cfg in "1 1 16 1" "2 0 16 1" "2 0 16 0"; do
set -- $cfg
ROITR_PIPELINE=$1 ROITR_MID_LEVEL=$2 ROITR_LN256=$4 timeout -k 10 600 python bench.py --steps 12 --warmup 4 --batch $3 --no-cpu-baseline > gpurun_out/bench_p$1_m$2_b$3_ln$4.json 2> gpurun_out/bench_p$1_m$2_b$3_ln$4.err
echo "pipeline=$1 mid=$2 B=$3 ln256=$4 rc=$? $(python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_p$1_m$2_b$3_ln$4.json"))
    print("value %.1f e2e %.1f ms/step %.2f serial %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["serial_replica_ms"]))
    print(json.dumps(d["kernel_shares_ms_per_step"]))
except Exception as e:
    print("ERR", e)
PY
)"
tail -2 gpurun_out/bench_p$1_m$2_b$3_ln$4.err
done
