#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout -k 10 600 python scripts/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; echo "bench_gemm rc=$?"; cat gpurun_out/bench_gemm.log | head -50
timeout -k 10 900 python -m pytest tests/test_gemm_gpu.py tests/test_forward_gpu.py tests/test_batch_gpu.py -q -m gpu --timeout 600 --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
for v3 in 1 0; do
ROITR_SELF_V3=$v3 timeout -k 10 600 python bench.py --steps 12 --warmup 4 --no-cpu-baseline > gpurun_out/bench_v3_$v3.json 2> gpurun_out/bench_v3_$v3.err
echo "self_v3=$v3 rc=$? $(python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_v3_$v3.json"))
    print("value %.1f e2e %.1f ms/step %.2f serial %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["serial_replica_ms"]))
    print(json.dumps(d["kernel_shares_ms_per_step"]))
except Exception as e:
    print("ERR", e)
PY
)"
tail -2 gpurun_out/bench_v3_$v3.err
done
