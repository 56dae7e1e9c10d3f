#!/usr/bin/env bash
# build, GPU tests, smoke, one bench line (optionally: BENCH_ENV="ROITR_PIPELINE=1" etc.)
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout -k 10 1500 python -m pytest tests -q -m gpu --timeout 900 --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
timeout -k 10 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout -k 10 900 python bench.py --steps ${BENCH_STEPS:-20} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    print("value %.1f e2e %.1f rec %.1f single %.2f ms serial %.2f ref_gpu %s" % (d["value"], d["e2e"]["value"], d["e2e_record"]["value"], d["single_pair_forward_ms"]["value"], d["serial_replica_ms"], d.get("reference_gpu", {}).get("value")))
    print(json.dumps(d["kernel_shares_ms_per_step"]))
except Exception as e:
    print("ERR", e)
PY
tail -3 gpurun_out/bench.err
