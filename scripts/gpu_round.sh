#!/usr/bin/env bash
# One gpurun call: tests, micro-bench, launch list + one full ncu capture. Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?" | tee -a gpurun_out/build.log
timeout 900 python -m pytest tests -q -m gpu --timeout 300 --tb=short ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 600 python scripts/bench_ops.py > gpurun_out/bench_ops.csv 2> gpurun_out/bench_ops.err; echo "bench_ops rc=$?"
cat gpurun_out/bench_ops.csv
