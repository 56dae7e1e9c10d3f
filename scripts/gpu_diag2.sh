#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 1500 python scripts/parity_report.py small unscaled 20k 30k 4d8k > gpurun_out/parity_report.txt 2>&1; echo "parity rc=$?"
grep -E "^====|FAILURES|flip" gpurun_out/parity_report.txt | cut -c1-400
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_gpu.log | cut -c1-300
