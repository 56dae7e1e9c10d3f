#!/usr/bin/env bash
# round-2 diagnostic call: parity tables at every BASELINE config (+fp64 yardstick), per-op ours-vs-reference kernels,
# LN256 experimental check, bench line with the new keys.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 1500 python scripts/parity_report.py small unscaled 20k 30k 4d8k > gpurun_out/parity_report.txt 2>&1; echo "parity rc=$?"
grep -E "^====|FAILURES" gpurun_out/parity_report.txt
timeout 600 python scripts/bench_ops.py > gpurun_out/bench_ops.txt 2>&1; echo "bench_ops rc=$?"; cat gpurun_out/bench_ops.txt
ROITR_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_experimental.py -q --tb=short > gpurun_out/experimental.log 2>&1; echo "experimental rc=$?"; tail -15 gpurun_out/experimental.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
