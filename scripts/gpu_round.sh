#!/usr/bin/env bash
# One gpurun call: build, GPU tests, smoke, bench, ncu launch list + full captures. Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 --tb=short -x ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -5
grep -E "^(FAILED|ERROR)|Error|assert" gpurun_out/pytest_gpu.log | head -20
fi
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
ROITR_PROFILE_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
for k in ${NCU_KERNELS:-fps_cluster_kernel knn_ppf_kernel}; do
ROITR_PROFILE_RANGE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c ${NCU_COUNT:-3} -f -o gpurun_out/prof_$k python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
fi
