set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout -k 10 600 python scripts/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; echo "bench_gemm rc=$?"; cat gpurun_out/bench_gemm.txt
bash scripts/gpu_check.sh
