#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY (see oracle/README.md).
#
# Compiles the reference's own two hot-path CUDA kernels, UNMODIFIED and from where
# they lie under /root/reference, for sm_100a. Output: oracle/_ref/libpointops_ref_cuda.so
# exporting the reference's extern "C" launchers
#   knnquery_cuda_launcher          (cpp_wrappers/pointops/src/knnquery/knnquery_cuda_kernel.h:13)
#   furthestsampling_cuda_launcher  (cpp_wrappers/pointops/src/sampling/sampling_cuda_kernel.h:13)
# The reference's *_cuda.cpp pybind wrappers are NOT built (they include THC/THC.h which
# modern torch no longer ships); the launchers are already a C ABI.
#
# Only runs where /root/reference exists (this container). The GPU box uses the prebuilt .so
# that travels with the gpurun snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).
set -euo pipefail
REF=${ROITR_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
SRC="$REF/cpp_wrappers/pointops/src"
if [ ! -d "$SRC" ]; then echo "reference not present at $REF; skipping oracle/_ref build"; exit 0; fi
mkdir -p "$OUT"
TORCH_INC=$(python - <<'EOF'
import os, torch
r = os.path.join(os.path.dirname(torch.__file__), "include")
print("-I%s -I%s" % (r, os.path.join(r, "torch/csrc/api/include")))
EOF
)
PY_INC=$(python -c "import sysconfig; print('-I'+sysconfig.get_paths()['include'])")
if [ "$OUT/libpointops_ref_cuda.so" -nt "$SRC/knnquery/knnquery_cuda_kernel.cu" ] && [ -z "${FORCE:-}" ]; then
  echo "oracle/_ref up to date"; exit 0; fi
# -O2 matches the reference's setup.py (cpp_wrappers/pointops/setup.py:27)
for f in knnquery/knnquery_cuda_kernel sampling/sampling_cuda_kernel; do
  nvcc -O2 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -std=c++17 \
       $TORCH_INC $PY_INC -c "$SRC/$f.cu" -o "$OUT/$(basename $f).o" &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libpointops_ref_cuda.so" \
     "$OUT/knnquery_cuda_kernel.o" "$OUT/sampling_cuda_kernel.o"
echo "built $OUT/libpointops_ref_cuda.so"
