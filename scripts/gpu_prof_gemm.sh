#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
for shp in "640000 64 64" "640000 192 64" "640000 256 64" "9984 1792 256"; do
  tag=$(echo $shp | tr ' ' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc3 -c 1 -f -o gpurun_out/gemm_$tag python scripts/prof_gemm.py $shp 0 > gpurun_out/ncu_gemm_$tag.log 2>&1; echo "ncu $tag rc=$?"
done
ls -la gpurun_out/*.ncu-rep
