#!/usr/bin/env bash
# pipeline depth / release point / batch size sweep of the bench (no CPU legs)
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
for cfg in "2 0 16" "3 0 16" "2 0 24" "2 1 16" "2 0 12" "2 0 32"; do
set -- $cfg
ROITR_PIPELINE=$1 ROITR_MID_LEVEL=$2 timeout -k 10 600 python bench.py --steps 12 --warmup 4 --batch $3 --no-cpu-baseline > gpurun_out/sw.json 2> gpurun_out/sw.err
echo "pipeline=$1 mid=$2 B=$3 rc=$? $(python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sw.json"))
    print("value %.1f e2e %.1f ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
except Exception as e:
    print("ERR", e)
PY
)"
tail -2 gpurun_out/sw.err
done
