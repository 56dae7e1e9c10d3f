# A/B of variant builds (roitr_b200/build.py:build_variant -> roitr_b200/lib/variants/*.so) on one box: the bench step with
# each library, one JSON line per variant in gpurun_out/ab.txt. usage: gpu_ab.sh [--breakdown]
set -u
mkdir -p gpurun_out
: > gpurun_out/ab.txt
for so in roitr_b200/lib/libroitr_b200.so roitr_b200/lib/variants/*.so; do
  [ -f "$so" ] || continue
  echo "== $so" >> gpurun_out/ab.txt
  ROITR_B200_LIB=$PWD/$so timeout -k 10 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2> gpurun_out/ab_err.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); k = d.get('kernel_shares_ms_per_step', {})
        print(json.dumps({'value': d['value'], 'e2e': d['e2e']['value'], 'ms': d['ms_per_step'], 'serial': d.get('serial_replica_ms'), 'top': dict(list(k.items())[:12])}))
" >> gpurun_out/ab.txt
done
if [ "${1:-}" = "--breakdown" ]; then timeout -k 10 300 python scripts/step_breakdown.py > gpurun_out/step_breakdown.log 2>&1; fi
cat gpurun_out/ab.txt
