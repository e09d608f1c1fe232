# Lanes-per-chain / units sweep on the existing kernels (generic instantiations where no tuned one exists): one line per point.
#   bash scripts/gpu_lanes_sweep.sh <tag> "<workload>:<lanes>:<units>:<T> ..."
TAG=${1:-sweep}; shift
mkdir -p gpurun_out
OUT=gpurun_out/lanes_sweep_$TAG.txt
: > $OUT
for p in $1; do
  IFS=: read wl lanes units T <<< "$p"
  line=$(timeout 300 python bench.py --workload $wl --lanes $lanes --units $units --T $T --steps 3 --warmup 3 --no-cpu --no-e2e --no-aux --also none 2>&1 | tail -1)
  echo "$p $(echo "$line" | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print(d['value'], d['config']['geometry_E_W'], round(d['roofline']['frac'], 4), d['acceptance_rate'], d.get('swap_acceptance_rate'), d['esjd'])
except Exception as e:
    print('FAILED', e)
")" | tee -a $OUT
done
