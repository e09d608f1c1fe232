# A/B of tuned geometries / loop variants through the environment knobs of pick_geometry (RWMPT_GEOM="E,W", RWMPT_VARIANT):
#   bash scripts/gpu_variant_ab.sh <tag> "<workload>:<E,W or ->:<variant>:<units>:<T> ..."
TAG=${1:-ab}; shift
mkdir -p gpurun_out
OUT=gpurun_out/variant_ab_$TAG.txt
: > $OUT
for p in $1; do
  IFS=: read wl geom var units T <<< "$p"
  [ "$geom" = "-" ] && unset RWMPT_GEOM || export RWMPT_GEOM=$geom
  export RWMPT_VARIANT=$var
  line=$(timeout 300 python bench.py --workload $wl --units $units --T $T --steps 3 --warmup 3 --no-cpu --no-e2e --no-aux --also none 2>&1 | tail -1)
  echo "$p $(echo "$line" | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print(d['value'], d['config']['geometry_E_W'], round(d['roofline']['frac'], 4), d['acceptance_rate'], d.get('swap_acceptance_rate'), d['esjd'])
except Exception as e:
    print('FAILED', e)
")" | tee -a $OUT
done
