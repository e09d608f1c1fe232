# config 2 at d = 30: fused kernel (auto 4 x 8, tuned 8 x 4) against the warp-specialised kernel on 8 x 4 / 4 x 8 lanes with
# 1..4 producer warps.   bash scripts/gpu_c2d30_sweep.sh <tag>
TAG=${1:-c2d30}
mkdir -p gpurun_out
OUT=gpurun_out/c2d30_sweep_$TAG.txt
: > $OUT
run() {  # geom schedule producers
  [ "$1" = "-" ] && unset RWMPT_GEOM || export RWMPT_GEOM=$1
  export RWMPT_SCHEDULE=$2
  [ "$3" = "-" ] && unset RWMPT_SPEC_NP || export RWMPT_SPEC_NP=$3
  line=$(timeout 300 python bench.py --workload c2d30 --T 400000 --steps 3 --warmup 3 --no-cpu --no-e2e --no-aux --also none 2>&1 | tail -1)
  echo "geom=$1 schedule=$2 producers=$3 $(echo "$line" | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print(d['value'], d['config']['geometry_E_W'], round(d['roofline']['frac'], 4), d['acceptance_rate'], d['esjd'])
except Exception as e:
    print('FAILED', e)
")" | tee -a $OUT
}
run - 1 -
run 8,4 1 -
for np in 1 2 3 4; do run 8,4 3 $np; done
for np in 1 2 3; do run 4,8 3 $np; done
run - 0 -
