# config 4 beyond its named size: ladders per GPU 512 (named) ... 4096 with all 8 chains stored (T shrinks to what fits in HBM), and
# the named size without the stores -- is the kernel or the size of the workload the limit?   bash scripts/gpu_c4_units.sh <tag>
TAG=${1:-c4u}
mkdir -p gpurun_out
OUT=gpurun_out/c4_units_$TAG.txt
: > $OUT
run() {  # units T store
  line=$(timeout 300 python bench.py --workload c4 --units $1 --T $2 --store $3 --steps 3 --warmup 3 --no-cpu --no-e2e --no-aux --also none 2>&1 | tail -1)
  echo "c4 units=$1 T=$2 store=$3 $(echo "$line" | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print(d['value'], d['config']['geometry_E_W'], d['roofline']['bound'], round(d['roofline']['frac'], 4), 'GB/s', round(d['roofline']['achieved'], 1), d['config'].get('steps_per_launch'), d['acceptance_rate'])
except Exception as e:
    print('FAILED', e)
")" | tee -a $OUT
}
run 512 100000 all
run 512 100000 none
run 1024 50000 all
run 2048 25000 all
run 4096 12000 all
run 4096 12000 none
