# sweep an environment variable over values for c3/c2: VAR=NAME VALS="a b c" bash scripts/gpu_env_sweep.sh
mkdir -p gpurun_out
for v in $VALS; do for wl in ${WLS:-c3 c2}; do echo -n "$VAR=$v $wl: "; env $VAR=$v ${EXTRA_ENV:-RWMPT_DUMMY=1} python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-e2e --lanes ${LANES:-4} --T 50000 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['value'], d['roofline']['frac'], d['acceptance_rate'])"; done; done 2>&1 | tee gpurun_out/env_sweep.log
