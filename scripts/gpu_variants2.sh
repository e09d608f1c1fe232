mkdir -p gpurun_out
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], round(d["roofline"]["frac"],4))'
for rep in 1 2 3; do for v in main _o2; do [ "$v" = main ] && v=""; lib=$PWD/rwm_pt_pytorch_b200/librwmpt$v.so
for wl in c2 c3; do echo -n "rep$rep variant[$v] $wl: "; RWMPT_LIB=$lib timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-e2e --T 400000 2>&1 | tail -1 | python -c "$P"; done; done; done 2>&1 | tee gpurun_out/variants2.log
