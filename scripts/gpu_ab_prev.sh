# A/B of the current library against librwmpt_prev.so (the previous build) on the workloads that use the ticketed schedule.
mkdir -p gpurun_out
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], round(d["roofline"]["frac"],4), d["acceptance_rate"], d.get("swap_acceptance_rate"), d["esjd"])'
for rep in 1 2; do for v in main _prev; do [ "$v" = main ] && v=""; lib=$PWD/rwm_pt_pytorch_b200/librwmpt$v.so; [ -f $lib ] || continue
for spec in "c3 500000" "c2 1000000"; do set -- $spec; echo -n "rep$rep variant[$v] $1: "; RWMPT_LIB=$lib timeout 600 python bench.py --workload $1 --steps 3 --warmup 3 --no-cpu --no-e2e --T $2 2>&1 | tail -1 | python -c "$P"; done; done; done 2>&1 | tee gpurun_out/ab_prev.log
