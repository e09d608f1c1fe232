# A/B of two library builds over the storing / non-storing workloads: VARIANTS="main _head"
mkdir -p gpurun_out
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], round(d["roofline"]["frac"],4), d["acceptance_rate"], d.get("swap_acceptance_rate"), d["esjd"])'
for v in ${VARIANTS:-main _head}; do [ "$v" = main ] && v=""; lib=$PWD/rwm_pt_pytorch_b200/librwmpt$v.so; [ -f $lib ] || continue
for spec in "c3 500000 none" "c4 2000 all" "c4 20000 all" "c2 500000 none" "c3 20000 all" "c3 100000 cold" "c2 20000 all"; do set -- $spec; echo -n "variant[$v] $1 T=$2 store=$3: "; RWMPT_LIB=$lib timeout 600 python bench.py --workload $1 --steps 3 --warmup 3 --no-cpu --no-e2e --T $2 --store $3 2>&1 | tail -1 | python -c "$P"; done; done 2>&1 | tee gpurun_out/ab_store.log
