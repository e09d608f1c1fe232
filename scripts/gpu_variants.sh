# A/B of compile-time variants of the fast loop (source order of the three streams, accumulator-flush chunk) on C3 and C2.
mkdir -p gpurun_out
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], round(d["roofline"]["frac"],4), d["acceptance_rate"], d.get("swap_acceptance_rate"), d["esjd"])'
for v in ${VARIANTS:-main _o1 _o2 _o3 _c16 _c64}; do [ "$v" = main ] && v=""; lib=$PWD/rwm_pt_pytorch_b200/librwmpt$v.so; [ -f $lib ] || continue
for wl in c3 c2; do echo -n "variant[$v] $wl: "; RWMPT_LIB=$lib timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-e2e --T 400000 2>&1 | tail -1 | python -c "$P"; done; done 2>&1 | tee gpurun_out/variants.log
