mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -W ignore > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], round(d["roofline"]["frac"],4), d["acceptance_rate"], d.get("swap_acceptance_rate"), d["esjd"])'
for spec in "c3 500000 none" "c2 1000000 none" "c4 2000 all" "c5 20000 none" "c5f 20000 none"; do set -- $spec; echo -n "$1 T=$2 store=$3: "; timeout 600 python bench.py --workload $1 --steps 3 --warmup 3 --no-cpu --no-e2e --T $2 --store $3 2>&1 | tail -1 | python -c "$P"; done 2>&1 | tee gpurun_out/final_ab.log
