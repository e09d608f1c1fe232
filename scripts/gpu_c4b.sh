# C4 tuned kernel: the new test, ncu of the C4 kernel, and C3 schedule / run-length sensitivity.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -W ignore -k "config4 or laplace or three_mixture or ThreeMixture" > gpurun_out/pytest_c4.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_c4.log; tail -15 gpurun_out/pytest_c4.log
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], round(d["roofline"]["frac"],4), d["acceptance_rate"], d.get("swap_acceptance_rate"), d["esjd"], d["clocks"])'
for spec in "0 100000" "1 100000" "0 200000" "0 500000" "1 500000"; do set -- $spec; echo -n "c3 schedule=$1 T=$2: "; RWMPT_SCHEDULE=$1 timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu --no-e2e --T $2 2>&1 | tail -1 | python -c "$P"; done 2>&1 | tee gpurun_out/c3_sched.log
python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain4_r1h.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mcmc_kernel -s 3 -c 1 -f -o gpurun_out/prof_r1h_c4 python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full4_r1h.log 2>&1
tail -2 gpurun_out/ncu_full4_r1h.log
