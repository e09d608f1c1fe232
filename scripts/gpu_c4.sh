# C4 tuned-kernel check: full GPU parity suite, then C4 bench at two run lengths (+ C3 / C2 / C5 regression lines).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -W ignore > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
for spec in "c4 2000" "c4 20000" "c3 100000" "c2 100000" "c5 20000"; do set -- $spec; echo -n "$1 T=$2: "; timeout 600 python bench.py --workload $1 --steps 3 --warmup 3 --no-cpu --no-e2e --T $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['bound'], round(d['roofline']['frac'],4), d['acceptance_rate'], d.get('swap_acceptance_rate'), d['esjd'])"; done 2>&1 | tee gpurun_out/c4_ab.log
