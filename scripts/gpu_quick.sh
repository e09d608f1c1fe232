# quick GPU check: parity tests + C3/C2 bench at a few lanes settings (results in gpurun_out/)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -W ignore > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
for L in ${C3_LANES:-4 8}; do echo "c3 lanes $L"; python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --lanes $L --T 50000 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['value'], d['roofline']['frac'], d['swap_acceptance_rate'], d['esjd'])"; done 2>&1 | tee gpurun_out/sweep_c3.log
for L in ${C2_LANES:-2 4 8 16}; do echo "c2 lanes $L"; python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu --no-e2e --lanes $L --T 50000 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['value'], d['roofline']['frac'], d['acceptance_rate'], d['esjd'])"; done 2>&1 | tee gpurun_out/sweep_c2.log
