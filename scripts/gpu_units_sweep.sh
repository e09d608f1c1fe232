# throughput vs number of ladders / chains per GPU (same kernel): shows how far the named sizes under-fill the machine
mkdir -p gpurun_out
for wl in c3 c2; do for u in ${UNITS:-512 1024 2048 4096 8192}; do echo -n "$wl units=$u: "; python bench.py --workload $wl --units $u --steps 3 --warmup 3 --no-cpu --no-e2e --T 20000 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], round(d['roofline']['frac'],4))"; done; done 2>&1 | tee gpurun_out/units_sweep.log
