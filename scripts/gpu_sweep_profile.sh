mkdir -p gpurun_out
for L in 2 4 8 16 32; do echo "c3 lanes $L"; python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --lanes $L --T 50000 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['value'], d['roofline']['frac'])"; done > gpurun_out/sweep_c3.log 2>&1
for L in 1 2 4 8 16 32; do echo "c2 lanes $L"; python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu --no-e2e --lanes $L --T 50000 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['value'], d['roofline']['frac'])"; done > gpurun_out/sweep_c2.log 2>&1
cat gpurun_out/sweep_c3.log gpurun_out/sweep_c2.log
python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --T 2000 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mcmc_kernel -s 3 -c 1 -o gpurun_out/prof_r1a python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --T 2000 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
python bench.py --steps 2 --warmup 3 --no-cpu --T 2000 > gpurun_out/plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 2 --warmup 3 --no-cpu --T 2000 > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log; ls -la gpurun_out
