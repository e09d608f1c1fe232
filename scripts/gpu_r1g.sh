# Round-1 final evidence for the current kernel: GPU tests, full bench line, launch list, ncu --set full of C3 and C4.
TAG=${1:-r1g}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -W ignore > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 3000 gpurun_out/bench_c3.json; echo
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_list_$TAG.log 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain2_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mcmc_kernel -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_c3 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1
python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain4_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mcmc_kernel -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_c4 python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full4_$TAG.log 2>&1
python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain5_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mcmc_kernel -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_c5 python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full5_$TAG.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json
ls -la gpurun_out | tail -8
