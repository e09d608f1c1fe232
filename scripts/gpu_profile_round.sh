# Round profile: launch list of the default bench command + full ncu capture of the hot kernel (C3 and C2).
TAG=${1:-r1}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_list_$TAG.log 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain2_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mcmc_kernel -s 3 -c 1 -o gpurun_out/prof_${TAG}_c3 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1
python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu --no-e2e --T 100000 > gpurun_out/plain3_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mcmc_kernel -s 3 -c 1 -o gpurun_out/prof_${TAG}_c2 python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu --no-e2e --T 100000 > gpurun_out/ncu_full2_$TAG.log 2>&1
ls -la gpurun_out | tail -8
