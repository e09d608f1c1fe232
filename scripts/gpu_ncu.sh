# ncu capture of the hot kernel for one workload: bash scripts/gpu_ncu.sh <workload> <tag> [extra bench args]
WL=${1:-c3}; TAG=${2:-r1}; shift 2
mkdir -p gpurun_out
python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu --no-e2e --T 2000 "$@" > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mcmc_kernel -s 3 -c 1 -o gpurun_out/prof_$TAG python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu --no-e2e --T 2000 "$@" > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
