# ncu --set full capture of the hot kernel: bash scripts/gpu_ncu2.sh <workload> <tag> <T> [env...]
WL=${1:-c3}; TAG=${2:-r1}; T=${3:-20000}
mkdir -p gpurun_out
python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu --no-e2e --T $T > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mcmc_kernel -s 3 -c 1 -f -o gpurun_out/prof_$TAG python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu --no-e2e --T $T > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
