# ncu --set full capture of one launch of a kernel (after the same command has run clean without the profiler):
#   bash scripts/gpu_ncu.sh <tag> <kernel regex> <skip> -- <command ...>
TAG=$1; KRE=$2; SKIP=$3; shift 4
mkdir -p gpurun_out
"$@" > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 1 -f -o gpurun_out/prof_$TAG "$@" > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
