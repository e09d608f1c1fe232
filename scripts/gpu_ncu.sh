# ncu --set full capture of one launch of a kernel (after the same command has run clean without the profiler); the text
# summary (scripts/ncu_summary.py) and the per-opcode instruction mix are written on the box, the .ncu-rep is dropped
# (gpurun_out/ is limited to 64 MiB):
#   bash scripts/gpu_ncu.sh <tag> <kernel regex> <skip> -- <command ...>
TAG=$1; KRE=$2; SKIP=$3; shift 4
mkdir -p gpurun_out
"$@" > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 1 -f -o /tmp/prof_$TAG "$@" > gpurun_out/ncu_$TAG.log 2>&1
python scripts/ncu_summary.py /tmp/prof_$TAG.ncu-rep > gpurun_out/ncu_summary_$TAG.txt 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv 2>/dev/null | python scripts/ncu_opcode_mix.py >> gpurun_out/ncu_summary_$TAG.txt 2>&1
head -12 gpurun_out/ncu_summary_$TAG.txt | cut -c1-160
rm -f /tmp/prof_$TAG.ncu-rep
