# Round-end evidence on the final code: GPU tests, smoke, full bench line (C3), reference arm, launch list, ncu --set full of the
# C3 / C2 / C4 hot kernels and of the ESJD reduction kernel.
TAG=${1:-r1j}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -W ignore > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 600 gpurun_out/bench_c3.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/bench_ref.json; echo
for wl in c2 c4 c5 c5f; do python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; done
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_list_$TAG.log 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain2_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mcmc_kernel -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_c3 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1
python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu --no-e2e --T 200000 > gpurun_out/plain3_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mcmc_kernel -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_c2 python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu --no-e2e --T 200000 > gpurun_out/ncu_full2_$TAG.log 2>&1
python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain4_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mcmc_kernel -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_c4 python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full4_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:esjd_flat -s 4 -c 1 -f -o gpurun_out/prof_${TAG}_esjd python bench.py --steps 1 --warmup 3 --no-cpu --T 20000 > gpurun_out/ncu_esjd_$TAG.log 2>&1
# the merge back is capped at 64 MiB: summarise on the box, keep only the C3 report itself
for w in c3 c2 c4 esjd; do [ -f gpurun_out/prof_${TAG}_$w.ncu-rep ] && python scripts/ncu_summary.py gpurun_out/prof_${TAG}_$w.ncu-rep > gpurun_out/summary_${TAG}_$w.txt 2>&1; done
rm -f gpurun_out/prof_${TAG}_c2.ncu-rep gpurun_out/prof_${TAG}_c4.ncu-rep
ls -la gpurun_out | grep $TAG; du -sh gpurun_out
