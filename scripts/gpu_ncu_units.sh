mkdir -p gpurun_out
python bench.py --units 8192 --steps 1 --warmup 3 --no-cpu --no-e2e --T 50000 > gpurun_out/plain_u8192.log 2>&1 && ncu --set full --clock-control none -k regex:mcmc_kernel -s 3 -c 1 -f -o gpurun_out/prof_r1k_c3_u8192 python bench.py --units 8192 --steps 1 --warmup 3 --no-cpu --no-e2e --T 50000 > gpurun_out/ncu_u8192.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_r1k_c3_u8192.ncu-rep > gpurun_out/summary_r1k_c3_u8192.txt 2>&1; rm -f gpurun_out/prof_r1k_c3_u8192.ncu-rep; cat gpurun_out/summary_r1k_c3_u8192.txt
