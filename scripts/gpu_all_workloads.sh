mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 2500 gpurun_out/bench_c3.json; echo
for wl in c2 c4 c5; do python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$wl.json').read().strip().splitlines()[-1])
print('$wl', d['value'], d['ms_per_step'], d['roofline']['bound'], round(d['roofline']['frac'],4), d['roofline'].get('frac_sfu_measured'), d['roofline'].get('frac_fp32_measured'), d.get('e2e',{}).get('value'))
" || tail -3 gpurun_out/bench_$wl.err; done
