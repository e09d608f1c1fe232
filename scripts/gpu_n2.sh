mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/n1_check.json 2> gpurun_out/n1_check.err; tail -c 1200 gpurun_out/n1_check.json; echo
for i in 1 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29510 + i)) bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/n2_check$i.json 2> gpurun_out/n2_check$i.err
python -c "
import json
d=json.loads(open('gpurun_out/n2_check$i.json').read().strip().splitlines()[-1])
print('N=2', d['value'], d['ms_per_step'], d['step_ms'], d['e2e']['value'], d['e2e']['calls_ms'], d['clocks'], d.get('all_ranks'))
" || tail -5 gpurun_out/n2_check$i.err; done
python bench.py --workload c5f --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_c5f.json 2> gpurun_out/bench_c5f.err; tail -c 900 gpurun_out/bench_c5f.json
