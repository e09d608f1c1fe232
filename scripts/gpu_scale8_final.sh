mkdir -p gpurun_out
for wl in c3 c2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 8 --workload $wl --steps 5 --warmup 3 > gpurun_out/scale3_${wl}_n8.json 2> gpurun_out/scale3_${wl}_n8.err
python -c "
import json
d=json.loads(open('gpurun_out/scale3_${wl}_n8.json').read().strip().splitlines()[-1])
print('$wl N=8', d['value'], d['ms_per_step'], d['step_ms'], d['e2e']['value'], d['e2e']['calls_ms'])
" || tail -3 gpurun_out/scale3_${wl}_n8.err; done
