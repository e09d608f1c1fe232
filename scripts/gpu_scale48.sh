# Re-measure N = 4 and 8 (and N = 1 for the same box) of the default bench with the asynchronous accumulator reduction.
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/scale2_n1.json 2> gpurun_out/scale2_n1.err
for n in 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale2_n$n.json 2> gpurun_out/scale2_n$n.err
done
for n in 1 4 8; do python -c "
import json
d=json.loads(open('gpurun_out/scale2_n$n.json').read().strip().splitlines()[-1])
print($n, d['value'], d['ms_per_step'], d['step_ms'], d['e2e']['value'], d['e2e']['calls_ms'])
" || tail -3 gpurun_out/scale2_n$n.err; done
