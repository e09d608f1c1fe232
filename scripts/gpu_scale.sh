# Weak-scaling run of the default bench (C3, 1024 ladders per GPU) at N = 1, 2, 4, 8 on one box, launched like the driver does.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/scale_gpus.txt
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
for n in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
done
for n in 1 2 4 8; do python -c "
import json
d=json.loads(open('gpurun_out/scale_n$n.json').read().strip().splitlines()[-1])
print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('all_ranks'))
" || tail -3 gpurun_out/scale_n$n.err; done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 8 --workload c5 --steps 3 --warmup 3 > gpurun_out/scale_c5_n8.json 2> gpurun_out/scale_c5_n8.err; tail -c 600 gpurun_out/scale_c5_n8.json
