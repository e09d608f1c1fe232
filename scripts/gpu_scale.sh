# Scaling of the default bench (C3) on one box, launched like the driver does (torch.distributed.run, one rank per GPU):
# weak (1024 ladders per GPU) and strong (1024 ladders in total) at N = 1, 2, 4, 8; every N > 1 line carries the NCCL
# sample-gather leg ("gather").     bash scripts/gpu_scale.sh <tag> "<N list>"
TAG=${1:-r2}; NS=${2:-"2 4 8"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/scale_gpus_$TAG.txt
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu --no-aux --also none > gpurun_out/scale_${TAG}_weak_n1.json 2> gpurun_out/scale_${TAG}_weak_n1.err
for n in $NS; do
  for mode in weak strong; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 3 --warmup 3 --scaling $mode --no-cpu --also none > gpurun_out/scale_${TAG}_${mode}_n$n.json 2> gpurun_out/scale_${TAG}_${mode}_n$n.err
  done
done
for f in gpurun_out/scale_${TAG}_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("scale_")[-1], d["n_gpus"], d["scaling"], "%.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"],
          "e2e %.4g" % d.get("e2e", {}).get("value", float("nan")), "clk", d["clocks"]["sm_mhz"], d["clocks"].get("sm_mhz_per_gpu"),
          "gather", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.get("gather", {}).items() if k != "what"})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
