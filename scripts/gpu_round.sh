# One GPU session: parity tests, the default bench line (all BASELINE configs + reference baselines), the reference arm,
# and the ncu launch list of the same command.
#   [SKIP_TESTS=1] [BENCH_ARGS="--quick"] bash scripts/gpu_round.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu_$TAG.txt
if [ -z "$SKIP_TESTS" ]; then
  timeout 2400 python -m pytest tests -m gpu -x -q -W ignore -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
  grep -E "^\[|passed|failed|error|pytest exit" gpurun_out/pytest_gpu_$TAG.log | tail -40
fi
timeout 1500 python bench.py --steps 5 --warmup 3 $BENCH_ARGS > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; tail -c 1500 gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
    print("C3", d["value"], d["roofline"]["frac"], "e2e", d.get("e2e", {}).get("value"), "clocks", d["clocks"])
    print("cpu_baseline", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("kind"), "torch", {k: v.get("chain_steps_per_s") for k, v in d.get("reference_torch", {}).items() if isinstance(v, dict)})
    for k, v in d.get("also", {}).items():
        if "value" in v:
            print(k, v["value"], v["roofline"]["bound"], round(v["roofline"]["frac"], 4), "e2e", v.get("e2e", {}).get("value"), v.get("note"))
        else:
            print(k, json.dumps(v.get("ladders")))
    print("aux", {k: (v.get("flat", {}).get("frac_of_hbm_peak"), v.get("with_moved_count", {}).get("frac_of_hbm_peak")) if "flat" in v else v.get("frac_of_hbm_peak") for k, v in d.get("aux_kernels", {}).items()})
except Exception as e:
    print("bench parse failed", e)
PY
if [ -z "$SKIP_REF" ]; then
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 700 gpurun_out/bench_ref_$TAG.json
fi
if [ -z "$SKIP_NCU" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-aux --also none > gpurun_out/ncu_list_$TAG.log 2>&1; tail -3 gpurun_out/launches_$TAG.csv
fi
