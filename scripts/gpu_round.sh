# One GPU session: packed-fp32 probe, parity tests, then C3 / C2 A/B over the variant libs that exist.
#   VARIANTS="'' _nox2 _exp2" bash scripts/gpu_round.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
[ -x scripts/probes/f32x2_probe ] && timeout 300 scripts/probes/f32x2_probe > gpurun_out/f32x2_probe.log 2>&1
if [ -z "$SKIP_TESTS" ]; then
timeout 1200 python -m pytest tests -m gpu -x -q -W ignore > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
fi
for v in ${VARIANTS:-main}; do [ "$v" = main ] && v=""; lib=$PWD/rwm_pt_pytorch_b200/librwmpt$v.so; [ -f $lib ] || continue
 for wl in ${WORKLOADS:-c3 c2}; do echo -n "variant[$v] $wl: "; RWMPT_LIB=$lib timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-e2e --T ${T:-100000} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], round(d['roofline']['frac'],4), d['acceptance_rate'], d.get('swap_acceptance_rate'), d['esjd'])"; done; done 2>&1 | tee gpurun_out/ab_round.log
# schedule A/B on the main library: plain (1) vs auto (0)
for sc in ${SCHEDULES:-}; do for wl in ${WORKLOADS:-c3 c2}; do echo -n "schedule[$sc] $wl: "; RWMPT_SCHEDULE=$sc timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-e2e --T ${T:-100000} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], round(d['roofline']['frac'],4), d['acceptance_rate'], d.get('swap_acceptance_rate'), d['esjd'])"; done; done 2>&1 | tee gpurun_out/ab_sched.log
