# A/B of variant libs on C3 at 1024 and 8192 ladders: VARIANTS="'' _exp2" bash scripts/gpu_ab2.sh
mkdir -p gpurun_out
for v in ${VARIANTS:-"" _exp2}; do lib=$PWD/rwm_pt_pytorch_b200/librwmpt$v.so; [ -f $lib ] || continue
 for u in 1024 8192; do echo -n "variant[$v] c3 units=$u: "; RWMPT_LIB=$lib python bench.py --units $u --steps 3 --warmup 3 --no-cpu --no-e2e --T 20000 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], round(d['roofline']['frac'],4), d['acceptance_rate'], d['swap_acceptance_rate'])"; done; done 2>&1 | tee gpurun_out/ab2.log
