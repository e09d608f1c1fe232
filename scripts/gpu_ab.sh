# A/B test variant builds: RWMPT_LIB selects the library (see rwm_pt_pytorch_b200/_lib.py)
mkdir -p gpurun_out
for v in "" _o1 _o2 _o3 _o4 _o5 _o6; do
  lib=rwm_pt_pytorch_b200/librwmpt$v.so
  [ -f $lib ] || continue
  for wl in c3 c2; do
    echo -n "variant[$v] $wl: "
    RWMPT_LIB=$PWD/$lib python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-e2e --lanes ${LANES:-4} --T 50000 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['value'], d['roofline']['frac'], d['acceptance_rate'])"
  done
done 2>&1 | tee gpurun_out/ab.log
