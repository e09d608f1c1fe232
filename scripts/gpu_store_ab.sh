# A/B: staged float4 stores (default lib) vs direct per-lane stores (librwmpt_direct.so) on store-heavy runs
for v in "" _direct; do lib=$PWD/rwm_pt_pytorch_b200/librwmpt$v.so; [ -f $lib ] || continue
 for cfg in "c2 all 2000" "c3 all 1000" "c3 cold 4000" "c4 all 2000"; do set -- $cfg
  echo -n "lib[$v] $1 store=$2 T=$3: "; RWMPT_LIB=$lib python bench.py --workload $1 --store $2 --T $3 --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], 'GB/s', round(d['value']*d['roofline']['per_chain_step']['stored_bytes']/1e9,1), d['acceptance_rate'])"
 done; done 2>&1 | tee gpurun_out/store_ab.log
