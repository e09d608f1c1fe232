mkdir -p gpurun_out
for v in ${VARIANTS:-main}; do [ "$v" = main ] && v=""; lib=$PWD/rwm_pt_pytorch_b200/librwmpt$v.so; [ -f $lib ] || continue
echo "=== variant [$v]"; RWMPT_LIB=$lib timeout 600 python -m pytest tests/test_gpu_parity.py -q -W ignore -k "increments_are_iid or reference_statistics" 2>&1 | tail -25
done 2>&1 | tee gpurun_out/inc_test.log
