mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -W ignore -k "esjd or host_buffer or simulation or step_api" > gpurun_out/pytest_esjd.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_esjd.log; tail -8 gpurun_out/pytest_esjd.log
python bench.py --steps 3 --warmup 3 --no-cpu --T 200000 > gpurun_out/bench_aux.json 2> gpurun_out/bench_aux.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_aux.json').read().strip().splitlines()[-1])
print(d['value'], json.dumps(d['aux_kernels'], indent=1))
" || tail gpurun_out/bench_aux.err
