cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "$@"; do
  L=$GRAFT_REPO_ROOT/multiply_b200/_variants/lib_$v.so
  MP_LIB=$L timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 > gpurun_out/bench_var_${v}.json
  python -c "
import json,sys
try:
  d=json.load(open('gpurun_out/bench_var_${v}.json')); r=d['roofline']
  print('$v', 'value',round(d['value']),'kernel_ms',round(r['kernel_ms_per_step'],3),'frac',round(r['frac'],4),'serial',round(r['ms_per_step_single_stream'],3), 'e2e', round(d['e2e']['value']))
except Exception as e: print('$v', 'FAILED', e, open('gpurun_out/bench_var_${v}.json').read()[-300:])"
  echo "--- trace $v"; MP_LIB=$L timeout 300 python scripts/gpu_trace.py 2>&1 | tail -34 | sed -n '4,6p;12,14p;19,21p'
done
