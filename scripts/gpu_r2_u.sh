cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 > gpurun_out/bench_r2_u.json; python -c "
import json
d=json.load(open('gpurun_out/bench_r2_u.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'serial', round(d['roofline']['ms_per_step_single_stream'],3), 'kernel', round(d['roofline']['kernel_ms_per_step'],3))"
