cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mirror.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 > gpurun_out/bench_r2_p.json; python -c "
import json
d=json.load(open('gpurun_out/bench_r2_p.json')); r=d['roofline']
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'e2e ms', round(d['e2e']['ms_per_step'],3), 'launches', d['e2e']['gpu_launches_per_step'], d['parity'])"
