cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_person_shard.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -3 > gpurun_out/bench_r2_r.log; tail -1 gpurun_out/bench_r2_r.log > gpurun_out/bench_r2_r.json; python -c "
import json
try:
  d=json.load(open('gpurun_out/bench_r2_r.json'))
  print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'graph', d['extras'].get('cuda_graph'))
except Exception as e:
  print('FAILED', e); print(open('gpurun_out/bench_r2_r.log').read()[-2500:])"
