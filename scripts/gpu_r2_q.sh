cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -3 > gpurun_out/bench_r2_q.log; tail -1 gpurun_out/bench_r2_q.log > gpurun_out/bench_r2_q.json; python -c "
import json
try:
  d=json.load(open('gpurun_out/bench_r2_q.json'))
  print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']), 'graph', d['extras'].get('cuda_graph'))
except Exception as e:
  print('FAILED', e); print(open('gpurun_out/bench_r2_q.log').read()[-2500:])"
