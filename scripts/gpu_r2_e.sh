cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
echo "=== trace"; timeout 300 python scripts/gpu_trace.py 2>&1 | tail -34 | head -24
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 > gpurun_out/bench_r2_e.json; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_r2_e.json')); r=d['roofline']
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'kernel_ms',r['kernel_ms_per_step'],'frac',r['frac'],'serial',r['ms_per_step_single_stream'], d['parity'])"
