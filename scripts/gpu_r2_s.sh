cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
mkdir -p gpurun_out
for sg in 1 0; do
MP_RENDER_STAGGER=$sg timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 > gpurun_out/bench_r2_s$sg.json; python -c "
import json
d=json.load(open('gpurun_out/bench_r2_s$sg.json'))
print('stagger $sg value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'graph', round(d['extras']['cuda_graph']['ms_per_step'],3))"
done
