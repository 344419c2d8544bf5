cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for c in 0 0.8 1.0 1.2; do
  echo "=== MP_TC_RZ_SCALE=$c"
  MP_TC_RZ_SCALE=$c timeout 600 python scripts/gpu_normal_diag.py 2>&1 | grep -v "^surface simt\|^origin simt\|torch fp32\|^simt\|oracle sample"
done
echo "=== trace (discard on)"; timeout 300 python scripts/gpu_trace.py 2>&1 | tail -32
echo "=== trace (discard off)"; MP_TC_KNOBS=6 timeout 300 python scripts/gpu_trace.py 2>&1 | tail -32
for k in 0 4; do
echo "=== bench knobs=$k"; MP_TC_KNOBS=$k timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 > gpurun_out/bench_r2_k$k.json; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_r2_k$k.json')); r=d['roofline']
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'kernel_ms',r['kernel_ms_per_step'],'frac',r['frac'],'serial',r['ms_per_step_single_stream'])"
done
