cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
for c in 0 1.0e-6 2.0e-6; do
  echo "=== MP_TC_RZ_COMP=$c"
  MP_TC_RZ_COMP=$c timeout 600 python scripts/gpu_normal_diag.py 2>&1 | grep -v "^surface simt\|^origin simt\|torch fp32\|^simt\|oracle sample"
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_r2_a.json; cat gpurun_out/bench_r2_a.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'kernel_ms',r['kernel_ms_per_step'],'frac',r['frac'],'serial',r['ms_per_step_single_stream'])"
