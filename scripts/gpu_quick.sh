cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:---no-cpu-baseline} 2>&1 | tail -1 | tee gpurun_out/bench_tc.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('BENCH',d['engine'],'value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'kernel_ms',round(d['roofline']['kernel_ms_per_step'],3),'ach',round(d['roofline']['achieved'],1),'frac',round(d['roofline']['frac'],4),'issued',round(d['roofline']['issued_tensor_tflops'] or 0,1),d.get('parity'),d.get('cpu_baseline'))"
