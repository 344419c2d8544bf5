cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python scripts/gpu_tc_debug.py 2>&1 | grep -E "^tc|first|err" | head
echo "--- tests"
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('BENCH v2',d['engine'],'value',round(d['value']),'ms',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['kernel_ms_per_step'],3),'frac',round(d['roofline']['frac'],4),'issued',round(d['roofline']['issued_tensor_tflops'] or 0,1))"
MP_TC_V2=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('BENCH v1',d['engine'],'value',round(d['value']),'ms',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['kernel_ms_per_step'],3),'frac',round(d['roofline']['frac'],4),'issued',round(d['roofline']['issued_tensor_tflops'] or 0,1))"
