cd $GRAFT_REPO_ROOT
run() { env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$*','ms',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['kernel_ms_per_step'],3))"; }
run MP_TC_V2=0
run MP_TC_V2=1
run MP_TC_V2=1 MP_TC_VARIANT=1
run MP_TC_V2=1 MP_TC_VARIANT=2
run MP_TC_V2=1 MP_TC_VARIANT=3
