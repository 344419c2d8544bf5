cd $GRAFT_REPO_ROOT
# A/B over one environment variable: VAR=name VALS="a b c"
for v in $VALS; do
  env $VAR=$v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$VAR=$v value',round(d['value']),'ms',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['kernel_ms_per_step'],3))"
done
