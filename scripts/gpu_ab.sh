cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# A/B of kernel experiment switches: one short bench per MP_TC_KNOBS value given in $KNOBS (space separated)
for k in ${KNOBS:-0 1}; do
  MP_TC_KNOBS=$k timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('KNOBS $k value',round(d['value']),'ms',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['kernel_ms_per_step'],3))"
done
if [ -n "$AB_TESTS" ]; then timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3; fi
