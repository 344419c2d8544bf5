cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r2_full.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_full.json')); r=d['roofline']
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'kernel_ms',round(r['kernel_ms_per_step'],3),'frac',round(r['frac'],4))
print('parity', d['parity']); print('cpu', d.get('cpu_baseline'))
for k,v in d['extras'].items(): print(k, v)
PY
for prec in colour1 throughput; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --precision $prec 2>&1 | tail -1 > gpurun_out/bench_r2_$prec.json; python -c "
import json
d=json.load(open('gpurun_out/bench_r2_$prec.json')); r=d['roofline']
print('$prec value',round(d['value']),'kernel_ms',round(r['kernel_ms_per_step'],3),'frac',round(r['frac'],4))"; done
