cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
mkdir -p gpurun_out
cp multiply_b200/libmultiply_b200.so multiply_b200/_variants/lib_main.so
bash scripts/gpu_r2_l.sh main vQ
export MP_LIB=$GRAFT_REPO_ROOT/multiply_b200/_variants/lib_vQ.so
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python scripts/gpu_normal_diag.py 2>&1 | grep "^surface tc\|^origin tc\|^tc rendered"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:tc_chain -s 13 -c 13 --csv --log-file gpurun_out/dram_vQ.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.DictReader([l for l in open('gpurun_out/dram_vQ.csv') if not l.startswith('==')])]
tot=0
for r in rows:
    if r['Metric Name'].startswith('dram__bytes'):
        v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
        tot+=v*{'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}[u]
print('vQ DRAM bytes over the 13 tc launches of one step: %.3f GB' % (tot/1e9))
PY
