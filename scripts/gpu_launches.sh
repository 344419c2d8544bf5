cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tmp.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_stdout.log 2>&1
python - <<'PY'
import csv, collections
rows=[]
lines=[l for l in open('gpurun_out/launches_tmp.csv') if not l.startswith('==')]
for r in csv.DictReader(lines):
    if r.get('Metric Name')=='gpu__time_duration.sum':
        rows.append((int(r['ID']), r['Kernel Name'].split('(')[0], float(r['Metric Value'].replace(',',''))))
idx=[i for i,r in enumerate(rows) if 'camera_rays' in r[1]]
s,e=idx[1],idx[2]
agg=collections.OrderedDict(); tot=0
for r in rows[s:e]:
    agg.setdefault(r[1],[0,0.0]); agg[r[1]][0]+=1; agg[r[1]][1]+=r[2]; tot+=r[2]
print("launches",e-s,"total ms",tot/1e6)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:8]: print("%-50s n=%3d %9.1f us %5.1f%%"%(k[:50],v[0],v[1]/1e3,100*v[1]/tot))
print([ (r[1][-22:], round(r[2]/1e3,1)) for r in rows[s:e] if 'deform' in r[1]])
PY
