cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 > gpurun_out/bench_r2_n$N.log
tail -1 gpurun_out/bench_r2_n$N.log > gpurun_out/bench_r2_n$N.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_r2_n$N.json'))
    print('N', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))
    print('parity', d['parity'])
    for k,v in d['extras'].items(): print(k, v)
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/bench_r2_n$N.log').read()[-3000:])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 1 --warmup 0 2>&1 | tail -1 | cut -c1-300
