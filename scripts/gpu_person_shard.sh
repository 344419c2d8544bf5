cd $GRAFT_REPO_ROOT
N=${NGPU:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 scripts/person_shard_check.py 2>&1 | grep -v "^W\|warn" | tail -5
