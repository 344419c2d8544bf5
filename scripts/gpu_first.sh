set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q -k "not tc" 2>&1 | tail -30
MP_ENGINE=simt timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
