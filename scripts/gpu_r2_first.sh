set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python scripts/gpu_normal_diag.py > gpurun_out/normal_diag.log 2>&1; tail -30 gpurun_out/normal_diag.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
