cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in 0 0.7e-6 1.4e-6 2.1e-6; do
  echo "=== MP_TC_RZ_COMP=$c"
  MP_TC_RZ_COMP=$c timeout 600 python scripts/gpu_normal_diag.py 2>&1 | grep -v "^surface simt\|^origin simt\|torch fp32\|^simt\|oracle sample"
done
