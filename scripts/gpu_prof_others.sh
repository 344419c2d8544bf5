cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# ncu --set full of the non-MLP kernels of one warm step (single-stream schedule so the launch order is fixed)
MP_RENDER_STREAMS=0 timeout 1200 ncu --set full --clock-control none --import-source on \
   -k regex:'deform|sampler|composite|bg_points|final_compose|camera_rays|gather_rays' -s 45 -c 45 -o gpurun_out/prof_others_r1 \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_others_stdout.log 2>&1
ls -la gpurun_out/prof_others_r1.ncu-rep
