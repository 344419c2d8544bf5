cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:deform -s 12 -c 14 -o gpurun_out/prof_deform \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_stdout.log 2>&1
ls -la gpurun_out/prof_deform.ncu-rep
