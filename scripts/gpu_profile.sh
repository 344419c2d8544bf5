set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# launch list of one warm step (device time per launch; shares matter, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_stdout.log 2>&1
# full capture of the dominant kernel: first shade launch + the following launches of the step
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tc_chain -s 5 -c 8 -o gpurun_out/prof_tc_r1 \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_stdout.log 2>&1
# the real (un-profiled) bench line of the same build
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r1.json
ls -la gpurun_out
