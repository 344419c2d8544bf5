set -x
cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
mkdir -p gpurun_out
# launch list of one warm step (device time per launch; shares matter, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench_stdout.log 2>&1
# full capture of the dominant kernel: the launches of one warm step (sdf-only trips, both shade launches, background)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:tc_chain -s 13 -c 13 -o gpurun_out/prof_tc_r2 \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_stdout.log 2>&1
# memcheck of smoke()
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/sanitizer_r2.log 2>&1
tail -5 gpurun_out/sanitizer_r2.log
ls -la gpurun_out
