set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# what the driver runs at round end: gpu tests, smoke, the default bench line and the reference arm
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json
