set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2g_pytest.log
timeout 900 python bench.py > gpurun_out/r2g_bench_128.json 2> gpurun_out/r2g_bench_128.err; echo "bench rc=$?"
timeout 600 python bench.py --size 256 --reset-every 2 --no-cpu-baseline --no-small > gpurun_out/r2g_bench_256.json 2> gpurun_out/r2g_bench_256.err; echo "256 rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2g_smoke.log 2>&1; echo "smoke rc=$?"
