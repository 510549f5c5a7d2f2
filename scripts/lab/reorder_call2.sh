set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_scale.py tests/test_gpu_linalg.py tests/test_gpu_steady.py tests/test_gpu_golden.py -m gpu -x -q -s > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
for r in 0 1 2; do
  ORC_B200_REORDER=$r python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-small --no-e2e > gpurun_out/r2b_bench_reorder$r.json 2> gpurun_out/r2b_bench_reorder$r.err; echo "bench $r rc=$?"
done
tail -3 gpurun_out/r2b_pytest.log
