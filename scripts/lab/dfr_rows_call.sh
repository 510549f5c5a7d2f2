mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_linalg.py tests/test_gpu_scale.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2k_pytest.log
{ echo "== tuned ticket shape (default)"; timeout 900 python scripts/lab/restriction_shapes.py; } > gpurun_out/r2k_restriction_tuned.txt 2>&1
cat gpurun_out/r2k_restriction_tuned.txt
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-small --no-e2e > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --mesh tet --size 48 --steps 4 --warmup 3 --reset-every 2 --no-e2e > gpurun_out/r2k_tet48.json 2> gpurun_out/r2k_tet48.err; echo "tet48 rc=$?"
timeout 600 python bench.py --mesh tet --size 80 --steps 4 --warmup 3 --reset-every 2 --no-e2e > gpurun_out/r2k_tet80.json 2> gpurun_out/r2k_tet80.err; echo "tet80 rc=$?"
python - <<'PY'
import json
for f in ("bench","tet48","tet80"):
    try:
        d=json.loads(open(f"gpurun_out/r2k_{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["kernel_classes_ms_per_step"], d["config"]["divergence_restarts"])
    except Exception as e: print(f, "ERR", e)
PY
