mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r2o_pytest.log
timeout 400 python bench.py > gpurun_out/r2o_bench_128.json 2> gpurun_out/r2o_bench_128.err; echo "bench rc=$?"
timeout 300 python bench.py --mesh tet --size 48 --steps 4 --warmup 3 --reset-every 2 --no-e2e > gpurun_out/r2o_tet48.json 2> gpurun_out/r2o_tet48.err; echo "tet48 rc=$?"
python - <<'PY'
import json
for f in ("bench_128","tet48"):
    d=json.loads(open(f"gpurun_out/r2o_{f}.json").read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["kernel_classes_ms_per_step"]["restriction"], d["phases_ms_per_step"]["momentum_solves"], d["roofline"]["frac"])
PY
