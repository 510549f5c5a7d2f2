set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29561 bench.py --gpus 8 --mesh tet --size 75 --steps 3 --warmup 3 --reset-every 2 --no-e2e > gpurun_out/r2l_bench_n8_tet.json 2> gpurun_out/r2l_bench_n8_tet.err; echo "tet rc=$?"
timeout 400 $TR --master-port 29562 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/r2l_bench_n8_weak.json 2> gpurun_out/r2l_bench_n8_weak.err; echo "weak rc=$?"
python - <<'PY'
import json
for f in ("tet","weak"):
    try:
        d=json.loads(open(f"gpurun_out/r2l_bench_n8_{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["global_iters_per_s"], d["ms_per_step"], d["kernel_classes_ms_per_step"], d["config"].get("divergence_restarts"))
    except Exception as e: print(f, "ERR", e)
PY
