set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/r2d_bench_n8_weak.json 2> gpurun_out/r2d_bench_n8_weak.err; echo "weak rc=$?"
timeout 600 $TR --master-port 29542 bench.py --gpus 8 --steps 6 --warmup 3 --scaling strong > gpurun_out/r2d_bench_n8_strong.json 2> gpurun_out/r2d_bench_n8_strong.err; echo "strong rc=$?"
timeout 900 $TR --master-port 29543 bench.py --gpus 8 --mesh tet --size 75 --steps 2 --warmup 2 --reset-every 2 > gpurun_out/r2d_bench_n8_tet.json 2> gpurun_out/r2d_bench_n8_tet.err; echo "tet rc=$?"
python - <<'PY'
import json
for f in ("weak","strong","tet"):
    try:
        d=json.loads(open(f"gpurun_out/r2d_bench_n8_{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["kernel_classes_ms_per_step"], d["config"].get("divergence_restarts"))
    except Exception as e: print(f, "ERR", e)
PY
