mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 6 --warmup 3 > gpurun_out/r2h_bench_n4.json 2> gpurun_out/r2h_bench_n4.err; echo "n4 rc=$?"
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2h_mgpu_test.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2h_mgpu_test.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2h_bench_n4.json").read().strip().splitlines()[-1])
print(d["value"], d["global_iters_per_s"], d["ms_per_step"], d["kernel_classes_ms_per_step"], d["config"].get("divergence_restarts"), d["config"]["workload"])
PY
