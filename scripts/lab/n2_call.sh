set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2d_mgpu_test.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2d_mgpu_test.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2d_bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2d_bench_n2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["kernel_classes_ms_per_step"], d["config"].get("divergence_restarts"), d["partition_parity"]["vs_oracle_partitioned"] if d.get("partition_parity") else None)
PY
