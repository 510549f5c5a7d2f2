set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_linalg.py tests/test_gpu_scale.py -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2f_pytest.log
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-small"
timeout 600 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-small > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 4300 --csv --log-file gpurun_out/r2f_launches.csv $B > gpurun_out/r2f_ncu1.log 2>&1; echo "ncu1 rc=$?"
python scripts/summarize_launches.py gpurun_out/r2f_launches.csv > gpurun_out/r2f_launches_summary.txt 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_spmv_vec -s 400 -c 4 -o /tmp/spmv_vec_full -f python scripts/one_iter.py 128 1 > gpurun_out/r2f_ncu2.log 2>&1; echo "ncu2 rc=$?"
python scripts/ncu_summary.py /tmp/spmv_vec_full.ncu-rep > gpurun_out/r2f_spmv_vec_ncu_full.txt 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_scale_permute_rows|k_strongest_dataflow" -s 2 -c 4 -o /tmp/setup_full -f python scripts/one_iter.py 128 1 > gpurun_out/r2f_ncu3.log 2>&1; echo "ncu3 rc=$?"
python scripts/ncu_summary.py /tmp/setup_full.ncu-rep > gpurun_out/r2f_setup_ncu_full.txt 2>&1
rm -f gpurun_out/r2f_launches.csv.tmp; gzip -f gpurun_out/r2f_launches.csv
du -sh gpurun_out
