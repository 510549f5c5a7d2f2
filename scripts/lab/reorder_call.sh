set -x
mkdir -p gpurun_out
python scripts/lab/dump_levels.py 128 /tmp --morton > gpurun_out/r2_reorder_dump.log 2>&1
for l in 1 2 3; do for v in "" _morton; do ./scripts/lab/reorder_lab /tmp/lvl${l}${v}.bin 40; done; done > gpurun_out/r2_reorder_lab.txt 2>&1
tail -5 gpurun_out/r2_reorder_dump.log
