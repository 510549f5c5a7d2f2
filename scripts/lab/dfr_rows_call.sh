mkdir -p gpurun_out
for r in 4 1; do echo "== ORC_B200_DFR_ROWS=$r"; ORC_B200_DFR_ROWS=$r timeout 200 python scripts/lab/restriction_shapes.py "256x256x32"; done > gpurun_out/r2p_restriction_hexslab.txt 2>&1
cat gpurun_out/r2p_restriction_hexslab.txt
