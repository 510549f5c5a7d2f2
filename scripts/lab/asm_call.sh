mkdir -p gpurun_out
for lab in 4 8 12; do
  ORC_B200_ASM_LAB=$lab timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-small --no-e2e > gpurun_out/r2c_lab$lab.json 2> gpurun_out/r2c_lab$lab.err; echo "lab $lab rc=$?"
done
python - <<'PY'
import json
for r in (4, 8, 12):
    try:
        d=json.loads(open(f"gpurun_out/r2c_lab{r}.json").read().strip().splitlines()[-1])
        print(r, d["ms_per_step"], d["phases_ms_per_step"]["momentum_assembly"])
    except Exception as e: print(r, "ERR", e)
PY
