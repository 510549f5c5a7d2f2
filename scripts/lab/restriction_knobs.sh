#!/bin/bash
# sweep of the restriction kernel's lab knobs on the bench workload: device time of the "restriction" class per SIMPLE iteration
for B in 8 4 2; do for S in 0 100 500; do
  ORC_B200_DFR_BLOCKS=$B ORC_B200_DFR_SLEEP=$S python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-small 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('blocks/SM $B sleep_ns $S: restriction %.2f ms/step, step %.1f ms' % (d['kernel_classes_ms_per_step']['restriction'], d['ms_per_step']))"
done; done
