#!/bin/bash
# c2 tick kernel: warps per chain (B2H_TILE_WPC) and the shared-memory stash (B2H_TILE_STASH)
for cfg in "4 1" "2 1" "2 0"; do
  set -- $cfg
  B2H_TILE_WPC=$1 B2H_TILE_STASH=$2 python bench.py --workload c2 --no-ess --no-cpu --no-secondary --steps 10 > gpurun_out/c2_wpc.json 2> gpurun_out/c2_wpc.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/c2_wpc.json").read().strip().splitlines()[-1])
print("WPC=$1 STASH=$2", "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "tick us", round(d["roofline_elementwise"]["avg_launch_us"], 1))
PY
done
