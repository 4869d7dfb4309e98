#!/bin/bash
# c2: split-K slices of the momentum contractions of restarting chains (B2H_RIDER_SPLIT)
for rs in 5 3 4; do
  B2H_RIDER_SPLIT=$rs python bench.py --workload c2 --no-ess --no-cpu --no-secondary > gpurun_out/c2_rs.json 2> gpurun_out/c2_rs.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/c2_rs.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("B2H_RIDER_SPLIT=$rs", "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "gemm ms", round(r["avg_launch_ms"], 4),
      "grad call ms", round(r["gradient_call_in_step_ms"], 4), "tick us", round(d["roofline_elementwise"]["avg_launch_us"], 1))
PY
done
