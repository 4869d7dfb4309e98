#!/bin/bash
# c2 step with / without the high-priority stream hop (B2H_HI_STREAM) of dense-metric runs
for hs in 0 1; do
  B2H_HI_STREAM=$hs python bench.py --workload c2 --no-ess --no-cpu --no-secondary > gpurun_out/c2_ab_$hs.json 2> gpurun_out/c2_ab_$hs.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/c2_ab_$hs.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("B2H_HI_STREAM=$hs", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3),
      "gemm ms", round(r["avg_launch_ms"], 4), "grad call ms", round(r["gradient_call_in_step_ms"], 4),
      "tick us", round(d["roofline_elementwise"]["avg_launch_us"], 1))
PY
done
