#!/bin/bash
# c2: potential of the Gaussian target formed by the tick kernel (default) or by its own kernel (B2H_TICK_POTENTIAL=0)
for tp in 0 1; do
  B2H_TICK_POTENTIAL=$tp python bench.py --workload c2 --no-ess --no-cpu --no-secondary > gpurun_out/c2_tp.json 2> gpurun_out/c2_tp.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/c2_tp.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("B2H_TICK_POTENTIAL=$tp", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3), "gemm ms", round(r["avg_launch_ms"], 4),
      "grad call ms", round(r["gradient_call_in_step_ms"], 4), "tick us", round(d["roofline_elementwise"]["avg_launch_us"], 1), "accept", round(d["config"]["mean_accept"], 4))
PY
done
