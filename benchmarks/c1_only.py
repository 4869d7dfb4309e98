"""bench.py's c1 leg alone (HMC L = 10, d = 100, 1 / 65536 chains), three repetitions: run-to-run spread of the one-shot
figure the default bench line carries; B2H_HI_STREAM / B2H_LIB select the build under test."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

D = bench.Dist()
peaks = bench.load_peaks()
for rep in range(3):
    r = bench.secondary_c1(D, peaks)
    print(rep, {k: (round(v["value"]), round(v["ms"], 2)) for k, v in r.items() if isinstance(v, dict) and "ms" in v}, flush=True)
