python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02c_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
tail -c 600 gpurun_out/r02c_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02c_bench.json").read().strip().splitlines()[-1])
print("c5", d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["issued_frac"], d["roofline_elementwise"]["frac"], d["clocks"])
for k, v in d["secondary"].items():
    print(k, json.dumps(v)[:700])
PY
