#!/bin/bash
# last check of the shipped build: GPU test suite, smoke, default bench line
mkdir -p gpurun_out/final2
O=gpurun_out/final2
timeout 1500 python -m pytest tests -x -q -m gpu > $O/tests.log 2>&1; tail -2 $O/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 400 python bench.py > $O/bench_800x800_thr0.2_K8.json 2> $O/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/final2/bench_800x800_thr0.2_K8.json").read().strip().splitlines()[-1])
print("fps %.2f e2e %.2f ms %.3f frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"]), d["stage_ms"], d["clocks"], d["cpu_baseline"]["value"])
PY
