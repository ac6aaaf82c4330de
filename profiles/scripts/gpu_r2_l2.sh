#!/bin/bash
mkdir -p gpurun_out/l
for v in 0 1 0 1 0 1; do
  ADN_SH_REFETCH=$v timeout 300 python bench.py --cpu-seconds 0 --steps 60 > gpurun_out/l/bench2_$v.json 2> gpurun_out/l/bench2_$v.err
  python - $v <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/l/bench2_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("refetch", sys.argv[1], "fps %.2f ms %.3f mlp1 %.4f" % (d["value"], d["ms_per_step"], d["stage_ms"]["mlp1"]), d["clocks"]["sm_mhz"])
PY
done
