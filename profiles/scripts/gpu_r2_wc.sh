#!/bin/bash
# same-box A/B: replicas of the packed weight blobs (ADN_WEIGHT_COPIES)
mkdir -p gpurun_out/wc
for c in "$@"; do
ADN_WEIGHT_COPIES=$c timeout 300 python bench.py --cpu-seconds 0 > gpurun_out/wc/bench_c$c.json 2> gpurun_out/wc/bench_c$c.err
python - $c <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/wc/bench_c{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("copies", sys.argv[1], "fps %.2f ms %.3f" % (d["value"], d["ms_per_step"]), d["stage_ms"], d["clocks"])
PY
done
