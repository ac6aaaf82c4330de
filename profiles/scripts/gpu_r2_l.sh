#!/bin/bash
# A/B: tile inputs of the shading kernel fetched once per layer (default) vs once per N half (ADN_SH_REFETCH=1)
mkdir -p gpurun_out/l
ADN_SH_REFETCH=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mlp1 or render_matches or band" 2>&1 | tail -2
for v in 0 1 0 1; do
  ADN_SH_REFETCH=$v timeout 300 python bench.py --cpu-seconds 0 > gpurun_out/l/bench_$v.json 2> gpurun_out/l/bench_$v.err
  python - $v <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/l/bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("refetch", sys.argv[1], "fps %.2f ms %.3f mlp1 %.4f" % (d["value"], d["ms_per_step"], d["stage_ms"]["mlp1"]), d["clocks"]["sm_mhz"])
PY
done
