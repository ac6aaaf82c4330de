#!/bin/bash
# shading-kernel change: correctness (parity tests), same-box bench against the previous build (libadn_prev.so), timeline
mkdir -p gpurun_out/h
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_parity_gate.py -x -q -m gpu > gpurun_out/h/tests.log 2>&1; tail -3 gpurun_out/h/tests.log
for v in new old new old; do
  if [ $v = old ]; then export ADN_LIB_PATH=$PWD/adanerf_b200/libadn_prev.so; else unset ADN_LIB_PATH; fi
  timeout 300 python bench.py --cpu-seconds 0 > gpurun_out/h/bench_$v.json 2> gpurun_out/h/bench_$v.err
  python - $v <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/h/bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], "fps %.2f ms %.3f" % (d["value"], d["ms_per_step"]), d["stage_ms"], d["clocks"]["sm_mhz"])
PY
done
unset ADN_LIB_PATH
timeout 300 python profiles/trace_sh.py > gpurun_out/h/trace.txt 2>&1; grep "kernel (CTA 0)\|tile-pair period:\|issuer 0\|producer" gpurun_out/h/trace.txt
