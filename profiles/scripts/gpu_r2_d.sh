#!/bin/bash
# same-box bench of several builds of the library: default first, then each ADN_LIB_PATH variant named on the command line
mkdir -p gpurun_out
run() { # name libpath
  ADN_LIB_PATH=$2 timeout 300 python bench.py --cpu-seconds 0 > gpurun_out/d_$1.json 2> gpurun_out/d_$1.err
  python - "$1" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/d_{n}.json"))
    print(f"{n:10s} value {d['value']:.1f} e2e {d['e2e']['value']:.1f} mlp1 {d['stage_ms']['mlp1']:.3f} ms  mlp0 {d['stage_ms']['mlp0']:.3f}  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(n, "FAILED", e); print(open(f"gpurun_out/d_{n}.err").read()[-1500:])
PY
}
run default ""
for v in "$@"; do run $v $PWD/adanerf_b200/libadn_$v.so; done
run default2 ""
