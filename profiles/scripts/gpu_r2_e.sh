#!/bin/bash
# round 2, call E: bench lines of every workload at N = 1 (BASELINE configs 2, 3, 4 at N=1, 5 = Pavillon threshold sweep)
mkdir -p gpurun_out/e
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/e/smi.txt 2>&1
run() { # name, extra args
  timeout 400 python bench.py --workload "$1" ${@:2} > gpurun_out/e/bench_$1.json 2> gpurun_out/e/bench_$1.err || { echo "FAILED $1"; tail -5 gpurun_out/e/bench_$1.err; }
  python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/e/bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "fps %.2f e2e %.2f ms %.3f spr %.2f frac %.3f launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["samples_per_ray"], d["roofline"]["frac"], d["gpu_launches"]), d["stage_ms"])
except Exception as e:
    print(sys.argv[1], "no line", e)
PY
}
run 800x800_thr0.2_K8
run 800x800_dense_K128 --cpu-seconds 0 --steps 5
run 800x800_thr0.2_K8_shaped --cpu-seconds 0
run 1600x1600_thr0.2_K8 --cpu-seconds 0
for k in 8 16; do for t in 0.05 0.1 0.2 0.3 0.5; do run 800x800_pav_thr${t}_K${k} --cpu-seconds 0; done; done
