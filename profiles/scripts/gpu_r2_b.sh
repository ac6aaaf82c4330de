#!/bin/bash
# quick loop: bench of the default build (+ optional old-kernel A/B) and the mlp_sh_kernel timeline
mkdir -p gpurun_out
timeout 300 python bench.py --cpu-seconds 0 > gpurun_out/b_bench_sh.json 2> gpurun_out/b_bench_sh.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/b_bench_sh.json"))
    print("sh: value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "stage_ms", d["stage_ms"], "clocks", d["clocks"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/b_bench_sh.err").read()[-2000:])
PY
if [ "$1" = "ab" ]; then
ADN_SHADING_KERNEL=0 timeout 300 python bench.py --cpu-seconds 0 > gpurun_out/b_bench_old.json 2> gpurun_out/b_bench_old.err
python -c "
import json; d=json.load(open('gpurun_out/b_bench_old.json')); print('old: value', round(d['value'],1), 'stage_ms', d['stage_ms'])"
fi
timeout 300 python profiles/trace_sh.py > gpurun_out/b_trace_sh.txt 2>&1; grep -v "^  *[0-9]* *\(mma\|epi\)" gpurun_out/b_trace_sh.txt | tail -150
