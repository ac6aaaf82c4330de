#!/bin/bash
# small correctness check, bench + trace of the default build, plus the DIAG 1 (empty epilogue) trace
mkdir -p gpurun_out
timeout 120 python profiles/scripts/dbg_small.py 400 || exit 1
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "mlp1 or render_matches or umma" 2>&1 | tail -3
timeout 300 python bench.py --cpu-seconds 0 > gpurun_out/b_bench_sh.json 2> gpurun_out/b_bench_sh.err; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/b_bench_sh.json"))
    print("sh: value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "stage_ms", d["stage_ms"], "clocks", d["clocks"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/b_bench_sh.err").read()[-2000:])
PY
timeout 300 python profiles/trace_sh.py > gpurun_out/b_trace_sh.txt 2>&1
grep "kernel (CTA 0)\|tile-pair period:\|issuer 0\|producer\|epilogue event\|acc seen  " gpurun_out/b_trace_sh.txt | grep -v "^ "
if [ -f adanerf_b200/libadn_diag1.so ]; then
ADN_LIB_PATH=$PWD/adanerf_b200/libadn_diag1.so timeout 300 python profiles/trace_sh.py > gpurun_out/diag1_trace.txt 2>&1
echo "== DIAG 1"; grep "kernel (CTA 0)\|tile-pair period:\|issuer 0: wait\|producer" gpurun_out/diag1_trace.txt
fi
