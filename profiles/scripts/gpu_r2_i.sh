#!/bin/bash
# NDC path with the fused tile writer, dense parity rows, NDC bench line + ncu capture
mkdir -p gpurun_out/i
O=gpurun_out/i
timeout 900 python -m pytest tests/test_parity_gate.py::test_dense_rows_against_oracle tests/test_gpu_parity.py -x -q -m gpu -k "dense or ndc or render_matches" -s > $O/tests.log 2>&1; grep -i "passed\|failed\|error\|dense K=128\|ndc:" $O/tests.log | tail -8
timeout 300 python bench.py --workload 800x800_ndc_thr0.15_K16 --cpu-seconds 0 > $O/bench_800x800_ndc_thr0.15_K16.json 2> $O/bench_ndc.err || tail -5 $O/bench_ndc.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/i/bench_800x800_ndc_thr0.15_K16.json").read().strip().splitlines()[-1])
print("ndc fps %.2f e2e %.2f ms %.3f spr %.2f launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["samples_per_ray"], d["gpu_launches"]), d["stage_ms"])
PY
timeout 900 ncu --clock-control none --set full --metrics lts__t_bytes.sum --import-source on -s 24 -c 6 -o $O/prof_ndc -f python profiles/ncu_frame.py 800x800_ndc_thr0.15_K16 2 > $O/prof_ndc.log 2>&1
ncu -i $O/prof_ndc.ncu-rep --page raw --csv > $O/raw_ndc.csv 2>/dev/null
tail -3 $O/prof_ndc.log
