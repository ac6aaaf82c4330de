#!/bin/bash
# 2 GPUs: bench with the gather-to-rank-0 default (weak + strong), multi-GPU tests
N=2
mkdir -p gpurun_out/k
for wl in 800x800_thr0.2_K8 1600x1600_thr0.2_K8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload $wl --cpu-seconds 0 > gpurun_out/k/bench_${wl}_n$N.json 2> gpurun_out/k/bench_${wl}_n$N.err || tail -5 gpurun_out/k/bench_${wl}_n$N.err
python - $wl <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/k/bench_{sys.argv[1]}_n2.json").read().strip().splitlines()[-1])
print(sys.argv[1], "fps %.1f ms %.3f e2e %.1f finite %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["finite"]), d["ms_per_step_by_rank"], d["config"]["parallelism"])
PY
done
timeout 600 python -m pytest tests/test_multi.py tests/test_viewer_host.py tests/test_host_api.py -x -q -m gpu 2>&1 | tail -3
