#!/bin/bash
# N-GPU weak-scaling A/B of the gather: overlapped all-gather (default), serialised all-gather, gather to rank 0, no collective
N=$1
mkdir -p gpurun_out/j
for m in overlap sync root none overlap; do
ADN_BENCH_GATHER=$m timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --cpu-seconds 0 > gpurun_out/j/bench_n${N}_$m.json 2> gpurun_out/j/bench_n${N}_$m.err || tail -3 gpurun_out/j/bench_n${N}_$m.err
python - $N $m <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/j/bench_n{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1])
print(sys.argv[2], "fps %.1f ms %.3f e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), "by rank", d["ms_per_step_by_rank"], "rank0 stage sum %.3f" % sum(d["stage_ms"].values()))
PY
done
