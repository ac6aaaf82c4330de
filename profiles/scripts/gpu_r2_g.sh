#!/bin/bash
# round 2, call G: BASELINE config 4 (1600x1600, strong scaling over $1 GPUs), torchrun ranks and the single-process C ABI
N=$1
mkdir -p gpurun_out/g
O=gpurun_out/g
WL=1600x1600_thr0.2_K8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $WL --cpu-seconds 0 > $O/bench_${WL}_n$N.json 2> $O/bench_${WL}_n$N.err || tail -5 $O/bench_${WL}_n$N.err
tail -c 900 $O/bench_${WL}_n$N.json; echo
timeout 600 python bench.py --gpus $N --single-process --workload $WL > $O/bench_${WL}_n${N}_single_process.json 2> $O/bench_${WL}_n${N}_single_process.err || tail -5 $O/bench_${WL}_n${N}_single_process.err
tail -c 1200 $O/bench_${WL}_n${N}_single_process.json; echo
if [ "$2" = "weak" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --cpu-seconds 0 > $O/bench_800x800_thr0.2_K8_n$N.json 2> $O/bench_800x800_n$N.err || tail -5 $O/bench_800x800_n$N.err
tail -c 700 $O/bench_800x800_thr0.2_K8_n$N.json; echo
fi
