#!/bin/bash
# round 2, call A: correctness of mlp_sh_kernel, same-box A/B against the round-1 shading kernel, timeline
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/a_tests.log
tail -5 gpurun_out/a_tests.log
timeout 300 python bench.py --cpu-seconds 0 > gpurun_out/a_bench_sh.json 2> gpurun_out/a_bench_sh.err; tail -c 1500 gpurun_out/a_bench_sh.json
ADN_SHADING_KERNEL=0 timeout 300 python bench.py --cpu-seconds 0 > gpurun_out/a_bench_old.json 2> gpurun_out/a_bench_old.err; tail -c 600 gpurun_out/a_bench_old.json
timeout 300 python bench.py --cpu-seconds 0 > gpurun_out/a_bench_sh2.json 2> gpurun_out/a_bench_sh2.err
timeout 300 python profiles/trace_sh.py > gpurun_out/a_trace_sh.txt 2>&1; tail -80 gpurun_out/a_trace_sh.txt
