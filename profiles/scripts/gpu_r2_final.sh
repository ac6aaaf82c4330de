#!/bin/bash
# final verification of the shipped build: GPU test suite, smoke, default bench line, reference arm (one short step), ncu re-capture
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 1500 python -m pytest tests -x -q -m gpu > $O/tests.log 2>&1; tail -3 $O/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 400 python bench.py > $O/bench_800x800_thr0.2_K8.json 2> $O/bench.err; tail -c 1200 $O/bench_800x800_thr0.2_K8.json; echo
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference_arm.json 2> $O/bench_ref.err; tail -c 600 $O/bench_reference_arm.json; echo
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -s 18 -c 60 --csv --log-file $O/launches_r2.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > $O/launches_bench.log 2>&1
timeout 900 $NCU --set full --metrics lts__t_bytes.sum --import-source on -s 12 -c 6 -o $O/prof_default -f python profiles/ncu_frame.py 800x800_thr0.2_K8 2 > $O/prof_default.log 2>&1
ncu -i $O/prof_default.ncu-rep --page raw --csv > $O/raw_default.csv 2>/dev/null
tail -2 $O/prof_default.log
