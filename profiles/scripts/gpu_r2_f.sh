#!/bin/bash
# round 2, call F: evidence -- launch list of the bench command, ncu --set full (+ L2 bytes) of one chunk of the default, dense and
# ragged Pavillon workloads, compute-sanitizer memcheck / synccheck / racecheck logs.
mkdir -p gpurun_out/f
O=gpurun_out/f
NCU="ncu --clock-control none"
# 1. launch list of the bench command (2 timed steps after 3 warm-ups: 18 launches skipped, then every launch)
timeout 600 $NCU --metrics gpu__time_duration.sum -s 18 -c 60 --csv --log-file $O/launches_r2.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > $O/launches_bench.log 2>&1
# 2. --set full of one chunk per workload (skip = warm frames x launches per frame)
cap() { # workload, launches per frame, tag
  timeout 900 $NCU --set full --metrics lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum --import-source on -s $((2 * $2)) -c 6 -o $O/prof_$3 -f python profiles/ncu_frame.py $1 2 > $O/prof_$3.log 2>&1
  ncu -i $O/prof_$3.ncu-rep --page raw --csv > $O/raw_$3.csv 2>/dev/null
  tail -3 $O/prof_$3.log
}
cap 800x800_thr0.2_K8 6 default
cap 800x800_dense_K128 60 dense
cap 800x800_pav_thr0.5_K16 12 pav_thr0.5_K16
# 3. sanitizers on the mixed small workload; the library variant has the device watchdog widened (tools slow the kernels ~100x)
export ADN_LIB_PATH=$PWD/adanerf_b200/libadanerf_b200_san.so
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python profiles/sanitizer_smoke.py > $O/sanitizer_$tool.log 2>&1; echo "rc=$?" >> $O/sanitizer_$tool.log
  tail -4 $O/sanitizer_$tool.log
done
# racecheck: the SIMT stages (shared-memory hazards are meaningful there) ...
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 30 --kernel-regex kns=stage python profiles/sanitizer_smoke.py > $O/sanitizer_racecheck_stages.log 2>&1; echo "rc=$?" >> $O/sanitizer_racecheck_stages.log
tail -4 $O/sanitizer_racecheck_stages.log
# ... and the MLP kernels (async-proxy / mbarrier ordering is not modelled by the tool: hazards it prints are listed in DESIGN.md)
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 30 --kernel-regex kns=mlp python profiles/sanitizer_smoke.py > $O/sanitizer_racecheck_mlp.log 2>&1; echo "rc=$?" >> $O/sanitizer_racecheck_mlp.log
tail -6 $O/sanitizer_racecheck_mlp.log
ls -la $O | head -40
