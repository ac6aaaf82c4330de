#!/bin/bash
mkdir -p gpurun_out
for d in "$@"; do
ADN_LIB_PATH=$PWD/adanerf_b200/libadn_t2_$d.so timeout 300 python profiles/trace_sh.py > gpurun_out/t2_${d}_trace.txt 2>&1
echo "== T2 $d"; grep "kernel (CTA 0)\|tile-pair period:" gpurun_out/t2_${d}_trace.txt
python profiles/trace_sh2.py | tee gpurun_out/t2_${d}_segments.txt
cp gpurun_out/trace_sh_raw.npy gpurun_out/t2_${d}_raw.npy
done
