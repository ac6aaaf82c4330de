#!/bin/bash
# timing experiments with deliberately incomplete kernels (ADN_SH_DIAG builds): what bounds the MMA stream?
mkdir -p gpurun_out
for d in "$@"; do
ADN_LIB_PATH=$PWD/adanerf_b200/libadn_diag$d.so timeout 300 python profiles/trace_sh.py > gpurun_out/diag${d}_trace.txt 2>&1
echo "== DIAG $d"; grep "kernel (CTA 0)\|tile-pair period:\|issuer 0\|producer" gpurun_out/diag${d}_trace.txt
done
