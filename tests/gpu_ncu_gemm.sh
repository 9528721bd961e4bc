#!/bin/bash
# ncu --set full of ONE GEMM launch (the 4th) with SASS-level stall samples exported as CSV
W=${1:-project}
ncu --set full --clock-control none --import-source on --kernel-name regex:gemm_tc --launch-skip 3 --launch-count 1 \
    -o /tmp/prof_gemm -f timeout 300 python tests/gpu_gemm_one.py $W > gpurun_out/ncu_gemm_$W.log 2>&1
ncu -i /tmp/prof_gemm.ncu-rep --page source --csv > gpurun_out/ncu_gemm_${W}_src.csv 2>/dev/null
ncu -i /tmp/prof_gemm.ncu-rep --page raw --csv > gpurun_out/ncu_gemm_${W}_raw.csv 2>/dev/null
ls -la gpurun_out | tail -4
