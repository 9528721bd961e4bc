#!/bin/bash
# Run on the GPU box (under gpurun): ncu evidence for profiles/.
#  1. every launch of ONE eager training step (MobileNetV2-1.0, N=256) with its device time and
#     DRAM traffic (serialised, cold caches: compare SHARES with bench.py, not absolutes)
#  2. a --set full capture of the kernels of one block (b3: 24->144->24 at 56x56), exported as CSV
set -x
R=${1:-r01}
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${R}_launches_step.csv timeout 600 python tests/gpu_step_once.py \
    > gpurun_out/${R}_step_once.log 2>&1
tail -2 gpurun_out/${R}_step_once.log
ncu --set full --clock-control none --profile-from-start off -o /tmp/prof_${R} -f \
    timeout 600 python tests/gpu_profile_block.py b3 > gpurun_out/${R}_ncu_b3.log 2>&1
ncu -i /tmp/prof_${R}.ncu-rep --page raw --csv > gpurun_out/${R}_ncu_b3_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -8
ncu -i /tmp/prof_${R}.ncu-rep --page source --csv --kernel-name regex:dw_ \
    > gpurun_out/${R}_ncu_b3_src_dw.csv 2>/dev/null
ls -la gpurun_out/ | tail -4
