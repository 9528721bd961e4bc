#!/bin/bash
# Run on the GPU box (under gpurun): ncu evidence for profiles/.
#  1. every launch of ONE eager training step (MobileNetV2-1.0, N=256) with its device time and
#     DRAM traffic (serialised, cold caches: compare SHARES with bench.py, not absolutes)
#  2. --set full captures of the kernels of single blocks (default b1 b3 b8 b15: the 112x112
#     non-expanding block, a 56x56, a 14x14 and a 7x7 block), exported as CSV
# usage: bash tests/gpu_collect_profiles.sh r02 [blocks...]
set -x
R=${1:-r02}
shift
BLOCKS=${@:-b1 b3 b8 b15}
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${R}_launches_step.csv timeout 600 python tests/gpu_step_once.py \
    > gpurun_out/${R}_step_once.log 2>&1
tail -2 gpurun_out/${R}_step_once.log
for B in $BLOCKS; do
  ncu --set full --import-source on --clock-control none --profile-from-start off -o /tmp/prof_${R}_${B} -f \
      timeout 600 python tests/gpu_profile_block.py $B > gpurun_out/${R}_ncu_${B}.log 2>&1
  ncu -i /tmp/prof_${R}_${B}.ncu-rep --page raw --csv > gpurun_out/${R}_ncu_${B}_raw.csv 2>/dev/null
  if [ "$B" = "b3" ]; then ncu -i /tmp/prof_${R}_${B}.ncu-rep --page source --csv --kernel-name regex:dw_bwd > gpurun_out/${R}_ncu_b3_src_dwbwd.csv 2>/dev/null; ncu -i /tmp/prof_${R}_${B}.ncu-rep --page source --csv --kernel-name regex:dw_fwd > gpurun_out/${R}_ncu_b3_src_dwfwd.csv 2>/dev/null; fi
done
ls -la gpurun_out/ | tail -12
