#!/bin/bash
# One `ncu --set full` capture per kernel (run on the GPU box under gpurun, 1 GPU), exported as raw CSV into gpurun_out/.
#   bash tools/ncu_capture.sh xattn_fwd xattn_bwd sattn_fwd sattn_bwd
# Read here with:  python tools/condense_ncu.py gpurun_out/r2_full_*.csv
set -u
mkdir -p gpurun_out
for k in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s 3 -c 1 -f \
      -o gpurun_out/r2_full_$k python tools/ncu_kernels.py $k > gpurun_out/r2_full_$k.log 2>&1
  ncu -i gpurun_out/r2_full_$k.ncu-rep --page raw --csv > gpurun_out/r2_full_$k.csv 2>/dev/null
  ncu -i gpurun_out/r2_full_$k.ncu-rep --page source --csv > gpurun_out/r2_source_$k.csv 2>/dev/null
  rm -f gpurun_out/r2_full_$k.ncu-rep
done
