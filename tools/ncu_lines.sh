#!/bin/bash
# ncu --set full of ONE kernel (K, regex) of a J-pair ticket; exports the source page with CUDA-line correlation.
TAG=${TAG:-ncu_lines}; J=${J:-20000}; ALGO=${ALGO:-guided}; K=${K:-prep_guided}
mkdir -p gpurun_out/$TAG
ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -f -o /tmp/$K python tools/profile_target.py $J $ALGO > gpurun_out/$TAG/$K.log 2>&1
ncu -i /tmp/$K.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/$TAG/${K}_cuda_sass.csv 2>/dev/null
ncu -i /tmp/$K.ncu-rep --page source --csv --print-source cuda > gpurun_out/$TAG/${K}_cuda.csv 2>/dev/null
ls -la gpurun_out/$TAG
