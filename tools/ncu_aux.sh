#!/bin/bash
# ncu --set full of the helper kernels (prep / trace / emit) of one 20k-pair ticket, source pages exported as CSV.
TAG=${TAG:-ncu_aux}; J=${J:-20000}; ALGO=${ALGO:-guided}
mkdir -p gpurun_out/$TAG
for K in ${KERNELS:-prep_guided trace_guided emit_kernel}; do
  ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -f -o /tmp/$K python tools/profile_target.py $J $ALGO > gpurun_out/$TAG/$K.log 2>&1
  ncu -i /tmp/$K.ncu-rep --page source --csv > gpurun_out/$TAG/${K}_source.csv 2>/dev/null
  ncu -i /tmp/$K.ncu-rep --page raw --csv > gpurun_out/$TAG/${K}_raw.csv 2>/dev/null
  ncu -i /tmp/$K.ncu-rep --page details > gpurun_out/$TAG/${K}_details.txt 2>/dev/null
done
du -sh gpurun_out/$TAG
