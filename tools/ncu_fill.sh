#!/bin/bash
# ncu of one whole-shard ticket (J pairs): launch list of every kernel, then --set full of the fill kernels with source pages.
TAG=${TAG:-ncu_fill}; J=${J:-20000}; ALGO=${ALGO:-guided}
mkdir -p gpurun_out/$TAG
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/$TAG/launches_${ALGO}_J$J.csv python tools/profile_target.py $J $ALGO > gpurun_out/$TAG/target_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fill_guided -c ${NK:-5} -f -o /tmp/fill_$ALGO python tools/profile_target.py $J $ALGO > gpurun_out/$TAG/target_full.log 2>&1
ncu -i /tmp/fill_$ALGO.ncu-rep --page raw --csv > gpurun_out/$TAG/fill_${ALGO}_raw.csv 2>/dev/null
ncu -i /tmp/fill_$ALGO.ncu-rep --page details > gpurun_out/$TAG/fill_${ALGO}_details.txt 2>/dev/null
ncu -i /tmp/fill_$ALGO.ncu-rep --page source --csv > gpurun_out/$TAG/fill_${ALGO}_source.csv 2>/dev/null
du -sh gpurun_out/$TAG
