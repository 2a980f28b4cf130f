#!/bin/bash
# Profiling recipe of /opt/skills/guides/B200_PROFILING.md applied to bench.py (run under gpurun, 1 GPU).
# Numbers printed by these runs are never bench values; only the ncu outputs are kept.
set -x
mkdir -p gpurun_out
J=${J:-30000}
ALGO=${ALGO:-guided}
KREGEX=${KREGEX:-fill_guided}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${ALGO}.csv \
    python bench.py --jobs $J --steps 2 --warmup 1 --algo $ALGO --cpu-seconds 0.5 > gpurun_out/bench_under_ncu_${ALGO}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s ${SKIP:-4} -c ${COUNT:-4} -f -o gpurun_out/fill_${ALGO} \
    python bench.py --jobs $J --steps 1 --warmup 1 --algo $ALGO --cpu-seconds 0.5 > gpurun_out/ncu_full_${ALGO}.log 2>&1
ls -la gpurun_out
