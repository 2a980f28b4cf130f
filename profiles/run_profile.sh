#!/bin/bash
# Profiling recipe of /opt/skills/guides/B200_PROFILING.md (run under gpurun, 1 GPU).
# Numbers printed by runs under ncu are never bench values; only the ncu outputs are kept.
set -x
mkdir -p gpurun_out
J=${J:-20000}
for ALGO in ${ALGOS:-guided affine}; do
  # launch list (cold-cache, serialised per-launch times) of one whole-shard ticket: submit+collect, then one rerun
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${ALGO}.csv \
      python tools/profile_target.py $J $ALGO > gpurun_out/target_under_ncu_${ALGO}.log 2>&1
  # full sections of every kernel of the same command
  ncu --set full --clock-control none --import-source on -c 60 -f -o gpurun_out/full_${ALGO} \
      python tools/profile_target.py $J $ALGO > gpurun_out/ncu_full_${ALGO}.log 2>&1
done
ls -la gpurun_out
