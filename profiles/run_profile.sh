#!/bin/bash
# Profiling recipe of /opt/skills/guides/B200_PROFILING.md (run under gpurun, 1 GPU).
# Numbers printed by runs under ncu are never bench values; only the ncu outputs are kept.
# The .ncu-rep files stay in /tmp on the box (gpurun_out/ is capped at 64 MiB); their raw / source pages are exported as CSV.
set -x
mkdir -p gpurun_out
J=${J:-20000}
for ALGO in ${ALGOS:-guided affine}; do
  # launch list (cold-cache, serialised per-launch times) of one whole-shard ticket: submit+collect, then one rerun
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${ALGO}.csv \
      python tools/profile_target.py $J $ALGO > gpurun_out/target_under_ncu_${ALGO}.log 2>&1
  # full sections of every kernel of the same command
  ncu --set full --clock-control none -c 60 -f -o /tmp/full_${ALGO} \
      python tools/profile_target.py $J $ALGO > gpurun_out/ncu_full_${ALGO}.log 2>&1
  ncu -i /tmp/full_${ALGO}.ncu-rep --page raw --csv > gpurun_out/ncu_full_${ALGO}_raw.csv 2>/dev/null
  # source-level counters of the widest-used fill kernel
  ncu --set full --clock-control none --import-source on -k regex:fill_guided -s 5 -c 3 -f -o /tmp/fillsrc_${ALGO} \
      python tools/profile_target.py $J $ALGO > gpurun_out/ncu_fillsrc_${ALGO}.log 2>&1
  ncu -i /tmp/fillsrc_${ALGO}.ncu-rep --page source --csv > gpurun_out/ncu_fillsrc_${ALGO}_source.csv 2>/dev/null
  ls -la /tmp/*.ncu-rep
done
du -sh gpurun_out
