#!/bin/bash
# GPU check of round 2e (1 GPU): the whole -m gpu suite, the pipeline leg (anchoring + refinement on the device), ncu of the anchoring kernels
TAG=${TAG:-r02e}
mkdir -p gpurun_out/$TAG
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/$TAG/pytest_gpu.log 2>&1; tail -3 gpurun_out/$TAG/pytest_gpu.log
timeout 600 python tools/pipeline_bench.py --reads 2000 > gpurun_out/$TAG/pipeline.json 2> gpurun_out/$TAG/pipeline.err; tail -c 1800 gpurun_out/$TAG/pipeline.json; tail -3 gpurun_out/$TAG/pipeline.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:locate_kernel -c 1 -f -o gpurun_out/$TAG/locate python tools/anchor_probe.py 500 > gpurun_out/$TAG/ncu_locate.log 2>&1; tail -2 gpurun_out/$TAG/ncu_locate.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/$TAG/launches_anchor.csv python tools/anchor_probe.py 500 > /dev/null 2>&1
