#!/bin/bash
# round 2e, second check (1 GPU): rescore + anchoring tests, anchoring probe with the 8-byte index entries (+ ncu), headline bench with the pipelined resident leg
TAG=${TAG:-r02e}
mkdir -p gpurun_out/$TAG
timeout 600 python -m pytest tests/test_gpu_rescore.py tests/test_gpu_anchor.py tests/test_gpu_adapter.py -x -q 2>&1 | tail -4
python tools/anchor_probe.py > gpurun_out/$TAG/anchor2.json 2> gpurun_out/$TAG/anchor2.err
python -c "
import json; d=json.load(open('gpurun_out/$TAG/anchor2.json')); print(d['value'], d['device_ms'], d['e2e']['value'], d['e2e']['ms_per_call'], d['parity'], d['cpu_baseline']['value'])"; tail -2 gpurun_out/$TAG/anchor2.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:locate_kernel -c 1 -f -o gpurun_out/$TAG/locate2 python tools/anchor_probe.py 500 > gpurun_out/$TAG/ncu_locate2.log 2>&1; tail -1 gpurun_out/$TAG/ncu_locate2.log
timeout 500 python bench.py --no-subrecords --no-pipeline --steps 5 --warmup 3 > gpurun_out/$TAG/bench_lin_pipelined.json 2> gpurun_out/$TAG/bench_lin_pipelined.err
python -c "
import json; d=json.loads(open('gpurun_out/$TAG/bench_lin_pipelined.json').read().strip().splitlines()[-1]); print('value', d['value'], 'resident', d['resident'], 'e2e', d['e2e']['value'], d['stage_ms'])"; tail -3 gpurun_out/$TAG/bench_lin_pipelined.err
