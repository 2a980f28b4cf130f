#!/bin/bash
# Quick GPU check under gpurun (1 GPU): parity tests, then the linear / affine / production-shape bench lines without sub-records.
TAG=${TAG:-quick}
mkdir -p gpurun_out/$TAG
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/$TAG/pytest_gpu.log 2>&1; tail -3 gpurun_out/$TAG/pytest_gpu.log
B="--no-subrecords --no-pipeline --steps 5 --warmup 3"
timeout 240 python bench.py $B > gpurun_out/$TAG/bench_lin.json 2> gpurun_out/$TAG/bench_lin.err
timeout 240 python bench.py $B --algo affine > gpurun_out/$TAG/bench_aff.json 2> gpurun_out/$TAG/bench_aff.err
timeout 240 python bench.py $B --algo affine --len-lo 10000 --len-hi 10000 --bands 16 --jobs 40000 > gpurun_out/$TAG/bench_prod.json 2> gpurun_out/$TAG/bench_prod.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/$TAG/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {k: round(v,2) for k,v in d["stage_ms"].items()}, "int_frac %.3f slots/cell %.3f" % (d["int_roofline"]["frac"], d["int_roofline"]["lane_steps_per_cell"]), "ok", d["jobs_ok"], "parity", (d.get("parity_sample") or {}).get("mismatches"))
    except Exception as e:
        print(f, "FAILED", e)
PY
