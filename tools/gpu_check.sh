#!/bin/bash
# GPU check of HEAD under gpurun (1 GPU): parity tests, then both bench lines.  TAG names the output directory.
TAG=${TAG:-check}
mkdir -p gpurun_out/$TAG
python -m pytest tests -x -q -m gpu > gpurun_out/$TAG/pytest_gpu.log 2>&1; tail -3 gpurun_out/$TAG/pytest_gpu.log
python bench.py > gpurun_out/$TAG/bench_guided_100k.json 2> gpurun_out/$TAG/bench_guided.err
if [ "${AFFINE:-1}" = "1" ]; then python bench.py --algo affine > gpurun_out/$TAG/bench_affine_100k.json 2> gpurun_out/$TAG/bench_affine.err; fi
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/$TAG/bench_*_100k.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {k: round(v,2) for k,v in d["stage_ms"].items()}, "int_frac %.3f" % d["int_roofline"]["frac"], "ok", d["jobs_ok"])
    except Exception as e:
        print(f, "FAILED", e)
PY
