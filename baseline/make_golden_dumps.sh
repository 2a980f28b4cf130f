#!/bin/bash
# Regenerates tests/golden/jobs_{c0,c2,c4}_small.bgj.gz: the refinement job sets of the UNMODIFIED reference pipeline on three
# small seeded data sets (configs[0] / [2] / [4] in miniature), dumped by baseline/_ref/blasrmc_dump.  Needs /root/reference
# (run `make -C baseline all` first); the committed .gz files are what the tests on the GPU box read.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
W="${1:-/tmp/bgpu_golden}"
B="$HERE/_ref/blasrmc_dump"
mkdir -p "$W"
python3 "$HERE/make_data.py" c0 "$W/c0" --genome 300000 --reads 40 --len 3000
python3 "$HERE/make_data.py" c2 "$W/c2" --genome 900000 --reads 12 --lo 4000 --hi 8000
python3 "$HERE/make_data.py" c4 "$W/c4" --genome 1000000 --contigs 2 --len 100000
( cd "$W/c0" && BGPU_DUMP=jobs.bgj "$B" reads.fa genome.fa -sam -nproc 1 -out out.sam )
( cd "$W/c2" && BGPU_DUMP=jobs.bgj "$B" reads.fa genome.fa -sam -nproc 1 -bestn 10 -out out.sam )
( cd "$W/c4" && BGPU_DUMP=jobs.bgj "$B" reads.fa genome.fa -sam -nproc 1 -alignContigs -out out.sam )
for c in c0 c2 c4; do gzip -9 -n -c "$W/$c/jobs.bgj" > "$HERE/../tests/golden/jobs_${c}_small.bgj.gz"; done
ls -la "$HERE/../tests/golden/"
