"""The anchoring sub-record of bench.py alone: python tools/anchor_probe.py [reads] [genome_len]."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from blasr_b200 import Aligner  # noqa: E402

al = Aligner(0)
rec = bench.anchoring_record(al, int(sys.argv[1]) if len(sys.argv) > 1 else 1000, int(sys.argv[2]) if len(sys.argv) > 2 else 4_600_000)
print(json.dumps(rec))
al.close()
