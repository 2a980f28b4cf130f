"""Loader of tests/golden/refinement_golden.npz (made by tests/golden/make_golden.py from the reference itself)."""
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refinement_golden.npz")
PKEYS = ("algo", "at", "band", "bndIns", "bndDel", "statsAffine", "doStats", "ins", "del_", "open", "ext", "kind")


def load():
    z = np.load(PATH)
    fields = [str(x) for x in z["fields"]]
    out = []
    for i in range(int(z["n"])):
        p = dict(zip(PKEYS, (int(x) for x in z[f"p{i}"])))
        p["M"] = z[f"M{i}"]
        want = dict(zip(fields, (int(x) for x in z[f"r{i}"])))
        want["pctSimilarity"] = np.float32(z[f"s{i}"])
        want["blocks"] = z[f"b{i}"]
        gaps, k = [], 0
        a = z[f"a{i}"]
        for c in z[f"c{i}"]:
            gaps.append([(int(x), int(y)) for x, y in a[k:k + int(c)]]); k += int(c)
        want["gaps"] = gaps
        qv = z[f"v{i}"]
        g = z[f"g{i}"]
        out.append(dict(p=p, q=z[f"q{i}"], t=z[f"t{i}"], guide=g if len(g) else None, qual=qv if len(qv) else None, want=want))
    return out, fields
