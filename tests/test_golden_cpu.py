"""CPU: the C restatement (oracle/orc_align.c) against the committed golden vectors (reference outputs)."""
from . import cases, golden_io, oracle as O


def test_oracle_matches_golden():
    recs, fields = golden_io.load()
    assert len(recs) > 150
    for i, r in enumerate(recs):
        p = r["p"]
        fn = O.score_fn(p["M"], p["ins"], p["del_"], p["open"], p["ext"], p["kind"])
        j, keep = O.make_job(p["algo"], p["at"], p["band"], r["q"], r["t"], r["guide"], r["qual"], p["bndIns"], p["bndDel"],
                             p["doStats"], p["statsAffine"])
        got = O.align("orc", fn, j)
        bad = cases.compare(got, r["want"], fields + ["pctSimilarity"])
        assert not bad, f"golden case {i} (algo {p['algo']}, type {p['at']}): {bad}"
