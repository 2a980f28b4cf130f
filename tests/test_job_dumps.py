"""CPU: the C restatement against the unmodified reference on the job sets the reference PIPELINE itself issues
(tests/golden/jobs_*_small.bgj.gz: every RefineAlignment / AlignSubstring call of `blasr -sam`, `-bestn 10` on a repeat
genome, and `-alignContigs`, dumped by the instrumented twin baseline/_ref/blasrmc_dump)."""
import numpy as np
import pytest

from blasr_b200 import capi, jobdump
from . import cases, dumps, oracle as O


@pytest.mark.parametrize("cfg", ["c0", "c2", "c4"])
def test_dump_loads_and_oracles_agree(cfg):
    groups = jobdump.load(dumps.DUMPS[cfg])
    assert sum(g.batch.n for g in groups) > 0
    kinds = {g.kind for g in groups}
    assert capi.AFFINE_GUIDED in kinds                      # MakeSane() forces affineAlign
    if cfg == "c4":
        assert capi.AFFINE_KBAND in kinds                   # the gap fills of AlignSubstring (Blasr.cpp:1067)
    if not O.have_ref():
        pytest.skip("reference build absent")
    rng = np.random.default_rng(5)
    for g in groups:
        idx = np.arange(g.batch.n) if g.batch.n <= 400 else np.sort(rng.choice(g.batch.n, 400, replace=False))
        ref = dumps.oracle_group("ref", g, idx)
        orc = dumps.oracle_group("orc", g, idx)
        for k, (a, b) in enumerate(zip(orc, ref)):
            assert not cases.compare(a, b), (cfg, jobdump.KIND_NAMES[g.kind], int(idx[k]), cases.compare(a, b))
