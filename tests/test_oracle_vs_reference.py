"""CPU: oracle restatement against the live reference shim (oracle/_ref), when it
has been built in this checkout (it needs /root/reference at build time)."""
import numpy as np
import pytest

import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="oracle/_ref not built")


def test_oracle_vs_live_reference(oracle):
    from spaln_b200 import workload
    ref = ref_harness.Reference("-Q0 -A2 -S1 -yX0 -TDictyost")
    p = ref.params()
    assert p["nelem"] == 16, "the oracle is pinned to the AVX2 (16-lane) build"
    rng = np.random.default_rng(2025)
    n = 0
    for i in range(12):
        g, q, _ = workload.plant_gene(rng, qlen_range=(40, 400), flank=(40, 250))
        for kw in ({}, {"a_exgl": 0, "a_exgr": 0, "b_exgl": 0, "b_exgr": 0}):
            t = ref.task(g, q, comrev_query=(i % 4 == 3))
            if kw:
                t.set(**kw)
            lw, up = t.stripe(p["sh"])
            ex = t.export()
            ex.update(lw=lw, up=up)
            r = t.kernel(lw, up, 0)
            o = oracle.forward_wip(p, ex)
            assert r["score"] == o["score"]
            assert np.array_equal(r["skl"], o["skl"])
            assert t.kernel(lw, up, 1)["score"] == oracle.scoreonly_wip(p, ex)["score"]
            # our stripe() restatement equals the reference's
            assert workload.stripe(ex["a_left"], ex["a_right"], ex["b_left"], ex["b_right"], p["sh"]) == (lw, up)
            t.close()
            n += 1
    assert n == 24
