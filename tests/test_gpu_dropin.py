"""GPU: true drop-in check.  The reference process (oracle/_ref) builds its own Seq / PwdB /
Exinon objects from FASTA text; the same objects go (a) through the reference's
SimdAln2s1::forwardS1_wip on the CPU and (b) through include/gspaln_spaln_adapter.hpp ->
libgspaln -> CUDA.  Scores and Mfile corner lists must be identical."""
import numpy as np
import pytest

import ref_harness

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_harness.available(), reason="oracle/_ref not built")]


def test_adapter_dropin_matches_reference():
    from spaln_b200 import workload
    ref = ref_harness.Reference._instance or ref_harness.Reference("-Q0 -A2 -S1 -yX0 -TDictyost")
    p = ref.params()
    rng = np.random.default_rng(99)
    for i in range(10):
        g, q, _ = workload.plant_gene(rng, qlen_range=(60, 900), flank=(50, 600))
        t = ref.task(g, q, comrev_query=(i % 5 == 4))
        if i % 3 == 2:
            t.set(a_exgl=0, a_exgr=0, b_exgl=0, b_exgr=0)
        lw, up = t.stripe(p["sh"])
        r = t.kernel(lw, up, 0)
        a = t.adapter(lw, up, 0)
        assert a["score"] == r["score"], i
        assert np.array_equal(a["skl"], r["skl"]), i
        assert t.adapter(lw, up, 1)["score"] == t.kernel(lw, up, 1)["score"], i
        t.close()
