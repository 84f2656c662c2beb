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


def test_adapter_dropin_driver_matches_reference():
    """the whole driver: the reference's Aln2s1::lspS_ng on the CPU against SpalnEngine::lspS_ng
    (gspaln_lsp on the GPU) on the reference's own objects -- trace-back problems, Hirschberg
    route (the set-up keeps the default -V, so the larger problems take it) and tiny problems that
    the reference aligns with its scalar kernel"""
    from spaln_b200 import workload
    ref = ref_harness.Reference._instance or ref_harness.Reference("-Q0 -A2 -S1 -yX0 -TDictyost")
    p = ref.params()
    rng = np.random.default_rng(123)
    shapes = [((60, 600), (50, 600))] * 5 + [((2, 7), (30, 300))] * 3 + [((1800, 2600), (1500, 4000))] * 2
    n_udh = 0
    for i, (qr, fl) in enumerate(shapes):
        g, q, _ = workload.plant_gene(rng, qlen_range=qr, flank=fl)
        t = ref.task(g, q, comrev_query=(i % 5 == 4))
        lw, up = t.stripe(p["sh"])
        r = t.lsp(lw, up, cap=1 << 17)
        a = t.adapter_lsp(lw, up, cap=1 << 17)
        assert a is not None, i
        assert a["score"] == r["score"], (i, a["score"], r["score"])
        assert np.array_equal(a["skl"], r["skl"]), i
        n_udh += 2.0 * len(q) * (len(g) + len(q)) >= p["MaxVmfSpace"]
        t.close()
    assert n_udh >= 2


def test_adapter_dropin_driver_carries_cip_score():
    """a query annotated with intron positions (SigII -> Cip_score, src/gsinfo.h:127-139): the
    adapter hands the bonus table to the device; few-row problems around an exon junction (what
    block re-alignment produces) reach the exact-ILD kernel, where the bonus is read"""
    from spaln_b200 import workload
    ref = ref_harness.Reference._instance or ref_harness.Reference("-Q0 -A2 -S1 -yX0 -TDictyost")
    p = ref.params()
    rng = np.random.default_rng(321)
    n_matter = 0
    for i in range(12):
        g, q, tr = workload.plant_gene(rng, qlen_range=(120, 300), n_exons=3, flank=(40, 120), sub=0, indel=0)
        j = i % 2
        J = sum(e - s for s, e in tr[: j + 1])
        lo, hi = int(rng.integers(1, 5)), int(rng.integers(1, 4))
        t = ref.task(g, q)
        t.set(a_left=J - lo, a_right=J + hi, b_left=tr[j][1] - lo, b_right=tr[j + 1][0] + hi,
              a_exgl=0, a_exgr=0, b_exgl=0, b_exgr=0)
        lw, up = t.stripe(p["sh"])
        bare = t.lsp(lw, up)
        pos = sorted({J, J + int(rng.integers(-2, 3)), int(rng.integers(1, len(q)))})
        t.set_cip(pos, rng.integers(1, 4, size=len(pos)))
        r = t.lsp(lw, up)
        a = t.adapter_lsp(lw, up)
        assert a is not None, i
        assert a["score"] == r["score"] and np.array_equal(a["skl"], r["skl"]), (i, a["score"], r["score"])
        n_matter += r["score"] != bare["score"]
        t.close()
    assert n_matter >= 6, n_matter
