"""CPU: the plain-C restatement (oracle/) against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import golden_io


@pytest.mark.parametrize("name", golden_io.GOLDEN_NAMES + golden_io.DAGP_NAMES)
def test_oracle_matches_reference_golden(oracle, name):
    prm, probs = golden_io.load(name)
    assert len(probs) >= 25
    for i, pb in enumerate(probs):
        o = oracle.forward_wip(prm, pb)
        assert o["score"] == pb["score"], (name, i, pb["tag"])
        assert np.array_equal(o["skl"], pb["skl"]), (name, i, pb["tag"])
        s = oracle.scoreonly_wip(prm, pb)
        assert s["score"] == pb["score_only"], (name, i, pb["tag"])


@pytest.mark.parametrize("name", golden_io.GOLDEN_NAMES + golden_io.DAGP_NAMES + golden_io.UDH_NAMES + golden_io.CIP_NAMES)
def test_oracle_scalar_kernel_matches_reference_golden(oracle, name):
    """Aln2s1::trcbkalignS_ng on its scalar branch (forwardS_ng + Vmf, exact intron scoring):
    the kernel the reference uses for blocks with fewer than 8 query rows"""
    prm, probs = golden_io.load(name)
    assert "penalty" in prm and "sig53tab" in prm and int(prm["intpot"]) == 0
    for i, pb in enumerate(probs):
        o = oracle.trcbk_ng(prm, pb)
        assert o["score"] == pb["ng_score"], (name, i, pb["tag"])
        assert np.array_equal(o["skl"], pb["ng_skl"]), (name, i, pb["tag"])
        # Aln2s1::scorealoneS_ng, the scalar score-only kernel
        assert oracle.scorealone_ng(prm, pb)["score"] == pb["ng_score_only"], (name, i, pb["tag"])


def test_cip_fixtures_exercise_the_bonus(oracle):
    """Cip_score (src/gsinfo.h:127-139): the annotated-query fixtures are only worth something if
    dropping the annotation changes what the exact-ILD kernel and the driver return"""
    prm, probs = golden_io.load("dna_A2_cip")
    n_kernel = n_driver = 0
    for pb in probs:
        bare = dict(pb, cip=None)
        n_kernel += oracle.trcbk_ng(prm, bare)["score"] != pb["ng_score"]
        n_driver += oracle.lsp(prm, bare)["score"] != pb["lsp_score"]
    assert n_kernel >= 20 and n_driver >= 5, (n_kernel, n_driver)
    prm, probs = golden_io.load_protein("prot_A2_cip")
    n_kernel = sum(oracle.trcbk_h_ng(prm, dict(pb, cip=None))["score"] != pb["ng_score"] for pb in probs)
    n_driver = sum(oracle.lsp_h(prm, dict(pb, cip=None))["score"] != pb["lsp_score"] for pb in probs)
    assert n_kernel >= 10 and n_driver >= 4, (n_kernel, n_driver)


def test_golden_covers_edge_cases():
    prm, probs = golden_io.load("dna_A2_global")
    tags = {p["tag"] for p in probs}
    for need in ("gene", "gene_rc", "global_all", "subrange", "tiny1", "tiny16", "tiny17", "rebase"):
        assert need in tags
    # the re-basing case really exceeds the int16 range
    rb = [p for p in probs if p["tag"] == "rebase"][0]
    assert rb["score"] > 32767


EOU = 2 ** 31 - 1 - 2     # end_of_ulk, src/aln.h:49


def cpos_equal(a, b):
    """Dim10 records are only defined up to their end_of_ulk terminator"""
    for ra, rb in zip(a.tolist(), b.tolist()):
        if ra[0] == EOU and rb[0] == EOU and ra[2] == rb[2]:
            continue
        ka = ra.index(EOU) if EOU in ra else 10
        kb = rb.index(EOU) if EOU in rb else 10
        if ra[:ka] != rb[:kb]:
            return False
    return True


@pytest.mark.parametrize("name", golden_io.UDH_NAMES + golden_io.CIP_NAMES)
def test_oracle_udh_matches_reference_golden(oracle, name):
    """hirschbergS1_wip alone and the whole lspS_ng driver under a small -V"""
    prm, probs = golden_io.load(name)
    n_udh = n_lsp = 0
    for i, pb in enumerate(probs):
        rg = pb.get("udh_ranges")
        # Degenerate outputs (alignment "ending" left of its start) come from link lanes the
        # reference never initialises before the first strip (hc_a: src/fwd2s1_simd.h:253-255,
        # only hb_a is cleared at src/fwd2s1_wip_simd.h:524); they are not reproducible.
        if "udh_nim" in pb and rg[0] <= rg[1] and rg[2] <= rg[3]:
            o = oracle.hirschberg_wip(prm, pb, pb["udh_nim"])
            assert o["score"] == pb["udh_score"], (name, i, pb["tag"])
            if o["score"] > -(1 << 28):     # else no path found (NEVSEL): records left as they were
                assert o["ranges"] == pb["udh_ranges"].tolist(), (name, i, pb["tag"])
                assert cpos_equal(o["cpos"], pb["udh_cpos"]), (name, i, pb["tag"])
            n_udh += 1
        o = oracle.lsp(prm, pb)
        assert not o["unsupported"], (name, i, pb["tag"])   # blocks with < 8 rows: scalar kernel
        assert o["score"] == pb["lsp_score"], (name, i, pb["tag"])
        assert np.array_equal(o["skl"], pb["lsp_skl"]), (name, i, pb["tag"])
        n_lsp += 1
    assert n_udh >= 20 and n_lsp == len(probs)


def sudh_cpos_equal(a, b):
    """records of the scalar pass: the crossing list up to its end_of_ulk terminator and, for rows
    that were crossed, the diagonal bounds in [8], [9]"""
    for ra, rb in zip(a.tolist(), b.tolist()):
        ka = ra.index(EOU) if EOU in ra[:8] else 8
        kb = rb.index(EOU) if EOU in rb[:8] else 8
        if ra[:ka] != rb[:kb] or (ka > 0 and ra[8:] != rb[8:]):
            return False
    return True


@pytest.mark.parametrize("name", golden_io.A0_NAMES)
def test_oracle_scalar_mode_matches_reference_golden(oracle, name):
    """`-A0`, the reference's default mode: the scalar kernels, the scalar Hirschberg pass
    hirschbergS_ng (1, 2 and 5 intermediate rows) and the whole driver with blocks banded by the
    recorded diagonal bounds -- global, local and double affine"""
    prm, probs = golden_io.load(name)
    assert int(prm["alg"]) & 3 == 0
    n_pass = n_route = 0
    for i, pb in enumerate(probs):
        o = oracle.trcbk_ng(prm, pb)
        assert o["score"] == pb["ng_score"] and np.array_equal(o["skl"], pb["ng_skl"]), (name, i, pb["tag"])
        for nn in (1, 2, 5):
            if f"sudh{nn}_nim" not in pb:
                continue
            o = oracle.hirschberg_ng(prm, pb, pb[f"sudh{nn}_nim"], pb[f"sudh{nn}_intvl"])
            assert o["score"] == pb[f"sudh{nn}_score"], (name, i, pb["tag"], nn)
            if o["score"] > -(1 << 28):
                assert o["ranges"] == pb[f"sudh{nn}_ranges"].tolist(), (name, i, pb["tag"], nn)
                assert sudh_cpos_equal(o["cpos"], pb[f"sudh{nn}_cpos"]), (name, i, pb["tag"], nn)
            n_pass += 1
        o = oracle.lsp(prm, pb)
        assert not o["unsupported"], (name, i, pb["tag"])
        assert o["score"] == pb["lsp_score"] and np.array_equal(o["skl"], pb["lsp_skl"]), (name, i, pb["tag"])
        m, n = pb["a_right"] - pb["a_left"], pb["b_right"] - pb["b_left"]
        k, q = pb["lw"] - pb["b_left"] + pb["a_right"], pb["b_right"] - pb["a_left"] - pb["up"]
        n_route += 2.0 * (m * n - (k * k + q * q) / 2) >= prm["MaxVmfSpace"]
    assert n_pass >= 50 and n_route >= 12, (n_pass, n_route)
