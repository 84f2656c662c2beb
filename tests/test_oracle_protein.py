"""CPU: protein x genome restatement (oracle/spaln_oracle_h.c) against golden vectors produced
by the unmodified reference's SimdAln2h1::forwardH1_wip."""
import numpy as np
import pytest

import golden_io


@pytest.mark.parametrize("name", golden_io.PROTEIN_NAMES)
def test_oracle_protein_matches_reference_golden(oracle, name):
    prm, probs = golden_io.load_protein(name)
    assert len(probs) >= 20
    for i, pb in enumerate(probs):
        o = oracle.forward_h1_wip(prm, pb)
        assert o["score"] == pb["score"], (name, i, pb["tag"])
        assert np.array_equal(o["skl"], pb["skl"]), (name, i, pb["tag"])
        s = oracle.forward_h1_wip(prm, pb, want_trace=False)
        assert s["score"] == pb["score_only"], (name, i, pb["tag"])
    # exons of the planted genes are recovered (corner lists have several segments)
    assert max(len(pb["skl"]) for pb in probs) >= 6


@pytest.mark.parametrize("name", golden_io.PROTEIN_UDH_NAMES + golden_io.PROTEIN_CIP_NAMES)
def test_oracle_protein_udh_and_driver_match_reference_golden(oracle, name):
    """hirschbergH1_wip (crossing records, narrowed ranges) and the whole driver Aln2h1::lspH_ng
    (trace-back vs Hirschberg dispatch + block re-alignment) against the reference's outputs"""
    prm, probs = golden_io.load_protein(name)
    n_udh = n_lsp = n_unsup = 0
    for i, pb in enumerate(probs):
        if "udh_nim" in pb:
            o = oracle.hirschberg_h1_wip(prm, pb, pb["udh_nim"])
            assert o["score"] == pb["udh_score"], (name, i, pb["tag"])
            assert np.array_equal(o["cpos"][:, :8], pb["udh_cpos"][:, :8]), (name, i, pb["tag"])
            assert o["ranges"] == pb["udh_ranges"].tolist(), (name, i, pb["tag"])
            n_udh += 1
        o = oracle.lsp_h(prm, pb)
        if o["unsupported"]:
            # only a range a Hirschberg pass narrowed to outside the sequences is left unsupported
            # (blocks with < 8 query rows run through the scalar restatement)
            assert pb["a_right"] - pb["a_left"] >= 8, (name, i, pb["tag"])
            n_unsup += 1
            continue
        assert o["score"] == pb["lsp_score"], (name, i, pb["tag"])
        assert np.array_equal(o["skl"], pb["lsp_skl"]), (name, i, pb["tag"])
        n_lsp += 1
    assert n_udh >= 10 and n_lsp >= 15


@pytest.mark.parametrize("name", golden_io.PROTEIN_NAMES + golden_io.PROTEIN_UDH_NAMES + golden_io.PROTEIN_CIP_NAMES)
def test_oracle_scalar_protein_kernel_matches_reference_golden(oracle, name):
    """Aln2h1::trcbkalignH_ng on its scalar branch (forwardH_ng + initH_ng / lastH_ng + Vmf,
    split-codon translation): the kernel the reference uses for blocks with fewer than 8 rows"""
    prm, probs = golden_io.load_protein(name)
    for i, pb in enumerate(probs):
        o = oracle.trcbk_h_ng(prm, pb)
        assert o["score"] == pb["ng_score"], (name, i, pb["tag"])
        assert np.array_equal(o["skl"], pb["ng_skl"]), (name, i, pb["tag"])


EOU = 2 ** 31 - 1 - 2


def _sudh_cpos_equal(a, b):
    for ra, rb in zip(a.tolist(), b.tolist()):
        ka = ra.index(EOU) if EOU in ra[:8] else 8
        kb = rb.index(EOU) if EOU in rb[:8] else 8
        if ra[:ka] != rb[:kb] or (ka > 0 and ra[8:] != rb[8:]):
            return False
    return True


@pytest.mark.parametrize("name", golden_io.PROTEIN_A0_NAMES)
def test_oracle_protein_scalar_mode_matches_reference_golden(oracle, name):
    """`-A0` for protein queries: forwardH_ng, the scalar Hirschberg pass hirschbergH_ng (1, 2 and 5
    intermediate rows) and the whole driver lspH_ng with blocks banded by the recorded diagonal bounds"""
    prm, probs = golden_io.load_protein(name)
    assert int(prm["alg"]) & 3 == 0
    n_pass = n_route = 0
    for i, pb in enumerate(probs):
        o = oracle.trcbk_h_ng(prm, pb)
        assert o["score"] == pb["ng_score"] and np.array_equal(o["skl"], pb["ng_skl"]), (name, i, pb["tag"])
        for nn in (1, 2, 5):
            if f"sudh{nn}_nim" not in pb:
                continue
            o = oracle.hirschberg_h_ng(prm, pb, pb[f"sudh{nn}_nim"], pb[f"sudh{nn}_intvl"])
            assert o["score"] == pb[f"sudh{nn}_score"], (name, i, pb["tag"], nn)
            if o["score"] > -(1 << 28):
                assert o["ranges"] == pb[f"sudh{nn}_ranges"].tolist(), (name, i, pb["tag"], nn)
                assert _sudh_cpos_equal(o["cpos"], pb[f"sudh{nn}_cpos"]), (name, i, pb["tag"], nn)
            n_pass += 1
        o = oracle.lsp_h(prm, pb)
        assert not o["unsupported"], (name, i, pb["tag"])
        assert o["score"] == pb["lsp_score"] and np.array_equal(o["skl"], pb["lsp_skl"]), (name, i, pb["tag"])
        m, n = pb["a_right"] - pb["a_left"], pb["b_right"] - pb["b_left"]
        k, q = pb["lw"] - pb["b_left"] + 3 * pb["a_right"], pb["b_right"] - 3 * pb["a_left"] - pb["up"]
        n_route += 2.0 * (m * n - (k * k + q * q) / 6) >= prm["MaxVmfSpace"]
    assert n_pass >= 40 and n_route >= 6, (n_pass, n_route)
