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
