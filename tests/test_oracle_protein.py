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
