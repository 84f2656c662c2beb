"""CPU: the plain-C restatement (oracle/) against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import golden_io


@pytest.mark.parametrize("name", golden_io.GOLDEN_NAMES)
def test_oracle_matches_reference_golden(oracle, name):
    prm, probs = golden_io.load(name)
    assert len(probs) >= 25
    for i, pb in enumerate(probs):
        o = oracle.forward_wip(prm, pb)
        assert o["score"] == pb["score"], (name, i, pb["tag"])
        assert np.array_equal(o["skl"], pb["skl"]), (name, i, pb["tag"])
        s = oracle.scoreonly_wip(prm, pb)
        assert s["score"] == pb["score_only"], (name, i, pb["tag"])


def test_golden_covers_edge_cases():
    prm, probs = golden_io.load("dna_A2_global")
    tags = {p["tag"] for p in probs}
    for need in ("gene", "gene_rc", "global_all", "subrange", "tiny1", "tiny16", "tiny17", "rebase"):
        assert need in tags
    # the re-basing case really exceeds the int16 range
    rb = [p for p in probs if p["tag"] == "rebase"][0]
    assert rb["score"] > 32767
