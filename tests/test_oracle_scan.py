"""CPU: the splice-signal scan restatement (oracle/spaln_oracle_scan.c: Exinon::intron53_c /
intron53_n + PatMat::calcPatMat) against the Exinon tables of the unmodified reference that the
golden fixtures carry (sig5, sig3, int53 of every fixture problem)."""
import numpy as np
import pytest

import golden_io


def scan_equal(got, pb):
    """got: dict sig5 / sig3 / int53 by column; pb: fixture problem.  sig3[0], sig5[len - 1] and the
    INT53 halves the reference never writes are not reproducible (uninitialised memory there)."""
    L = len(pb["b"]) - 2
    ok = np.array_equal(got["sig5"][:L - 1], pb["sig5"][:L - 1]) and np.array_equal(got["sig3"][1:L], pb["sig3"][1:L])
    ok = ok and np.array_equal(got["int53"][0:L - 1] & 0x0f0f, pb["int53"][0:L - 1] & 0x0f0f)
    return ok and np.array_equal(got["int53"][1:L + 1] & 0xf0f0, pb["int53"][1:L + 1] & 0xf0f0)


@pytest.mark.parametrize("name", ["dna_A2_global", "dna_A2_tetrapod", "dna_A2_udh"])
def test_oracle_scan_matches_reference_tables(oracle, name):
    prm, probs = golden_io.load(name)
    assert int(prm["pat5_meta"][4]) == 2 and int(prm["pat3_meta"][4]) == 2     # Markov order 2 PSSMs
    for i, pb in enumerate(probs):
        got = oracle.exinon_scan(prm, pb["b"][1:-1])
        assert scan_equal(got, pb), (name, i, pb["tag"])


def test_oracle_scan_handles_ambiguity_and_short_segments(oracle):
    prm, _ = golden_io.load("dna_A2_global")
    from spaln_b200 import workload
    for s in ("A", "ACG", "ACGTNNACGT" * 3, "GTAAGT" + "N" * 30 + "TTTCAG"):
        got = oracle.exinon_scan(prm, workload.encode_dna(s))
        assert np.array_equal(got["int53"], workload.synthetic_int53(workload.encode_dna(s)))
        assert got["sig5"][len(s)] == 0 and got["sig5"][len(s) + 1] == 0


def test_oracle_nuc2tron_matches_reference(oracle):
    """Seq::nuc2tron: residue codes as the reference reads the DNA (DNA set-up) against the tron
    codes the protein set-up holds for the same segments (tests/golden/make_golden_nuc2tron.py),
    all fifteen IUPAC codes included"""
    z = np.load(golden_io.GOLDEN_DIR / "nuc2tron.npz")
    assert int(z["n"]) >= 10
    seen = set()
    for i in range(int(z["n"])):
        d, t = z[f"dna{i}"], z[f"tron{i}"]
        assert np.array_equal(oracle.nuc2tron(z["gencode"], d), t[1:-1]), i
        seen |= set(d.tolist())
    assert len(seen) >= 16


SGPT6_NAMES = ["sig5", "sig3", "sigS", "sigT", "sigE", "sigI", "phs5", "phs3"]


def load_scan_p():
    z = np.load(golden_io.GOLDEN_DIR / "scan_p.npz")
    prm = {k[4:]: (z[k] if z[k].ndim else z[k].item()) for k in z.files if k.startswith("prm_")}
    segs = [{"tron": z[f"s{i}_tron"], "sgpt6": z[f"s{i}_sgpt6"], "int53": z[f"s{i}_int53"]} for i in range(int(z["n"]))]
    return prm, segs


def sgpt6_equal(got, want, L):
    """got / want: (len + 2, 8) tables.  Entries that depend on the INT53 halves the reference never
    writes (column 0's 3' half, column len - 1's 5' half: uninitialised memory) are skipped."""
    for c, nm in enumerate(SGPT6_NAMES):
        lo, hi = {"sig3": (1, L), "sig5": (0, L - 1), "phs3": (2, L), "phs5": (0, L - 2)}.get(nm, (0, L))
        if not np.array_equal(got[lo:hi, c], want[lo:hi, c]):
            return False
    return True


def test_oracle_protein_scan_matches_reference_tables(oracle):
    """Exinon::intron53_p (four PSSMs on tron codes, ExinPot::calcScr_3, termination-codon rules,
    intron phases) against the SGPT6 tables of the unmodified reference"""
    prm, segs = load_scan_p()
    assert len(segs) >= 8 and prm["codepot"].size == 3 * 4096
    for i, sg in enumerate(segs):
        tron = sg["tron"][1:-1]
        o = oracle.exinon_scan_p(prm, tron)
        assert sgpt6_equal(o["sgpt6"], sg["sgpt6"], len(tron)), i
        L = len(tron)
        assert np.array_equal(o["int53"][0:L - 1] & 0x0f0f, sg["int53"][0:L - 1] & 0x0f0f), i
        assert np.array_equal(o["int53"][1:L + 1] & 0xf0f0, sg["int53"][1:L + 1] & 0xf0f0), i
