"""GPU: the splice-signal scan kernel (gspaln_exinon_scan, through the C-ABI) against (1) the
Exinon tables of the unmodified reference carried by the golden fixtures, (2) the C oracle on
seeded segments, (3) size-independent properties at genome scale.  Bar: bit-exact shorts."""
import numpy as np
import pytest

import golden_io
from test_oracle_scan import scan_equal

pytestmark = pytest.mark.gpu


def _scan(prm):
    from spaln_b200 import ExinonScan
    return ExinonScan(prm, device=0)


@pytest.mark.parametrize("name", ["dna_A2_global", "dna_A2_tetrapod", "dna_A2_udh"])
def test_scan_matches_reference_tables(name):
    prm, probs = golden_io.load(name)
    sc = _scan(prm)
    for i, pb in enumerate(probs):
        s5, s3, i53 = sc.scan(pb["b"][1:-1])
        assert scan_equal({"sig5": s5, "sig3": s3, "int53": i53}, pb), (name, i, pb["tag"])
    sc.close()


def _random_codes(rng, n, amb=0.002):
    c = rng.choice(np.array([2, 3, 5, 9], np.uint8), size=n, p=[0.295, 0.205, 0.205, 0.295])
    k = rng.random(n) < amb
    c[k] = rng.choice(np.array([16, 4, 6, 10, 1, 0], np.uint8), size=int(k.sum()))
    return c


@pytest.mark.parametrize("n", [0, 1, 2, 3, 17, 18, 19, 1023, 1024, 1025, 2047, 100003])
def test_scan_matches_oracle_edge_lengths(oracle, n):
    """empty and tiny segments, lengths around the CTA tile (1024 positions), ambiguity codes"""
    prm, _ = golden_io.load("dna_A2_global")
    rng = np.random.default_rng(1000 + n)
    codes = _random_codes(rng, n, amb=0.01)
    sc = _scan(prm)
    s5, s3, i53 = sc.scan(codes)
    o = oracle.exinon_scan(prm, codes)
    assert np.array_equal(s5, o["sig5"]) and np.array_equal(s3, o["sig3"]) and np.array_equal(i53, o["int53"])
    sc.close()


def test_scan_markov_order_0_and_1(oracle):
    """lower-order PSSMs take the other branch of PatMat::calcPatMat"""
    prm, _ = golden_io.load("dna_A2_global")
    rng = np.random.default_rng(5)
    codes = _random_codes(rng, 5000, amb=0.01)
    for order, rows in ((0, 4), (1, 20)):
        p = dict(prm)
        for nm in ("pat5", "pat3"):
            meta = np.array(prm[nm + "_meta"]).copy()
            cols = int(meta[1])
            meta[0], meta[4] = rows, order
            p[nm + "_meta"] = meta
            p[nm + "_mtx"] = rng.normal(-0.5, 1.0, rows * cols).astype(np.float32)
        sc = _scan(p)
        s5, s3, i53 = sc.scan(codes)
        o = oracle.exinon_scan(p, codes)
        assert np.array_equal(s5, o["sig5"]) and np.array_equal(s3, o["sig3"]), order
        sc.close()


def test_scan_genome_scale_properties(oracle):
    """16 Mb segment (the oracle checks a window): the scan is local, so any window of the big
    result equals the scan of that window alone away from the window's ends; resident re-runs
    are idempotent"""
    prm, _ = golden_io.load("dna_A2_global")
    rng = np.random.default_rng(8)
    n = 16 * 1024 * 1024 + 7
    codes = _random_codes(rng, n)
    sc = _scan(prm)
    sc.upload(codes)
    sc.run()
    a = sc.download()
    sc.run()
    b = sc.download()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    for lo in (0, 5_000_001, n - 200_000):
        hi = lo + 200_000
        o = oracle.exinon_scan(prm, codes[lo:hi])
        m = 64      # margin: PSSM windows reach 18 back and 10 ahead
        assert np.array_equal(a[0][lo + m:hi - m], o["sig5"][m:hi - lo - m])
        assert np.array_equal(a[1][lo + m:hi - m], o["sig3"][m:hi - lo - m])
        assert np.array_equal(a[2][lo + m:hi - m], o["int53"][m:hi - lo - m])
    # canonical sites carry their class: every GT is a class-3 donor, every AG a class-3 acceptor
    gt = np.nonzero((codes[:-1] == 5) & (codes[1:] == 9))[0]
    assert np.all(((a[2][gt] >> 8) & 15) == 3)
    ag = np.nonzero((codes[:-1] == 2) & (codes[1:] == 5))[0] + 2
    assert np.all(((a[2][ag] >> 12) & 15) == 3)
    t = sc.timing()
    assert t["kernel_ms"] > 0
    sc.close()


def test_nuc2tron_matches_reference_and_oracle(oracle):
    """gspaln_nuc2tron against the reference's tron codes (golden) and the oracle on seeded
    segments of every length around the 16-byte vector width"""
    from spaln_b200 import nuc2tron
    z = np.load(golden_io.GOLDEN_DIR / "nuc2tron.npz")
    gc = z["gencode"]
    for i in range(int(z["n"])):
        got, _ = nuc2tron(gc, z[f"dna{i}"])
        assert np.array_equal(got, z[f"tron{i}"][1:-1]), i
    rng = np.random.default_rng(12)
    for n in list(range(0, 40)) + [255, 256, 257, 4095, 4096, 4097, 1_000_003]:
        codes = np.concatenate([[0], _random_codes(rng, n, amb=0.02), [0]]).astype(np.uint8)
        got, ms = nuc2tron(gc, codes)
        assert np.array_equal(got, oracle.nuc2tron(gc, codes)), n
    # a non-standard table goes through unchanged semantics
    gc2 = gc.copy()
    gc2[[56, 58]] = 20
    codes = np.concatenate([[0], _random_codes(rng, 5000, amb=0.0), [0]]).astype(np.uint8)
    assert np.array_equal(nuc2tron(gc2, codes)[0], oracle.nuc2tron(gc2, codes))


def _sg_table(sg):
    """SGPT6 records -> (n, 8) int16 table in the fixture's column order"""
    return np.stack([sg[k].astype(np.int16) for k in ("sig5", "sig3", "sigS", "sigT", "sigE", "sigI", "phs5", "phs3")], 1)


def test_protein_scan_matches_reference_and_oracle(oracle):
    """gspaln_exinon_scan_p (Exinon::intron53_p on a TRON segment) against the reference's SGPT6
    tables (golden) and against the oracle on seeded segments incl. ambiguity codes, stop codons
    and lengths around the CTA tile"""
    from spaln_b200 import ExinonScanP, nuc2tron
    from test_oracle_scan import load_scan_p, sgpt6_equal
    prm, segs = load_scan_p()
    sc = ExinonScanP(prm, device=0)
    for i, s in enumerate(segs):
        tron = s["tron"][1:-1]
        sg, i53 = sc.scan(tron)
        assert sgpt6_equal(_sg_table(sg), s["sgpt6"], len(tron)), i
    gc = np.load(golden_io.GOLDEN_DIR / "nuc2tron.npz")["gencode"]
    rng = np.random.default_rng(77)
    for n in (0, 1, 2, 5, 6, 7, 8, 30, 1023, 1024, 1025, 50_001):
        codes = np.concatenate([[0], _random_codes(rng, n, amb=0.01), [0]]).astype(np.uint8)
        tron = oracle.nuc2tron(gc, codes)
        sg, i53 = sc.scan(tron)
        o = oracle.exinon_scan_p(prm, tron)
        assert np.array_equal(_sg_table(sg), o["sgpt6"]), n
        assert np.array_equal(i53, o["int53"]), n
    sc.close()
    # parameters outside this version's limits are refused, not approximated
    from spaln_b200 import EngineError
    bad = dict(prm)
    bad["any"] = 2
    with pytest.raises(EngineError):
        ExinonScanP(bad, device=0)
