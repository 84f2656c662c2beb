"""GPU: the CUDA path (through the C-ABI) against (1) golden vectors from the
unmodified reference, (2) the C oracle on seeded inputs, (3) size-independent
properties at larger sizes.  Bar: bit-exact scores and trace-back corners."""
import numpy as np
import pytest

import golden_io

pytestmark = pytest.mark.gpu


def _engine(prm):
    from spaln_b200 import Engine
    return Engine(prm, device=0)


def _problems(probs):
    from spaln_b200 import Problem
    return [Problem.from_export(pb, pb["lw"], pb["up"]) for pb in probs]


@pytest.mark.parametrize("name", golden_io.GOLDEN_NAMES + golden_io.DAGP_NAMES)
def test_forward_wip_matches_reference_golden(name):
    prm, probs = golden_io.load(name)
    eng = _engine(prm)
    res = eng.forwardS1_wip(_problems(probs))
    for i, (pb, r) in enumerate(zip(probs, res)):
        assert r.status == 0, (name, i, pb["tag"])
        assert r.score == pb["score"], (name, i, pb["tag"], r.score, pb["score"])
        assert np.array_equal(r.skl, pb["skl"]), (name, i, pb["tag"])
    eng.close()


@pytest.mark.parametrize("name", golden_io.GOLDEN_NAMES + golden_io.DAGP_NAMES)
def test_scoreonly_wip_matches_reference_golden(name):
    prm, probs = golden_io.load(name)
    eng = _engine(prm)
    res = eng.scoreonlyS1_wip(_problems(probs))
    for i, (pb, r) in enumerate(zip(probs, res)):
        assert r.score == pb["score_only"], (name, i, pb["tag"], r.score, pb["score_only"])
    eng.close()


def _synthetic(prm, rng, n, qlen, flank, intron_scale=1.0, flags=None, sub=None):
    from spaln_b200 import workload
    out = []
    for _ in range(n):
        g, q, _t = workload.plant_gene(rng, qlen_range=qlen, flank=flank, intron_scale=intron_scale)
        a = workload.encode_dna(q)
        b = workload.encode_dna(g)
        s5, s3 = workload.synthetic_signals(b, rng)
        al, ar, bl, br = 0, len(a), 0, len(b)
        if sub:
            al, ar, bl, br = sub[0], len(a) - sub[1], sub[2], len(b) - sub[3]
        lw, up = workload.stripe(al, ar, bl, br, int(prm["sh"]))
        f = flags or (1, 1, 1, 1)
        out.append({"a": np.concatenate([[0], a, [0]]).astype(np.uint8),
                    "b": np.concatenate([[0], b, [0]]).astype(np.uint8),
                    "sig5": s5, "sig3": s3, "int53": workload.synthetic_int53(b), "a_left": al, "a_right": ar, "b_left": bl,
                    "b_right": br, "a_exgl": f[0], "a_exgr": f[1], "b_exgl": f[2], "b_exgr": f[3],
                    "lw": lw, "up": up})
    return out


@pytest.mark.parametrize("name,flags,sub", [
    ("dna_A2_global", None, None),
    ("dna_A2_global", (0, 0, 0, 0), None),
    ("dna_A2_global", (1, 0, 0, 1), (7, 3, 11, 5)),
    ("dna_A2_local", None, None),
    ("dna_A3_global", None, None),
    ("dna_A2_dagp", None, None),
    ("dna_A2_dagp", (0, 0, 0, 0), None),
    ("dna_A2_dagp", (1, 0, 0, 1), (7, 3, 11, 5)),
])
def test_forward_wip_matches_oracle_seeded(oracle, name, flags, sub):
    prm, _ = golden_io.load(name)
    import zlib
    rng = np.random.default_rng(zlib.crc32(repr((name, flags, sub)).encode()))
    probs = _synthetic(prm, rng, 24, (30, 700), (30, 400), flags=flags, sub=sub)
    probs += _synthetic(prm, rng, 3, (1500, 2200), (100, 300), flags=flags, sub=sub)   # re-basing
    eng = _engine(prm)
    res = eng.forwardS1_wip(_problems(probs))
    sco = eng.scoreonlyS1_wip(_problems(probs))
    for i, (pb, r, s) in enumerate(zip(probs, res, sco)):
        o = oracle.forward_wip(prm, pb)
        assert r.status == 0
        assert r.score == o["score"], (i, r.score, o["score"])
        assert np.array_equal(r.skl, o["skl"]), i
        assert s.score == oracle.scoreonly_wip(prm, pb)["score"], i
    eng.close()


def test_long_introns_match_oracle(oracle):
    """intron-length counter beyond the quantile table / int16 range"""
    prm, _ = golden_io.load("dna_A2_global")
    rng = np.random.default_rng(77)
    probs = _synthetic(prm, rng, 2, (300, 500), (50, 100), intron_scale=300.0)
    eng = _engine(prm)
    res = eng.forwardS1_wip(_problems(probs))
    for pb, r in zip(probs, res):
        o = oracle.forward_wip(prm, pb, cap=1 << 17)
        assert r.score == o["score"]
        assert np.array_equal(r.skl, o["skl"])
    eng.close()


def test_batch_properties_full_size():
    """BASELINE config-2 sized problems (1-3 kb cDNA): properties that do not
    need the oracle -- resubmission is idempotent, batch order does not matter,
    corner lists are monotone walks inside the matrix.  (Score-only and
    trace-back scores may differ: the reference's score-only kernel stops one
    step earlier and lacks the empty-intron guard.)"""
    prm, _ = golden_io.load("dna_A2_global")
    rng = np.random.default_rng(5)
    probs = _synthetic(prm, rng, 48, (1000, 3000), (500, 3000))
    P = _problems(probs)
    eng = _engine(prm)
    r1 = eng.forwardS1_wip(P)
    r2 = eng.forwardS1_wip(P[::-1])[::-1]
    for pb, x, y in zip(probs, r1, r2):
        assert x.status == 0
        assert x.score == y.score and np.array_equal(x.skl, y.skl)
        skl = x.skl
        assert len(skl) >= 2
        assert np.all(np.diff(skl[:, 0]) <= 0) and np.all(np.diff(skl[:, 1]) <= 0)
        assert skl[:, 0].min() >= pb["a_left"] and skl[:, 0].max() <= pb["a_right"]
        assert skl[:, 1].min() >= pb["b_left"] and skl[:, 1].max() <= pb["b_right"]
        assert x.cells > 0
    # planted genes must be recovered with a clearly positive score
    assert np.median([x.score for x in r1]) > 5000
    eng.close()


def test_empty_batch_and_overflow():
    prm, probs = golden_io.load("dna_A2_global")
    eng = _engine(prm)
    assert eng.forwardS1_wip([]) == []
    P = _problems(probs[:1])
    P[0].skl_cap = 2
    r = eng.forwardS1_wip(P)[0]
    assert r.status == 1 and r.score == probs[0]["score"]
    assert np.array_equal(r.skl, probs[0]["skl"][:2])
    eng.close()


# ---------------------------------------------------------------------------
# unidirectional Hirschberg pass (hirschbergS1_wip)
# ---------------------------------------------------------------------------
def _cpos_equal(a, b):
    from spaln_b200.capi import END_OF_ULK as EOU
    for ra, rb in zip(a.tolist(), b.tolist()):
        if ra[0] == EOU and rb[0] == EOU and ra[2] == rb[2]:
            continue
        ka = ra.index(EOU) if EOU in ra else 10
        kb = rb.index(EOU) if EOU in rb else 10
        if ra[:ka] != rb[:kb]:
            return False
    return True


@pytest.mark.parametrize("name", ["dna_A2_udh", "dna_A6_udh_recursive", "dna_A2_udh_local"])
def test_hirschberg_wip_matches_reference_golden(name):
    prm, probs = golden_io.load(name)
    sel = [pb for pb in probs if "udh_nim" in pb
           and pb["udh_ranges"][0] <= pb["udh_ranges"][1] and pb["udh_ranges"][2] <= pb["udh_ranges"][3]]
    assert len(sel) >= 20
    P = _problems(sel)
    for p, pb in zip(P, sel):
        p.n_imd = pb["udh_nim"]
    eng = _engine(prm)
    res = eng.hirschbergS1_wip(P)
    for i, (pb, r) in enumerate(zip(sel, res)):
        assert r.status == 0
        assert r.score == pb["udh_score"], (name, i, pb["tag"], r.score, pb["udh_score"])
        assert list(r.ranges) == pb["udh_ranges"].tolist(), (name, i, pb["tag"])
        assert _cpos_equal(r.cpos, pb["udh_cpos"]), (name, i, pb["tag"])
    eng.close()


@pytest.mark.parametrize("flags,fixture", [(None, "dna_A2_global"), ((0, 0, 0, 0), "dna_A2_global"),
                                           ((1, 0, 0, 1), "dna_A2_global"), (None, "dna_A2_local"),
                                           ((1, 0, 1, 0), "dna_A2_local"), ((0, 1, 0, 1), "dna_A2_local")])
def test_hirschberg_wip_matches_oracle_seeded(oracle, flags, fixture):
    prm, _ = golden_io.load(fixture)
    rng = np.random.default_rng(4242 + (sum(flags) if flags else 9))
    probs = _synthetic(prm, rng, 20, (60, 900), (40, 500), flags=flags)
    probs += _synthetic(prm, rng, 3, (1500, 2400), (100, 400), flags=flags)     # re-basing
    nims = []
    for pb in probs:
        m = pb["a_right"] - pb["a_left"]
        nims.append(int(rng.integers(1, max(2, min(9, m // 16)))))
    P = _problems(probs)
    for p, k in zip(P, nims):
        p.n_imd = k
    eng = _engine(prm)
    res = eng.hirschbergS1_wip(P)
    for i, (pb, r, k) in enumerate(zip(probs, res, nims)):
        o = oracle.hirschberg_wip(prm, pb, k)
        assert r.score == o["score"], (i, k, r.score, o["score"])
        assert list(r.ranges) == o["ranges"], (i, k)
        assert _cpos_equal(r.cpos, o["cpos"]), (i, k)
    eng.close()


# ---------------------------------------------------------------------------
# scalar exact-ILD kernel: Aln2s1::trcbkalignS_ng's scalar branch (forwardS_ng + Vmf)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_io.GOLDEN_NAMES + golden_io.DAGP_NAMES + golden_io.UDH_NAMES + golden_io.CIP_NAMES)
def test_scalar_kernel_matches_reference_golden(name):
    prm, probs = golden_io.load(name)
    eng = _engine(prm)
    res = eng.forwardS_ng(_problems(probs))
    for i, (pb, r) in enumerate(zip(probs, res)):
        assert r.status == 0, (name, i, pb["tag"], r.status)
        assert r.score == pb["ng_score"], (name, i, pb["tag"], r.score, pb["ng_score"])
        assert np.array_equal(r.skl, pb["ng_skl"]), (name, i, pb["tag"])
    eng.close()


@pytest.mark.parametrize("name,flags,sub", [
    ("dna_A2_global", None, None),
    ("dna_A2_global", (0, 0, 0, 0), None),
    ("dna_A2_global", (1, 0, 0, 1), (2, 1, 11, 5)),
    ("dna_A2_local", None, None),
    ("dna_A2_local", (1, 1, 0, 0), None),
    ("dna_A2_dagp", None, None),
    ("dna_A2_dagp", (0, 0, 0, 0), None),
    ("dna_A2_tetrapod", None, None),
])
def test_scalar_kernel_matches_oracle_seeded(oracle, name, flags, sub):
    prm, _ = golden_io.load(name)
    import zlib
    rng = np.random.default_rng(zlib.crc32(repr(("ng", name, flags, sub)).encode()))
    probs = _synthetic(prm, rng, 40, (4, 8), (20, 2500), flags=flags, sub=sub)       # what the driver sends
    probs += _synthetic(prm, rng, 12, (30, 300), (30, 400), flags=flags, sub=sub)    # any size works
    eng = _engine(prm)
    res = eng.forwardS_ng(_problems(probs))
    for i, (pb, r) in enumerate(zip(probs, res)):
        o = oracle.trcbk_ng(prm, pb)
        assert r.status == 0, (i, r.status)
        assert r.score == o["score"], (i, r.score, o["score"])
        assert np.array_equal(r.skl, o["skl"]), i
    eng.close()


@pytest.mark.parametrize("name", ["dna_A2_global", "dna_A2_dagp", "dna_A2_local"])
def test_exact_ild_kernels_with_cip_score_match_oracle(oracle, name):
    """Cip_score (gspaln_task.cip; src/gsinfo.h:127-139, `sigB` at src/fwd2s1.cc:254,338,1191,1262):
    seeded problems with a random bonus table by query position, trace-back and score-only"""
    prm, _ = golden_io.load(name)
    import zlib
    rng = np.random.default_rng(zlib.crc32(("cip" + name).encode()))
    probs = _synthetic(prm, rng, 30, (3, 8), (20, 2500), flags=(0, 0, 0, 0))
    probs += _synthetic(prm, rng, 20, (30, 300), (30, 400))
    for pb in probs:
        tab = np.zeros(len(pb["a"]) + 2, np.int32)
        k = rng.integers(0, len(tab), size=max(2, len(tab) // 8))
        tab[k] = rng.integers(1, 4, size=len(k)) * 20 * int(prm.get("scale", 10)) // 10
        pb["cip"] = tab
    eng = _engine(prm)
    n_diff = 0
    for i, (pb, r) in enumerate(zip(probs, eng.forwardS_ng(_problems(probs)))):
        o = oracle.trcbk_ng(prm, pb)
        assert r.status == 0 and r.score == o["score"] and np.array_equal(r.skl, o["skl"]), (name, i)
        n_diff += o["score"] != oracle.trcbk_ng(prm, dict(pb, cip=None))["score"]
    assert n_diff >= 10, n_diff
    for i, (pb, r) in enumerate(zip(probs, eng.scorealoneS_ng(_problems(probs)))):
        assert r.status == 0 and r.score == oracle.scorealone_ng(prm, pb)["score"], (name, i)
    eng.close()


@pytest.mark.parametrize("name", golden_io.GOLDEN_NAMES + golden_io.DAGP_NAMES + golden_io.CIP_NAMES)
def test_scalar_scoreonly_kernel_matches_reference_and_oracle(oracle, name):
    """Aln2s1::scorealoneS_ng on the device: reference goldens (any size: three int rows of
    workspace per problem), then seeded problems against the oracle"""
    prm, probs = golden_io.load(name)
    eng = _engine(prm)
    for i, (pb, r) in enumerate(zip(probs, eng.scorealoneS_ng(_problems(probs)))):
        assert r.status == 0 and r.score == pb["ng_score_only"], (name, i, pb["tag"], r.score, pb["ng_score_only"])
    import zlib
    rng = np.random.default_rng(zlib.crc32(("sa" + name).encode()))
    more = _synthetic(prm, rng, 30, (2, 12), (20, 1500), flags=(1, 0, 0, 1))
    more += _synthetic(prm, rng, 20, (100, 900), (50, 900))
    more += _synthetic(prm, rng, 10, (30, 300), (30, 400), flags=(0, 0, 0, 0), sub=(2, 1, 11, 5))
    for i, (pb, r) in enumerate(zip(more, eng.scorealoneS_ng(_problems(more)))):
        assert r.status == 0 and r.score == oracle.scorealone_ng(prm, pb)["score"], (name, i)
    eng.close()


def test_homscore_dispatch_matches_reference_golden():
    """HomScoreS_ng: m < 4 -> scorealoneS_ng, otherwise scoreonlyS1_wip (src/fwd2s1.cc:2704-2712)"""
    prm, probs = golden_io.load("dna_A2_global")
    eng = _engine(prm)
    res = eng.HomScoreS_ng(_problems(probs))
    n_small = 0
    for pb, r in zip(probs, res):
        small = pb["a_right"] - pb["a_left"] < 4
        n_small += small
        assert r.score == (pb["ng_score_only"] if small else pb["score_only"]), pb["tag"]
    assert n_small >= 2
    eng.close()


def test_scalar_kernel_needs_its_tables():
    """without gspaln_set_ng_tables / int53 the kind is refused, and the driver reports
    GSPALN_ST_UNSUPPORTED for blocks with fewer than 8 query rows"""
    from spaln_b200 import EngineError
    prm, probs = golden_io.load("dna_A2_udh")
    bare = {k: v for k, v in prm.items() if k not in ("penalty", "sig53tab")}
    eng = _engine(bare)
    with pytest.raises(EngineError):
        eng.forwardS_ng(_problems(probs[:2]))
    tiny = [pb for pb in probs if pb["a_right"] - pb["a_left"] < 8]
    assert tiny
    res = eng.lspS_ng(_problems(tiny), max_vmf_space=int(prm["MaxVmfSpace"]), sh=int(prm["sh"]))
    assert all(r.status == 3 for r in res)
    eng.close()


# ---------------------------------------------------------------------------
# the whole driver: lspS_ng (trace-back vs UDH dispatch + block re-alignment)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["dna_A2_udh", "dna_A6_udh_recursive", "dna_A2_global", "dna_A2_udh_local", "dna_A2_cip"])
def test_lsp_driver_matches_reference_golden(oracle, name):
    prm, probs = golden_io.load(name)
    eng = _engine(prm)
    res = eng.lspS_ng(_problems(probs), max_vmf_space=int(prm["MaxVmfSpace"]), sh=int(prm["sh"]),
                      ubh=int(prm["ubh"]), alg=int(prm["alg"]))
    n_ok = 0
    for i, (pb, r) in enumerate(zip(probs, res)):
        # blocks with < 8 query rows run on the scalar kernel, like in the reference
        assert r.status == 0, (name, i, pb["tag"])
        if "lsp_score" in pb:
            want_score, want_skl = pb["lsp_score"], pb["lsp_skl"]
        elif pb["a_right"] - pb["a_left"] < 8:
            want_score, want_skl = pb["ng_score"], pb["ng_skl"]
        else:
            want_score, want_skl = pb["score"], pb["skl"]
        assert r.score == want_score, (name, i, pb["tag"], r.score, want_score)
        assert np.array_equal(r.skl, want_skl), (name, i, pb["tag"])
        n_ok += 1
    assert n_ok == len(probs)
    eng.close()


def test_lsp_driver_matches_oracle_seeded(oracle):
    """default -V (32 MiB): config-2 sized problems mix trace-back and UDH dispatch"""
    prm, _ = golden_io.load("dna_A2_global")
    rng = np.random.default_rng(31337)
    probs = _synthetic(prm, rng, 10, (1000, 3000), (500, 4000))
    probs += _synthetic(prm, rng, 10, (100, 900), (100, 900))
    eng = _engine(prm)
    for vmf in (int(prm["MaxVmfSpace"]), 1 << 20):
        res = eng.lspS_ng(_problems(probs), max_vmf_space=vmf, sh=int(prm["sh"]), alg=2)
        n_udh = 0
        for i, (pb, r) in enumerate(zip(probs, res)):
            o = oracle.lsp(prm, pb, cap=1 << 16, max_vmf_space=vmf)
            assert not o["unsupported"]
            m, n = pb["a_right"] - pb["a_left"], pb["b_right"] - pb["b_left"]
            n_udh += 2.0 * m * (n + m) >= vmf
            assert r.status == 0
            assert r.score == o["score"], (vmf, i, r.score, o["score"])
            assert np.array_equal(r.skl, o["skl"]), (vmf, i)
        assert n_udh >= 5
    eng.close()


def test_lsp_driver_config4_shape_local(oracle):
    """BASELINE config 4 shape: mRNA of ~2.5 kb against a locus with 20x longer introns (tens of kb),
    local mode (-LS), default -V: every problem takes the multi-intermediate Hirschberg route and
    its block re-alignments (some with fewer than 8 rows -> scalar kernel)"""
    from spaln_b200 import workload
    prm, _ = golden_io.load("dna_A2_local")
    rng = np.random.default_rng(404)
    probs = []
    for _ in range(4):
        r = workload.config2_problem(rng, qlen_range=(1500, 3500), intron_scale=20.0, flank=(500, 3000))
        b = r["b"]
        probs.append({"a": np.concatenate([[0], r["a"], [0]]).astype(np.uint8),
                      "b": np.concatenate([[0], b, [0]]).astype(np.uint8),
                      "sig5": r["sig5"], "sig3": r["sig3"], "int53": workload.synthetic_int53(b),
                      "a_left": 0, "a_right": len(r["a"]), "b_left": 0, "b_right": len(b),
                      "a_exgl": 1, "a_exgr": 1, "b_exgl": 1, "b_exgr": 1, "lw": r["lw"], "up": r["up"]})
    assert min(len(pb["b"]) for pb in probs) > 10000
    eng = _engine(prm)
    vmf = int(prm["MaxVmfSpace"])
    res = eng.lspS_ng(_problems(probs), max_vmf_space=vmf, sh=int(prm["sh"]), alg=2)
    for i, (pb, r) in enumerate(zip(probs, res)):
        o = oracle.lsp(prm, pb, cap=1 << 17, max_vmf_space=vmf)
        assert not o["unsupported"] and r.status == 0, (i, r.status)
        assert 2.0 * pb["a_right"] * (pb["b_right"] + pb["a_right"]) >= vmf         # Hirschberg route
        assert r.score == o["score"], (i, r.score, o["score"])
        assert np.array_equal(r.skl, o["skl"]), i
        assert r.score > 5000               # the planted gene is found across the long introns
    eng.close()


def test_lsp_packed_equals_object_api():
    prm, probs = golden_io.load("dna_A2_udh")
    eng = _engine(prm)
    P = _problems(probs)
    opts = dict(max_vmf_space=int(prm["MaxVmfSpace"]), sh=int(prm["sh"]), ubh=int(prm["ubh"]), alg=int(prm["alg"]))
    want = eng.lspS_ng(P, **opts)
    b = eng.lsp_packed(eng.pack(P), **opts)
    for i, w in enumerate(want):
        assert b.scores[i] == w.score and b.status[i] == w.status
        assert np.array_equal(b.corners(i), w.skl)
    eng.close()


def test_packed_batch_api_equals_object_api():
    prm, probs = golden_io.load("dna_A2_global")
    eng = _engine(prm)
    P = _problems(probs)
    want = eng.forwardS1_wip(P)
    b = eng.submit_packed(eng.pack(P))
    for i, w in enumerate(want):
        assert b.scores[i] == w.score and b.status[i] == w.status
        assert np.array_equal(b.corners(i), w.skl)
    eng.close()


def test_streamed_submit_equals_resident_path():
    """A batch large enough for the one-shot submit to stream it in behind the running kernel
    (chunks on a second stream + watermark) must give exactly what upload / run / download of the
    resident batch gives; a few problems are also checked against the oracle by the caller tests."""
    prm, _ = golden_io.load("dna_A2_global")
    rng = np.random.default_rng(77)
    probs = _synthetic(prm, rng, 700, (800, 2000), (500, 4000))
    assert sum(pb["b_right"] - pb["b_left"] for pb in probs) > (4 << 20)    # above the chunking threshold
    P = _problems(probs)
    eng = _engine(prm)
    eng.upload(P)
    eng.run()
    want = eng.download()
    for rep in range(2):
        got = eng.forwardS1_wip(P)
        for i, (w, g) in enumerate(zip(want, got)):
            assert g.status == 0 and w.status == 0, (rep, i, g.status)
            assert g.score == w.score and np.array_equal(g.skl, w.skl), (rep, i)
    so = eng.scoreonlyS1_wip(P)
    eng.upload(P, kind=1)
    eng.run()
    for w, g in zip(eng.download(), so):
        assert w.score == g.score
    eng.close()


def _config5_problem(rng, m, W):
    """query of m nt; genome = the query with 0 / 1 / 2 inserted GT..AG introns and short flanks,
    d = n - m extra nucleotides in total; band = stripe() with the shoulder that makes the band
    W diagonals wide"""
    from spaln_b200 import workload
    d = int(rng.integers(0, max(1, W // 2)))
    q = workload.random_dna(rng, m)
    nin = int(rng.integers(0, 3)) if d >= 60 and m >= 40 else 0
    lens = []
    rest = d
    for _ in range(nin):
        il = int(rng.integers(25, max(26, rest // (nin + 1) + 26)))
        il = min(il, rest)
        if il >= 25:
            lens.append(il)
            rest -= il
    fl = rest // 2
    parts = [workload.random_dna(rng, fl)]
    cuts = sorted(int(x) for x in rng.choice(np.arange(10, m - 10), size=len(lens), replace=False)) if lens else []
    prev = 0
    for c, il in zip(cuts, lens):
        parts.append(q[prev:c])
        it = workload.random_dna(rng, il)
        it[:2] = np.frombuffer(b"GT", np.uint8)
        it[-2:] = np.frombuffer(b"AG", np.uint8)
        parts.append(it)
        prev = c
    parts.append(q[prev:])
    parts.append(workload.random_dna(rng, rest - fl))
    g = np.concatenate(parts)
    a, b = workload.encode_dna(q.tobytes().decode()), workload.encode_dna(g.tobytes().decode())
    s5, s3 = workload.synthetic_signals(b, rng)
    sh = max(1, (W - 3 - (len(b) - len(a))) // 2)
    lw, up = workload.stripe(0, len(a), 0, len(b), sh)
    return {"a": np.concatenate([[0], a, [0]]).astype(np.uint8),
            "b": np.concatenate([[0], b, [0]]).astype(np.uint8),
            "sig5": s5, "sig3": s3, "a_left": 0, "a_right": len(a), "b_left": 0,
            "b_right": len(b), "a_exgl": 1, "a_exgr": 1, "b_exgl": 1, "b_exgr": 1, "lw": lw, "up": up}


def test_config5_sweep_matches_oracle(oracle):
    """BASELINE config 5 shapes: band width W in {16 .. 1024} x query length m in {16 .. 4096}
    with m * W in [256, 65536] cells, 0 / 1 / 2 planted introns: trace-back and score-only
    against the oracle, global and local parameter sets."""
    for fixture in ("dna_A2_global", "dna_A2_local"):
        prm, _ = golden_io.load(fixture)
        rng = np.random.default_rng(55555)
        probs = []
        for W in (16, 32, 64, 128, 256, 512, 1024):
            for m in (16, 33, 64, 250, 1024, 4096):
                if 256 <= m * W <= 65536:
                    probs += [_config5_problem(rng, m, W) for _ in range(2)]
        assert len(probs) >= 40
        eng = _engine(prm)
        res = eng.forwardS1_wip(_problems(probs))
        sco = eng.scoreonlyS1_wip(_problems(probs))
        for i, (pb, r, s) in enumerate(zip(probs, res, sco)):
            o = oracle.forward_wip(prm, pb)
            assert r.status == 0, (fixture, i)
            assert r.score == o["score"], (fixture, i, pb["lw"], pb["up"], r.score, o["score"])
            assert np.array_equal(r.skl, o["skl"]), (fixture, i)
            assert s.score == oracle.scoreonly_wip(prm, pb)["score"], (fixture, i)
        eng.close()


def test_coalescing_queue_from_many_threads():
    """the literal drop-in shape: 16 host threads submit one problem at a time through
    gspaln_queue_*; the dispatcher coalesces them into batches; results equal the batch API"""
    import threading
    prm, probs = golden_io.load("dna_A2_global")
    rng = np.random.default_rng(9)
    probs = probs + _synthetic(prm, rng, 40, (100, 600), (50, 400))
    P = _problems(probs)
    eng = _engine(prm)
    want = eng.forwardS1_wip(P)
    q = eng.queue(max_batch=64, max_wait_us=500)
    got = [None] * len(P)
    nthr = 16

    def work(tid):
        for i in range(tid, len(P), nthr):
            got[i] = q.forwardS1_wip(P[i])

    th = [threading.Thread(target=work, args=(k,)) for k in range(nthr)]
    [t.start() for t in th]
    [t.join() for t in th]
    tasks, batches = q.stats()
    q.close()
    assert tasks == len(P) and batches < tasks          # something was coalesced
    for i, (w, g) in enumerate(zip(want, got)):
        assert g is not None and g.status == w.status and g.score == w.score, i
        assert np.array_equal(g.skl, w.skl), i
    eng.close()


def test_packed_kernel_hands_back_and_skips_what_it_cannot_run(oracle):
    """The packed int16x2 kernel (gspaln_packed.cuh) runs the eligible problems and the 32-bit
    kernel the rest; both must give the oracle's answer:
      * long near-perfect matches with large positive splice signals push max H + addends past
        32767 -> the monitor hands the problem back (status 6 never reaches the caller),
      * IUPAC codes other than A, C, G, T, N and signals beyond +-8192 are not eligible at all,
      * values at the bottom of the int16 range (global mode, unrelated sequences) stay exact on
        the packed path itself."""
    from spaln_b200 import workload
    prm, _ = golden_io.load("dna_A2_global")
    rng = np.random.default_rng(4242)
    probs = []
    # (1) monitor: 1450-row queries, no mutations, signals up to +3000
    for _ in range(3):
        g, q, _t = workload.plant_gene(rng, qlen_range=(1440, 1460), flank=(30, 60), sub=0.0, indel=0.0)
        a, b = workload.encode_dna(q), workload.encode_dna(g)
        s5, s3 = workload.synthetic_signals(b, rng)
        s5 = (s5.astype(np.int32) + rng.integers(0, 3000, size=len(s5))).astype(np.int16)
        s3 = (s3.astype(np.int32) + rng.integers(0, 3000, size=len(s3))).astype(np.int16)
        probs.append((a, b, s5, s3, (1, 1, 1, 1)))
    # (2) not eligible: ambiguity codes / huge signals
    for k in range(4):
        g, q, _t = workload.plant_gene(rng, qlen_range=(100, 400), flank=(30, 200))
        a, b = workload.encode_dna(q), workload.encode_dna(g)
        s5, s3 = workload.synthetic_signals(b, rng)
        if k % 2 == 0:
            b = b.copy(); b[rng.integers(0, len(b), size=5)] = rng.choice([1, 4, 6, 7, 8, 10, 12, 15], size=5)
            a = a.copy(); a[rng.integers(0, len(a), size=2)] = 16
        else:
            s3 = s3.copy(); s3[rng.integers(0, len(s3), size=4)] = -20000
            s5 = s5.copy(); s5[rng.integers(0, len(s5), size=4)] = 12000
        probs.append((a, b, s5, s3, (1, 1, 1, 1)))
    # (3) deep negative values: unrelated sequences, no free end gaps, N runs
    for _ in range(6):
        a = workload.encode_dna(workload.random_dna(rng, int(rng.integers(300, 900))).tobytes().decode())
        b = workload.encode_dna(workload.random_dna(rng, int(rng.integers(400, 1200))).tobytes().decode())
        b = b.copy(); p = int(rng.integers(0, len(b) - 40)); b[p:p + 30] = 16
        s5, s3 = workload.synthetic_signals(b, rng)
        probs.append((a, b, s5, s3, (0, 0, 0, 0)))
    pbs = []
    for a, b, s5, s3, f in probs:
        lw, up = workload.stripe(0, len(a), 0, len(b), int(prm["sh"]))
        pbs.append({"a": np.concatenate([[0], a, [0]]).astype(np.uint8), "b": np.concatenate([[0], b, [0]]).astype(np.uint8),
                    "sig5": s5, "sig3": s3, "int53": workload.synthetic_int53(b), "a_left": 0, "a_right": len(a),
                    "b_left": 0, "b_right": len(b), "a_exgl": f[0], "a_exgr": f[1], "b_exgl": f[2], "b_exgr": f[3],
                    "lw": lw, "up": up})
    eng = _engine(prm)
    res = eng.forwardS1_wip(_problems(pbs))
    sco = eng.scoreonlyS1_wip(_problems(pbs))
    for i, (pb, r, s) in enumerate(zip(pbs, res, sco)):
        o = oracle.forward_wip(prm, pb, cap=1 << 16)
        assert r.status == 0, (i, r.status)
        assert r.score == o["score"], (i, r.score, o["score"])
        assert np.array_equal(r.skl, o["skl"]), i
        assert s.score == oracle.scoreonly_wip(prm, pb)["score"], i
    eng.close()


def test_packed_and_32bit_kernels_agree(monkeypatch):
    """same batch with the packed kernels disabled (GSPALN_NO_PACKED=1): identical results"""
    prm, _ = golden_io.load("dna_A2_global")
    rng = np.random.default_rng(99)
    probs = _synthetic(prm, rng, 40, (40, 900), (30, 600)) + _synthetic(prm, rng, 4, (1600, 2600), (100, 800))
    eng = _engine(prm)
    fast = eng.forwardS1_wip(_problems(probs))
    fast_s = eng.scoreonlyS1_wip(_problems(probs))
    eng.close()
    monkeypatch.setenv("GSPALN_NO_PACKED", "1")
    eng = _engine(prm)
    slow = eng.forwardS1_wip(_problems(probs))
    slow_s = eng.scoreonlyS1_wip(_problems(probs))
    eng.close()
    for f, s, fs, ss in zip(fast, slow, fast_s, slow_s):
        assert f.status == 0 and s.status == 0
        assert f.score == s.score and np.array_equal(f.skl, s.skl)
        assert fs.score == ss.score


def test_resident_runs_repeat_for_every_chain_class(oracle):
    """upload once, run three times (every kernel of every chain-width class gets its ticket counter
    reset), download: the same answers as the one-shot submit and the oracle"""
    prm, _ = golden_io.load("dna_A2_global")
    rng = np.random.default_rng(31337)
    probs = (_synthetic(prm, rng, 12, (10, 40), (3, 30)) + _synthetic(prm, rng, 10, (50, 120), (10, 80)) +
             _synthetic(prm, rng, 8, (130, 300), (30, 300)) + _synthetic(prm, rng, 4, (600, 900), (300, 900)))
    eng = _engine(prm)
    ps = _problems(probs)
    eng.upload(ps)
    for _ in range(3):
        eng.run()
    res = eng.download()
    one = eng.forwardS1_wip(ps)
    for i, (pb, r, o1) in enumerate(zip(probs, res, one)):
        o = oracle.forward_wip(prm, pb)
        assert r.status == 0 and o1.status == 0
        assert r.score == o["score"] == o1.score, i
        assert np.array_equal(r.skl, o["skl"]) and np.array_equal(o1.skl, o["skl"]), i
    eng.close()


# ---------------------------------------------------------------------------
# -A0, the reference's default mode: scalar Hirschberg pass + the driver on the exact-ILD kernels
# ---------------------------------------------------------------------------
def _sudh_cpos_equal(a, b):
    for ra, rb in zip(a.tolist(), b.tolist()):
        ka = ra.index(EOU) if EOU in ra[:8] else 8
        kb = rb.index(EOU) if EOU in rb[:8] else 8
        if ra[:ka] != rb[:kb] or (ka > 0 and ra[8:] != rb[8:]):
            return False
    return True


EOU = 2 ** 31 - 1 - 2


@pytest.mark.parametrize("name", golden_io.A0_NAMES)
def test_scalar_hirschberg_pass_matches_reference_golden(name):
    """GSPALN_HIRSCHBERG_NG == Aln2s1::hirschbergS_ng: score, crossing records with their diagonal
    bounds, narrowed ranges -- 1, 2 and 5 intermediate rows; global, local, double affine"""
    prm, probs = golden_io.load(name)
    eng = _engine(prm)
    n = 0
    for nn in (1, 2, 5):
        sel = [pb for pb in probs if f"sudh{nn}_nim" in pb]
        P = _problems(sel)
        for p in P:
            p.n_imd = nn
        for i, (pb, r) in enumerate(zip(sel, eng.hirschbergS_ng(P))):
            assert r.status == 0, (name, nn, i, pb["tag"], r.status)
            assert r.score == pb[f"sudh{nn}_score"], (name, nn, i, pb["tag"], r.score, pb[f"sudh{nn}_score"])
            if r.score > -(1 << 28):
                assert list(r.ranges) == pb[f"sudh{nn}_ranges"].tolist(), (name, nn, i, pb["tag"])
                assert _sudh_cpos_equal(r.cpos[: pb[f"sudh{nn}_nim"] + 1], pb[f"sudh{nn}_cpos"]), (name, nn, i, pb["tag"])
            n += 1
    assert n >= 50
    eng.close()


@pytest.mark.parametrize("name,flags,sub", [
    ("dna_A0_udh", None, None),
    ("dna_A0_udh", (0, 0, 0, 0), None),
    ("dna_A0_udh", (1, 0, 0, 1), (7, 3, 11, 5)),
    ("dna_A0_udh_local", None, None),
    ("dna_A0_udh_dagp", None, None),
    ("dna_A0_udh_dagp", (0, 1, 1, 0), None),
])
def test_scalar_hirschberg_pass_matches_oracle_seeded(oracle, name, flags, sub):
    """seeded planted genes (synthetic signals: many more candidate sites than real tables), any
    number of intermediate rows, rows spaced closer than one pass of the kernel included"""
    prm, _ = golden_io.load(name)
    import zlib
    rng = np.random.default_rng(zlib.crc32(repr(("xudh", name, flags, sub)).encode()))
    probs = _synthetic(prm, rng, 16, (40, 700), (30, 500), flags=flags, sub=sub)
    probs += _synthetic(prm, rng, 4, (900, 1500), (100, 300), flags=flags, sub=sub)
    P = _problems(probs)
    want = []
    for pb, p in zip(probs, P):
        m = pb["a_right"] - pb["a_left"]
        p.n_imd = int(rng.choice([1, 2, 3, 7, max(1, m // 16), max(1, m // 9)]))
        intvl = (m + p.n_imd) // (p.n_imd + 1)
        nq = p.n_imd - 1 if intvl * p.n_imd == m else p.n_imd
        want.append(oracle.hirschberg_ng(prm, pb, nq, intvl) if nq >= 1 else None)
    eng = _engine(prm)
    n = 0
    for i, (pb, p, r, o) in enumerate(zip(probs, P, eng.hirschbergS_ng(P), want)):
        if o is None:
            continue
        assert r.status == 0 and r.score == o["score"], (name, i, p.n_imd, r.status, r.score, o["score"])
        if r.score > -(1 << 28):
            assert list(r.ranges) == o["ranges"], (name, i, p.n_imd)
            assert _sudh_cpos_equal(r.cpos[: len(o["cpos"])], o["cpos"]), (name, i, p.n_imd)
        n += 1
    assert n >= 15
    eng.close()


@pytest.mark.parametrize("name", golden_io.A0_NAMES)
def test_lsp_driver_scalar_mode_matches_reference_golden(oracle, name):
    """gspaln_lsp with alg = 0 == Aln2s1::lspS_ng under -A0: hexagonal volume in the dispatch, exact-ILD
    trace-back for every block, the scalar Hirschberg pass, blocks banded by its diagonal bounds"""
    prm, probs = golden_io.load(name)
    eng = _engine(prm)
    n_route = 0
    for vmf in (int(prm["MaxVmfSpace"]), 32 * 1024 * 1024):
        res = eng.lspS_ng(_problems(probs), max_vmf_space=vmf, sh=int(prm["sh"]), ubh=int(prm["ubh"]), alg=0)
        for i, (pb, r) in enumerate(zip(probs, res)):
            assert r.status == 0, (name, vmf, i, pb["tag"], r.status)
            if vmf == int(prm["MaxVmfSpace"]):
                want_score, want_skl = pb["lsp_score"], pb["lsp_skl"]
                m, n = pb["a_right"] - pb["a_left"], pb["b_right"] - pb["b_left"]
                k, q = pb["lw"] - pb["b_left"] + pb["a_right"], pb["b_right"] - pb["a_left"] - pb["up"]
                n_route += 2.0 * (m * n - (k * k + q * q) / 2) >= vmf
            else:
                o = oracle.lsp(prm, pb, max_vmf_space=vmf)
                want_score, want_skl = o["score"], o["skl"]
            assert r.score == want_score, (name, vmf, i, pb["tag"], r.score, want_score)
            assert np.array_equal(r.skl, want_skl), (name, vmf, i, pb["tag"])
    assert n_route >= 12
    eng.close()


def test_exact_ild_trace_back_reruns_with_the_full_record_store(oracle, monkeypatch):
    """the first run of GSPALN_FORWARD_NG sizes the path-record store (the reference's Vmf) for a
    typical problem; problems that overflow it are run again with the worst-case bound inside
    gspaln_submit.  With a starved first run (1/8 record per cell) every larger problem takes that
    second run, and the results must not change."""
    monkeypatch.setenv("GSPALN_NG_REC_EIGHTHS", "1")
    prm, _ = golden_io.load("dna_A2_global")
    rng = np.random.default_rng(4242)
    probs = _synthetic(prm, rng, 24, (100, 600), (200, 1500))
    eng = _engine(prm)
    res = eng.forwardS_ng(_problems(probs))
    for i, (pb, r) in enumerate(zip(probs, res)):
        o = oracle.trcbk_ng(prm, pb)
        assert r.status == 0 and r.score == o["score"] and np.array_equal(r.skl, o["skl"]), (i, r.status)
    eng.close()
