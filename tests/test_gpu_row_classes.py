"""GPU: the exact intron-length kernels at the edges of their row classes.

The device runs a problem either with one warp (a CTA carries several problems) or with a CTA of
eight warps per problem (gspaln_ng.cuh NG_WIDE_ROWS = 128, gspaln_xudh.cuh XUDH_WIDE_ROWS = 128,
gspaln_hng.cuh HNG_WIDE_ROWS = 96); a pass of the wavefront is 32 resp. 256 rows tall.  Query
lengths one below, at and one above those numbers (class switch, last pass of one row, first pass
exactly full) against the C oracle, which the reference pins (tests/test_oracle_golden.py,
test_oracle_protein.py).  Bar: bit-exact scores, corners, ranges and crossing records."""
import numpy as np
import pytest

import golden_io

pytestmark = pytest.mark.gpu

ROWS = [31, 32, 33, 95, 96, 97, 127, 128, 129, 255, 256, 257, 300, 511, 512, 513]
EOU = 2 ** 31 - 1 - 2


def _sudh_equal(r, o):
    if r.status or r.score != o["score"]:
        return False
    if r.score <= -(1 << 28):           # no path: the reference leaves the rest undefined
        return True
    if list(r.ranges) != o["ranges"]:
        return False
    for ra, rb in zip(r.cpos[: len(o["cpos"])].tolist(), o["cpos"].tolist()):
        ka = ra.index(EOU) if EOU in ra[:8] else 8
        kb = rb.index(EOU) if EOU in rb[:8] else 8
        if ra[:ka] != rb[:kb] or (ka > 0 and ra[8:] != rb[8:]):
            return False
    return True


def _dna(prm, rng, rows, flags):
    from spaln_b200 import workload
    g, q, _ = workload.plant_gene(rng, qlen_range=(rows + 6, rows + 12), flank=(20, 200),
                                  intron_scale=float(rng.choice([0.3, 1.0])))
    a, b = workload.encode_dna(q), workload.encode_dna(g)
    assert len(a) >= rows
    s5, s3 = workload.synthetic_signals(b, rng)
    al = int(rng.integers(0, len(a) - rows + 1))
    ar = al + rows
    lw, up = workload.stripe(al, ar, 0, len(b), int(prm["sh"]))
    return {"a": np.concatenate([[0], a, [0]]).astype(np.uint8), "b": np.concatenate([[0], b, [0]]).astype(np.uint8),
            "sig5": s5, "sig3": s3, "int53": workload.synthetic_int53(b),
            "a_left": al, "a_right": ar, "b_left": 0, "b_right": len(b),
            "a_exgl": flags[0], "a_exgr": flags[1], "b_exgl": flags[2], "b_exgr": flags[3], "lw": lw, "up": up}


def _n_imd_cases(rows):
    """(n_imd as lspS_ng asks for it, number and spacing after its even-division correction)"""
    out = []
    for n_req in (1, 3, 7):
        intvl = (rows + n_req) // (n_req + 1)
        nq = n_req - 1 if intvl * n_req == rows else n_req
        if nq >= 1:
            out.append((n_req, nq, intvl))
    return out


@pytest.mark.parametrize("name", ["dna_A0_udh", "dna_A0_udh_local", "dna_A0_udh_dagp"])
def test_dna_exact_kernels_at_class_edges(oracle, name):
    from spaln_b200 import Engine, Problem
    prm, _ = golden_io.load(name)
    rng = np.random.default_rng(4100 + len(name))
    flags = [(1, 1, 1, 1), (0, 0, 0, 0), (1, 0, 1, 0), (0, 1, 0, 1)]
    probs = [_dna(prm, rng, rows, flags[i % 4]) for i, rows in enumerate(ROWS)]
    P = [Problem.from_export(pb, pb["lw"], pb["up"]) for pb in probs]
    eng = Engine(prm, device=0)
    bad = []
    for rows, pb, r, s in zip(ROWS, probs, eng.forwardS_ng(P), eng.scorealoneS_ng(P)):
        o = oracle.trcbk_ng(prm, pb, cap=1 << 17)
        if r.status or r.score != o["score"] or not np.array_equal(r.skl, o["skl"]):
            bad.append(("forwardS_ng", rows, r.status, r.score, o["score"]))
        so = oracle.scorealone_ng(prm, pb)
        if s.status or s.score != so["score"]:
            bad.append(("scorealoneS_ng", rows, s.status, s.score, so["score"]))
    PU, want, tag = [], [], []
    for rows, pb in zip(ROWS, probs):
        for n_req, nq, intvl in _n_imd_cases(rows):
            p = Problem.from_export(pb, pb["lw"], pb["up"])
            p.n_imd = n_req
            PU.append(p)
            want.append(oracle.hirschberg_ng(prm, pb, nq, intvl))
            tag.append((rows, n_req))
    for t, r, o in zip(tag, eng.hirschbergS_ng(PU), want):
        if not _sudh_equal(r, o):
            bad.append(("hirschbergS_ng", t, r.status, r.score, o["score"]))
    # the driver with alg 0 on the same problems: small -V (Hirschberg route) and the default
    for vmf in (1 << 17, 1 << 25):
        for rows, pb, r in zip(ROWS, probs, eng.lspS_ng(P, max_vmf_space=vmf, sh=int(prm["sh"]), alg=0)):
            o = oracle.lsp(prm, pb, cap=1 << 17, max_vmf_space=vmf)
            if o["unsupported"]:
                if r.status != 3:
                    bad.append(("lspS_ng", vmf, rows, "expected unsupported", r.status))
            elif r.status or r.score != o["score"] or not np.array_equal(r.skl, o["skl"]):
                bad.append(("lspS_ng", vmf, rows, r.status, r.score, o["score"]))
    eng.close()
    assert not bad, (name, bad)


def _protein(prm, rng, rows, er):
    from spaln_b200 import workload
    pb = workload.protein_problem(rng, plen_range=(rows + 4, rows + 10), flank=(30, 300), sh=int(prm["sh"]),
                                  intron_scale=float(rng.choice([0.5, 1.0])))
    assert pb["a_right"] >= rows
    pb["a_left"] = int(rng.integers(0, pb["a_right"] - rows + 1))
    pb["a_right"] = pb["a_left"] + rows
    pb["lw"], pb["up"] = workload.stripe31(pb["a_left"], pb["a_right"], pb["b_left"], pb["b_right"], int(prm["sh"]))
    pb.update(a_exgl=er[0], a_exgr=er[1], b_exgl=er[2], b_exgr=er[3])
    pb["alen"] = len(pb["a"]) - 2
    pb["int53"] = workload.synthetic_int53(workload.encode_dna(pb["genome"]))
    return pb


@pytest.mark.parametrize("name", ["prot_A0_udh", "prot_A0_udh_local"])
def test_protein_exact_kernels_at_class_edges(oracle, name):
    from spaln_b200 import EngineH, ProblemH
    prm, _ = golden_io.load_protein(name)
    rng = np.random.default_rng(4200 + len(name))
    flags = [(1, 1, 1, 1), (0, 0, 0, 0), (1, 1, 0, 0), (0, 1, 1, 1)]      # (a_exgr, b_exgr) = (0, 1) is undefined in the reference
    rows_h = [r for r in ROWS if r <= 300]
    probs = [_protein(prm, rng, rows, flags[i % 4]) for i, rows in enumerate(rows_h)]
    P = [ProblemH.from_export(pb, pb["lw"], pb["up"]) for pb in probs]
    eng = EngineH(prm, device=0)
    bad = []
    for rows, pb, r in zip(rows_h, probs, eng.forwardH_ng(P)):
        o = oracle.trcbk_h_ng(prm, pb, cap=1 << 17)
        if r.status or r.score != o["score"] or not np.array_equal(r.skl, o["skl"]):
            bad.append(("forwardH_ng", rows, r.status, r.score, o["score"]))
    PU, want, tag = [], [], []
    for rows, pb in zip(rows_h, probs):
        for n_req, nq, intvl in _n_imd_cases(rows):
            p = ProblemH.from_export(pb, pb["lw"], pb["up"])
            p.n_imd = n_req
            PU.append(p)
            want.append(oracle.hirschberg_h_ng(prm, pb, nq, intvl))
            tag.append((rows, n_req))
    for t, r, o in zip(tag, eng.hirschbergH_ng(PU), want):
        if not _sudh_equal(r, o):
            bad.append(("hirschbergH_ng", t, r.status, r.score, o["score"]))
    for vmf in (1 << 17, 1 << 25):
        for rows, pb, r in zip(rows_h, probs, eng.lspH_ng(P, max_vmf_space=vmf, sh=int(prm["sh"]), alg=0)):
            o = oracle.lsp_h(prm, pb, cap=1 << 17, max_vmf_space=vmf)
            if o["unsupported"]:
                if r.status != 3:
                    bad.append(("lspH_ng", vmf, rows, "expected unsupported", r.status))
            elif r.status or r.score != o["score"] or not np.array_equal(r.skl, o["skl"]):
                bad.append(("lspH_ng", vmf, rows, r.status, r.score, o["score"]))
    eng.close()
    assert not bad, (name, bad)
