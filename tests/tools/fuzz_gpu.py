#!/usr/bin/env python
"""GPU-box tool: randomized parity sweep of every device path against the C oracle.
usage: fuzz_gpu.py [n_per_case] [seed]   (prints one line per case; exit 1 on any mismatch)"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import golden_io                    # noqa: E402
import oracle_harness as O          # noqa: E402
from spaln_b200 import Engine, EngineH, Problem, ProblemH, workload  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 150
SEED = int(sys.argv[2]) if len(sys.argv) > 2 else 1
FLAGS = [(1, 1, 1, 1), (0, 0, 0, 0), (1, 0, 1, 0), (0, 1, 0, 1), (1, 1, 0, 0), (0, 0, 1, 1), (1, 0, 0, 0)]
bad_total = 0


def dna_problem(prm, rng, i):
    kind = i % 5
    qr = [(8, 40), (30, 300), (200, 900), (1400, 2100), (16, 16)][kind]
    fl = [(3, 40), (20, 300), (50, 1500), (30, 200), (5, 60)][kind]
    g, q, _ = workload.plant_gene(rng, qlen_range=qr, flank=fl, intron_scale=float(rng.choice([0.3, 1.0, 4.0])))
    a, b = workload.encode_dna(q), workload.encode_dna(g)
    s5, s3 = workload.synthetic_signals(b, rng)
    al, ar, bl, br = 0, len(a), 0, len(b)
    if i % 4 == 1 and len(a) > 30:
        al, ar = int(rng.integers(0, 9)), len(a) - int(rng.integers(0, 9))
        bl, br = int(rng.integers(0, 20)), len(b) - int(rng.integers(0, 20))
    sh = int(rng.choice([100, 100, 30, 8]))
    lw, up = workload.stripe(al, ar, bl, br, sh)
    f = FLAGS[int(rng.integers(0, len(FLAGS)))] if i % 3 == 0 else (1, 1, 1, 1)
    return {"a": np.concatenate([[0], a, [0]]).astype(np.uint8), "b": np.concatenate([[0], b, [0]]).astype(np.uint8),
            "sig5": s5, "sig3": s3, "int53": workload.synthetic_int53(b),
            "a_left": al, "a_right": ar, "b_left": bl, "b_right": br,
            "a_exgl": f[0], "a_exgr": f[1], "b_exgl": f[2], "b_exgr": f[3], "lw": lw, "up": up}


def prot_problem(prm, rng, i):
    kind = i % 4
    pr = [(8, 30), (20, 200), (150, 420), (560, 700)][kind]
    fl = [(10, 80), (30, 400), (40, 900), (40, 150)][kind]
    pb = workload.protein_problem(rng, plen_range=pr, flank=fl, sh=int(rng.choice([100, 100, 30])),
                                  intron_scale=float(rng.choice([0.5, 1.0, 3.0])))
    if i % 4 == 1 and pb["a_right"] > 30:
        pb["a_left"] = int(rng.integers(0, 9)); pb["a_right"] -= int(rng.integers(0, 9))
        pb["b_left"] = int(rng.integers(0, 40)); pb["b_right"] -= int(rng.integers(0, 40))
        pb["lw"], pb["up"] = workload.stripe31(pb["a_left"], pb["a_right"], pb["b_left"], pb["b_right"], 100)
    if i % 3 == 0:
        er = [(1, 1), (0, 0), (1, 0)][int(rng.integers(0, 3))]     # (0, 1) is undefined in the reference
        pb.update(a_exgl=int(rng.integers(0, 2)), a_exgr=er[0], b_exgl=int(rng.integers(0, 2)), b_exgr=er[1])
    pb["alen"] = len(pb["a"]) - 2
    return pb


def report(name, bad, n):
    global bad_total
    bad_total += len(bad)
    print(f"{name:44s} {n:5d} problems  {len(bad)} mismatches {bad[:3] if bad else ''}", flush=True)


ONLY = sys.argv[3] if len(sys.argv) > 3 else ""
for fixture in ("dna_A2_global", "dna_A2_local", "dna_A3_global", "dna_A2_dagp") if ONLY in ("", "dna") else ():
    prm, _ = golden_io.load(fixture)
    rng = np.random.default_rng(SEED * 1000 + hash(fixture) % 997)
    probs = [dna_problem(prm, rng, i) for i in range(N)]
    P = [Problem.from_export(pb, pb["lw"], pb["up"]) for pb in probs]
    eng = Engine(prm, device=0)
    res, sco = eng.forwardS1_wip(P), eng.scoreonlyS1_wip(P)
    bad = []
    for i, (pb, r, s) in enumerate(zip(probs, res, sco)):
        o = O.forward_wip(prm, pb, cap=1 << 17)
        if r.status or r.score != o["score"] or not np.array_equal(r.skl, o["skl"]) or \
                s.score != O.scoreonly_wip(prm, pb)["score"]:
            bad.append(i)
    report(f"{fixture} forward / score-only", bad, N)
    # scalar exact-ILD kernel (what the driver uses for blocks with < 8 rows), on the smaller problems
    sel = [i for i, pb in enumerate(probs) if (pb["a_right"] - pb["a_left"]) * (pb["up"] - pb["lw"]) < 400000]
    rn = eng.forwardS_ng([P[i] for i in sel])
    bad = []
    for i, r in zip(sel, rn):
        o = O.trcbk_ng(prm, probs[i], cap=1 << 17)
        if r.status or r.score != o["score"] or not np.array_equal(r.skl, o["skl"]):
            bad.append(i)
    report(f"{fixture} forwardS_ng (scalar)", bad, len(sel))
    if int(prm["Noll"]) == 3:       # double affine: no Hirschberg pass on the device (nor in the oracle)
        rl = eng.lspS_ng(P, max_vmf_space=1 << 25, sh=int(prm["sh"]), alg=2)
        bad = []
        for i, (pb, r) in enumerate(zip(probs, rl)):
            o = O.lsp(prm, pb, cap=1 << 17, max_vmf_space=1 << 25)
            if o["unsupported"]:
                if r.status != 3:
                    bad.append(i)
            elif r.status or r.score != o["score"] or not np.array_equal(r.skl, o["skl"]):
                bad.append(i)
        report(f"{fixture} lspS_ng -V32M", bad, N)
        eng.close()
        continue
    # Hirschberg pass + driver at small -V
    sel = [i for i, pb in enumerate(probs) if pb["a_right"] - pb["a_left"] >= 32]
    PU = [Problem.from_export(probs[i], probs[i]["lw"], probs[i]["up"]) for i in sel]
    nim = []
    for i, p in zip(sel, PU):
        m = probs[i]["a_right"] - probs[i]["a_left"]
        p.n_imd = int(rng.integers(1, max(2, min(7, m // 16))))
        nim.append(p.n_imd)
    ru = eng.hirschbergS1_wip(PU)
    bad = []
    for i, k, r in zip(sel, nim, ru):
        o = O.hirschberg_wip(prm, probs[i], k)
        if r.status or r.score != o["score"] or list(r.ranges) != o["ranges"]:
            bad.append(i)
    report(f"{fixture} hirschbergS1_wip", bad, len(sel))
    for vmf in (1 << 17, 1 << 20):
        rl = eng.lspS_ng(P, max_vmf_space=vmf, sh=int(prm["sh"]), alg=2)
        bad = []
        for i, (pb, r) in enumerate(zip(probs, rl)):
            o = O.lsp(prm, pb, cap=1 << 17, max_vmf_space=vmf)
            if o["unsupported"]:
                if r.status != 3:
                    bad.append(i)
                continue
            if r.status or r.score != o["score"] or not np.array_equal(r.skl, o["skl"]):
                bad.append(i)
        report(f"{fixture} lspS_ng -V{vmf >> 10}K", bad, N)
    eng.close()

for fixture in ("prot_A2_global", "prot_A2_local") if ONLY in ("", "prot") else ():
    prm, _ = golden_io.load_protein(fixture)
    rng = np.random.default_rng(SEED * 1000 + 17 + hash(fixture) % 997)
    probs = [prot_problem(prm, rng, i) for i in range(N)]
    P = [ProblemH.from_export(pb, pb["lw"], pb["up"]) for pb in probs]
    eng = EngineH(prm, device=0)
    res, sco = eng.forwardH1_wip(P), eng.forwardH1_wip(P, trace=False)
    bad = []
    for i, (pb, r, s) in enumerate(zip(probs, res, sco)):
        o = O.forward_h1_wip(prm, pb, cap=1 << 17)
        if r.status or r.score != o["score"] or not np.array_equal(r.skl, o["skl"]) or s.score != o["score"]:
            bad.append(i)
    report(f"{fixture} forwardH1_wip / score-only", bad, N)
    sel = [i for i, pb in enumerate(probs) if pb["a_right"] - pb["a_left"] >= 32]
    PU = [ProblemH.from_export(probs[i], probs[i]["lw"], probs[i]["up"]) for i in sel]
    nim = []
    for i, p in zip(sel, PU):
        m = probs[i]["a_right"] - probs[i]["a_left"]
        p.n_imd = int(rng.integers(1, max(2, min(7, m // 16))))
        nim.append(p.n_imd)
    ru = eng.hirschbergH1_wip(PU)
    bad = []
    for i, k, r in zip(sel, nim, ru):
        o = O.hirschberg_h1_wip(prm, probs[i], k)
        if r.status or r.score != o["score"] or list(r.ranges) != o["ranges"] or \
                not np.array_equal(r.cpos[:, :8], o["cpos"][:, :8]):
            bad.append(i)
    report(f"{fixture} hirschbergH1_wip", bad, len(sel))
    for vmf in (1 << 17, 1 << 19):
        rl = eng.lspH_ng(P, max_vmf_space=vmf, sh=int(prm["sh"]), alg=2)
        bad = []
        for i, (pb, r) in enumerate(zip(probs, rl)):
            o = O.lsp_h(prm, pb, cap=1 << 17, max_vmf_space=vmf)
            if o["unsupported"]:
                if r.status != 3:
                    bad.append(i)
                continue
            if r.status or r.score != o["score"] or not np.array_equal(r.skl, o["skl"]):
                bad.append(i)
        report(f"{fixture} lspH_ng -V{vmf >> 10}K", bad, N)
    eng.close()

EOU = 2 ** 31 - 1 - 2


def sudh_equal(r, o):
    """scalar Hirschberg pass: score; ranges and crossing records (with their diagonal bounds) if a path exists"""
    if r.status or r.score != o["score"]:
        return False
    if r.score <= -(1 << 28):
        return True
    if list(r.ranges) != o["ranges"]:
        return False
    for ra, rb in zip(r.cpos[: len(o["cpos"])].tolist(), o["cpos"].tolist()):
        ka = ra.index(EOU) if EOU in ra[:8] else 8
        kb = rb.index(EOU) if EOU in rb[:8] else 8
        if ra[:ka] != rb[:kb] or (ka > 0 and ra[8:] != rb[8:]):
            return False
    return True


# ---- the reference's default mode -A0: scalar Hirschberg passes and the drivers with alg = 0
for fixture in ("dna_A0_udh", "dna_A0_udh_local", "dna_A0_udh_dagp") if ONLY in ("", "a0", "dna_a0") else ():
    prm, _ = golden_io.load(fixture)
    rng = np.random.default_rng(SEED * 1000 + 29 + hash(fixture) % 997)
    probs = [dna_problem(prm, rng, i) for i in range(N)]
    P = [Problem.from_export(pb, pb["lw"], pb["up"]) for pb in probs]
    eng = Engine(prm, device=0)
    sel = [i for i, pb in enumerate(probs) if pb["a_right"] - pb["a_left"] >= 12]
    PU = [Problem.from_export(probs[i], probs[i]["lw"], probs[i]["up"]) for i in sel]
    want = []
    for i, p in zip(sel, PU):
        m = probs[i]["a_right"] - probs[i]["a_left"]
        p.n_imd = int(rng.integers(1, max(2, min(9, m // 5))))
        intvl = (m + p.n_imd) // (p.n_imd + 1)
        nq = p.n_imd - 1 if intvl * p.n_imd == m else p.n_imd
        want.append(O.hirschberg_ng(prm, probs[i], nq, intvl) if nq >= 1 else None)
    bad = [i for i, r, o in zip(sel, eng.hirschbergS_ng(PU), want) if o is not None and not sudh_equal(r, o)]
    report(f"{fixture} hirschbergS_ng", bad, len(sel))
    for vmf in (1 << 17, 1 << 20, 1 << 25):
        rl = eng.lspS_ng(P, max_vmf_space=vmf, sh=int(prm["sh"]), alg=0)
        bad = []
        for i, (pb, r) in enumerate(zip(probs, rl)):
            o = O.lsp(prm, pb, cap=1 << 17, max_vmf_space=vmf)
            if o["unsupported"]:
                if r.status != 3:
                    bad.append(i)
            elif r.status or r.score != o["score"] or not np.array_equal(r.skl, o["skl"]):
                bad.append(i)
        report(f"{fixture} lspS_ng alg 0 -V{vmf >> 10}K", bad, N)
    eng.close()

for fixture in ("prot_A0_udh", "prot_A0_udh_local") if ONLY in ("", "a0", "prot_a0") else ():
    prm, _ = golden_io.load_protein(fixture)
    rng = np.random.default_rng(SEED * 1000 + 31 + hash(fixture) % 997)
    probs = [prot_problem(prm, rng, i) for i in range(N)]
    for pb in probs:
        pb["int53"] = workload.synthetic_int53(workload.encode_dna(pb["genome"]))
    P = [ProblemH.from_export(pb, pb["lw"], pb["up"]) for pb in probs]
    eng = EngineH(prm, device=0)
    rn = eng.forwardH_ng(P)
    bad = []
    for i, (pb, r) in enumerate(zip(probs, rn)):
        o = O.trcbk_h_ng(prm, pb, cap=1 << 17)
        if r.status or r.score != o["score"] or not np.array_equal(r.skl, o["skl"]):
            bad.append(i)
    report(f"{fixture} forwardH_ng", bad, N)
    sel = [i for i, pb in enumerate(probs) if pb["a_right"] - pb["a_left"] >= 12]
    PU = [ProblemH.from_export(probs[i], probs[i]["lw"], probs[i]["up"]) for i in sel]
    want = []
    for i, p in zip(sel, PU):
        m = probs[i]["a_right"] - probs[i]["a_left"]
        p.n_imd = int(rng.integers(1, max(2, min(9, m // 5))))
        intvl = (m + p.n_imd) // (p.n_imd + 1)
        nq = p.n_imd - 1 if intvl * p.n_imd == m else p.n_imd
        want.append(O.hirschberg_h_ng(prm, probs[i], nq, intvl) if nq >= 1 else None)
    bad = [i for i, r, o in zip(sel, eng.hirschbergH_ng(PU), want) if o is not None and not sudh_equal(r, o)]
    report(f"{fixture} hirschbergH_ng", bad, len(sel))
    for vmf in (1 << 17, 1 << 19, 1 << 25):
        rl = eng.lspH_ng(P, max_vmf_space=vmf, sh=int(prm["sh"]), alg=0)
        bad = []
        for i, (pb, r) in enumerate(zip(probs, rl)):
            o = O.lsp_h(prm, pb, cap=1 << 17, max_vmf_space=vmf)
            if o["unsupported"]:
                if r.status != 3:
                    bad.append(i)
            elif r.status or r.score != o["score"] or not np.array_equal(r.skl, o["skl"]):
                bad.append(i)
        report(f"{fixture} lspH_ng alg 0 -V{vmf >> 10}K", bad, N)
    eng.close()

print("TOTAL mismatches:", bad_total)
sys.exit(1 if bad_total else 0)
