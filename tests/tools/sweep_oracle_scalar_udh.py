#!/usr/bin/env python
"""Build-container tool: pins the restatements of the scalar Hirschberg passes
(oracle/spaln_oracle_udhng.c so_hirschberg_ng, oracle/spaln_oracle_hng.c so_hirschberg_h_ng) against
the unmodified reference (Aln2s1::hirschbergS_ng src/fwd2s1.cc:764-1104, Aln2h1::hirschbergH_ng
src/fwd2h1.cc:941-1520: the passes of the default mode -A0) on random planted genes: score,
narrowed ranges, crossing records with their diagonal bounds, for 1 to 8 intermediate rows spaced
as lsp*_ng spaces them.
usage: sweep_oracle_scalar_udh.py dna|prot [n] [seed] [reference options]   (one option string per process)"""
import os
import pickle
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_harness as O          # noqa: E402
import ref_harness as R             # noqa: E402
from spaln_b200 import workload as synth    # noqa: E402

WHAT = sys.argv[1] if len(sys.argv) > 1 else "dna"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 200
SEED = int(sys.argv[3]) if len(sys.argv) > 3 else 5
PROT = WHAT == "prot"
OPTS = sys.argv[4] if len(sys.argv) > 4 else ("-Q0 -A0 -yX0 -TDictyost" if PROT else "-Q0 -A0 -S1 -yX0 -TDictyost")
FLAGS = [(1, 1, 1, 1), (0, 0, 0, 0), (1, 0, 1, 0), (0, 1, 0, 1), (1, 1, 0, 0), (0, 0, 1, 1), (1, 0, 0, 0)]
EOU = 2 ** 31 - 1 - 2


def equal(r, o):
    if r["score"] != o["score"]:
        return False
    if r["score"] <= -(1 << 28):            # no path: the reference leaves the rest undefined
        return True
    if list(r["ranges"]) != list(o["ranges"]):
        return False
    for ra, rb in zip(r["cpos"].tolist(), o["cpos"].tolist()):
        ka = ra.index(EOU) if EOU in ra[:8] else 8
        kb = rb.index(EOU) if EOU in rb[:8] else 8
        if ra[:ka] != rb[:kb] or (ka > 0 and ra[8:] != rb[8:]):
            return False
    return True


def guarded(fn):
    """fn() in a forked child: on some inputs without a path through an intermediate row the
    reference's pass dereferences a null intermediate (src/fwd2s1.cc:1093) -- those are counted,
    not compared"""
    rd, wr = os.pipe()
    pid = os.fork()
    if pid == 0:
        os.close(rd)
        try:
            with os.fdopen(wr, "wb") as f:
                pickle.dump(fn(), f)
        finally:
            os._exit(0)
    os.close(wr)
    with os.fdopen(rd, "rb") as f:
        data = f.read()
    _, st = os.waitpid(pid, 0)
    return pickle.loads(data) if st == 0 and data else None


ref = R.Reference(OPTS, protein=PROT)
p = ref.params()
if PROT:
    p.update(ref.scalar_p_tables())
rng = np.random.default_rng(SEED)
bad = done = crashed = 0
for i in range(N):
    kind = i % 3
    if PROT:
        g, q, _ = synth.plant_protein_gene(rng, plen_range=[(16, 40), (30, 150), (120, 300)][kind],
                                           flank=[(20, 200), (30, 400), (40, 300)][kind])
    else:
        g, q, _ = synth.plant_gene(rng, qlen_range=[(30, 80), (60, 300), (250, 700)][kind],
                                   flank=[(10, 200), (20, 500), (50, 800)][kind],
                                   intron_scale=float(rng.choice([0.3, 1.0, 4.0])))
    t = ref.task(g, q)          # (no unrelated pairs: without a path the reference dereferences a null intermediate, src/fwd2s1.cc:1093)
    if i % 3 == 0:
        if PROT:
            er = [(1, 1), (0, 0), (1, 0)][int(rng.integers(0, 3))]     # (0, 1) is undefined in the reference
            t.set(a_exgl=int(rng.integers(0, 2)), a_exgr=er[0], b_exgl=int(rng.integers(0, 2)), b_exgr=er[1])
        else:
            f = FLAGS[int(rng.integers(0, len(FLAGS)))]
            t.set(a_exgl=f[0], a_exgr=f[1], b_exgl=f[2], b_exgr=f[3])
    if i % 5 == 1 and len(q) > 20:
        t.set(a_left=int(rng.integers(0, 4)), a_right=len(q) - int(rng.integers(0, 4)),
              b_left=int(rng.integers(0, 20)), b_right=len(g) - int(rng.integers(0, 20)))
    lw, up = (t.stripe31 if PROT else t.stripe)(int(rng.choice([100, 100, 30])))
    ex = t.export_p() if PROT else t.export()
    ex.update(int53=t.export_int53(), lw=lw, up=up)
    pp = dict(p)
    pp.update(t.export_ng_tables(max(4096, ex["blen"] + 2)))
    m = ex["a_right"] - ex["a_left"]
    for nn in sorted(set(int(x) for x in rng.integers(1, 9, size=2))):
        if m < 4 * nn:
            continue
        intvl = (m + nn) // (nn + 1)
        nq = nn - 1 if intvl * nn == m else nn
        if nq < 1:
            continue
        rs = guarded(lambda: (t.scalar_udh_p if PROT else t.scalar_udh)(lw, up, nq, intvl))
        if rs is None:
            crashed += 1
            continue
        o = (O.hirschberg_h_ng if PROT else O.hirschberg_ng)(pp, ex, nq, intvl)
        done += 1
        if not equal(rs, o):
            bad += 1
            if bad < 5:
                print("MISMATCH", i, len(q), len(g), t.info(), nq, intvl, rs["score"], o["score"], rs["ranges"], o["ranges"])
    t.close()
print(f"{WHAT} {OPTS}: {bad} mismatches in {done} passes over {N} problems ({crashed} more passes crashed the reference)", flush=True)
sys.exit(1 if bad else 0)
