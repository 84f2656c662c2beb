#!/usr/bin/env python
"""Build-container tool: pins the scalar exact-ILD restatement (oracle/spaln_oracle_ng.c) against
the unmodified reference (Aln2s1::trcbkalignS_ng forced onto its scalar branch) on random problems.
usage: sweep_oracle_scalar.py [n] [seed] [reference options]     (one option string per process)"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_harness as O          # noqa: E402
import ref_harness as R             # noqa: E402
from spaln_b200 import workload as synth    # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
SEED = int(sys.argv[2]) if len(sys.argv) > 2 else 3
OPTS = sys.argv[3] if len(sys.argv) > 3 else "-Q0 -A2 -S1 -yX0 -TDictyost"
FLAGS = [(1, 1, 1, 1), (0, 0, 0, 0), (1, 0, 1, 0), (0, 1, 0, 1), (1, 1, 0, 0), (0, 0, 1, 1), (1, 0, 0, 0)]

ref = R.Reference(OPTS)
p = ref.params()
rng = np.random.default_rng(SEED)
bad = 0
for i in range(N):
    kind = i % 4
    qr = [(1, 8), (5, 40), (30, 300), (200, 600)][kind]
    fl = [(3, 60), (10, 200), (20, 500), (50, 800)][kind]
    g, q, _ = synth.plant_gene(rng, qlen_range=qr, flank=fl, intron_scale=float(rng.choice([0.3, 1.0, 4.0])))
    t = ref.task(g, q, comrev_query=(i % 7 == 3))
    f = FLAGS[i % len(FLAGS)] if i % 3 == 0 else (1, 1, 1, 1)
    t.set(a_exgl=f[0], a_exgr=f[1], b_exgl=f[2], b_exgr=f[3])
    if i % 5 == 1 and len(q) > 20:
        t.set(a_left=int(rng.integers(0, 5)), a_right=len(q) - int(rng.integers(0, 5)),
              b_left=int(rng.integers(0, 10)), b_right=len(g) - int(rng.integers(0, 10)))
    lw, up = t.stripe(int(rng.choice([100, 100, 30, 8])))
    ex = t.export()
    ex.update(int53=t.export_int53(), lw=lw, up=up)
    pp = dict(p)
    pp.update(t.export_ng_tables(max(4096, ex["blen"] + 2)))
    rs = t.scalar(lw, up)
    o = O.trcbk_ng(pp, ex)
    sa, so = t.scorealone(lw, up), O.scorealone_ng(pp, ex)["score"]
    if sa != so:
        bad += 1
        if bad < 4:
            print("SCOREALONE MISMATCH", i, len(q), len(g), f, sa, so)
    if rs["score"] != o["score"] or not np.array_equal(rs["skl"], o["skl"]):
        bad += 1
        if bad < 4:
            print("MISMATCH", i, len(q), len(g), f, rs["score"], o["score"], len(rs["skl"]), len(o["skl"]))
    t.close()
print(f"{OPTS}: {bad} mismatches in {N} problems")
sys.exit(1 if bad else 0)
