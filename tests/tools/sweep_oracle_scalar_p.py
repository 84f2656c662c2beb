#!/usr/bin/env python
"""Build-container tool: pins the scalar protein restatement (oracle/spaln_oracle_hng.c) against the
unmodified reference (Aln2h1::trcbkalignH_ng forced onto its scalar branch).
usage: sweep_oracle_scalar_p.py [n] [seed] [reference options]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_harness as O          # noqa: E402
import ref_harness as R             # noqa: E402
from spaln_b200 import workload as synth    # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
SEED = int(sys.argv[2]) if len(sys.argv) > 2 else 3
OPTS = sys.argv[3] if len(sys.argv) > 3 else "-Q0 -A2 -yX0 -TDictyost"

ref = R.Reference(OPTS, protein=True)
p = ref.params()
p.update(ref.scalar_p_tables())
rng = np.random.default_rng(SEED)
bad = 0
for i in range(N):
    kind = i % 4
    pr = [(1, 8), (3, 20), (20, 120), (100, 260)][kind]
    fl = [(10, 80), (20, 200), (30, 400), (40, 300)][kind]
    g, q, _ = synth.plant_protein_gene(rng, plen_range=pr, flank=fl)
    t = ref.task(g, q)
    if i % 3 == 0:
        er = [(1, 1), (0, 0), (1, 0)][int(rng.integers(0, 3))]
        t.set(a_exgl=int(rng.integers(0, 2)), a_exgr=er[0], b_exgl=int(rng.integers(0, 2)), b_exgr=er[1])
    if i % 5 == 1 and len(q) > 12:
        t.set(a_left=int(rng.integers(0, 4)), a_right=len(q) - int(rng.integers(0, 4)),
              b_left=int(rng.integers(0, 30)), b_right=len(g) - int(rng.integers(0, 30)))
    lw, up = t.stripe31(int(rng.choice([100, 100, 30])))
    ex = t.export_p()
    ex.update(int53=t.export_int53(), lw=lw, up=up)
    pp = dict(p)
    pp.update(t.export_ng_tables(max(4096, ex["blen"] + 2)))
    rs = t.scalar_p(lw, up)
    o = O.trcbk_h_ng(pp, ex)
    if rs["score"] != o["score"] or not np.array_equal(rs["skl"], o["skl"]):
        bad += 1
        if bad < 6:
            print("MISMATCH", i, len(q), len(g), t.info(), rs["score"], o["score"], len(rs["skl"]), len(o["skl"]),
                  rs["skl"][:4].tolist(), o["skl"][:4].tolist())
    t.close()
print(f"{OPTS}: {bad} mismatches in {N} problems")
sys.exit(1 if bad else 0)
