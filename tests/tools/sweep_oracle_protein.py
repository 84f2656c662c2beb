#!/usr/bin/env python
"""Build-container tool: protein x genome oracle (oracle/spaln_oracle_h.c) against the live
reference (oracle/_ref) on seeded random problems.  usage: sweep_oracle_protein.py [-LS] [n] [seed]"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))
import oracle_harness as O      # noqa: E402
import ref_harness as R         # noqa: E402
from spaln_b200 import workload as synth    # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("-")]
    local = "-LS" in sys.argv
    n = int(args[0]) if args else 100
    seed = int(args[1]) if len(args) > 1 else 1
    opts = "-Q0 -A2 -yX0 -TDictyost" + (" -LS" if local else "")
    ref = R.Reference(opts, protein=True)
    p = ref.params()
    rng = np.random.default_rng(seed)
    bad = 0
    for i in range(n):
        pl = int(rng.integers(8, 700 if i % 10 == 0 else 260))
        g, q, _ = synth.plant_protein_gene(rng, plen_range=(pl, pl), flank=(30, 400))
        t = ref.task(g, q)
        kw = {}
        if i % 3 == 1:
            # (a_exgr, b_exgr) = (0, 1) makes the reference start its walk outside the matrix
            er = [(1, 1), (0, 0), (1, 0)][int(rng.integers(0, 3))]
            kw = dict(a_exgl=int(rng.integers(0, 2)), a_exgr=er[0],
                      b_exgl=int(rng.integers(0, 2)), b_exgr=er[1])
        if i % 7 == 3 and len(q) > 30:
            kw.update(a_left=int(rng.integers(0, 10)), a_right=len(q) - int(rng.integers(0, 10)),
                      b_left=int(rng.integers(0, 40)), b_right=len(g) - int(rng.integers(0, 40)))
        if kw:
            t.set(**kw)
        lw, up = t.stripe31(p["sh"])
        ex = t.export_p()
        ex.update(lw=lw, up=up)
        print("case", i, pl, kw, lw, up, flush=True) if "-v" in sys.argv else None
        r = t.kernel_p(lw, up, 0)
        try:
            o = O.forward_h1_wip(p, ex)
        except RuntimeError as e:
            o = {"score": None, "skl": np.zeros((0, 2), np.int32)}
        ok = r["score"] == o["score"] and np.array_equal(r["skl"], o["skl"])
        if not ok:
            bad += 1
            print("MISMATCH", i, pl, kw, r["score"], o["score"], len(r["skl"]), len(o["skl"]))
        t.close()
    print(f"{opts}: {n} problems, {bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
