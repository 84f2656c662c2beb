#!/usr/bin/env python
"""Build-container tool: protein UDH oracle (so_hirschberg_h1_wip) against the live reference
on seeded random problems.  usage: sweep_oracle_protein_udh.py [-LS] [n] [seed]"""
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
    verbose = "-v" in sys.argv
    n = int(args[0]) if args else 60
    seed = int(args[1]) if len(args) > 1 else 1
    opts = "-Q0 -A2 -yX0 -TDictyost" + (" -LS" if local else "")
    ref = R.Reference(opts, protein=True)
    p = ref.params()
    rng = np.random.default_rng(seed)
    bad = 0
    for i in range(n):
        pl = int(rng.integers(40, 500 if i % 5 == 0 else 220))
        g, q, _ = synth.plant_protein_gene(rng, plen_range=(pl, pl), flank=(30, 400))
        t = ref.task(g, q)
        kw = {}
        if i % 3 == 1 and not local:
            er = [(1, 1), (0, 0), (1, 0)][int(rng.integers(0, 3))]
            kw = dict(a_exgl=int(rng.integers(0, 2)), a_exgr=er[0],
                      b_exgl=int(rng.integers(0, 2)), b_exgr=er[1])
        if i % 7 == 3:
            kw.update(a_left=int(rng.integers(0, 10)), a_right=len(q) - int(rng.integers(0, 10)),
                      b_left=int(rng.integers(0, 40)), b_right=len(g) - int(rng.integers(0, 40)))
        if kw:
            t.set(**kw)
        lw, up = t.stripe31(p["sh"])
        ex = t.export_p()
        ex.update(lw=lw, up=up)
        m = ex["a_right"] - ex["a_left"]
        n_im = int(rng.integers(1, max(2, min(8, m // 16))))
        if verbose:
            print("case", i, pl, kw, lw, up, n_im, flush=True)
        r = t.udh_p(lw, up, n_im)
        o = O.hirschberg_h1_wip(p, ex, n_im)
        ok = r["score"] == o["score"] and np.array_equal(r["cpos"][:n_im + 1, :8], o["cpos"][:n_im + 1, :8]) \
            and r["ranges"] == o["ranges"]
        if not ok:
            bad += 1
            print("MISMATCH", i, pl, kw, n_im, r["score"], o["score"], r["ranges"], o["ranges"])
            if verbose:
                print(r["cpos"][:, :8], o["cpos"][:, :8], sep="\n")
        t.close()
    print(f"{opts}: {n} problems, {bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
