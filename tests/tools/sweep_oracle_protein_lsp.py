#!/usr/bin/env python
"""Build-container tool: protein driver oracle (so_lsp_h) against the live reference's
Aln2h1::lspH_ng.  usage: sweep_oracle_protein_lsp.py [-LS] [-A6] [-V64K] [n] [seed]"""
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
    flags = [a for a in sys.argv[1:] if a.startswith("-") and a != "-v"]
    verbose = "-v" in sys.argv
    n = int(args[0]) if args else 60
    seed = int(args[1]) if len(args) > 1 else 1
    alg = "-A6" if "-A6" in flags else "-A2"
    vopt = next((f for f in flags if f.startswith("-V")), "-V64K")
    opts = f"-Q0 {alg} -yX0 {vopt} -TDictyost" + (" -LS" if "-LS" in flags else "")
    ref = R.Reference(opts, protein=True)
    p = ref.params()
    rng = np.random.default_rng(seed)
    bad = unsup = udh = 0
    for i in range(n):
        pl = int(rng.integers(40, 500 if i % 5 == 0 else 220))
        g, q, _ = synth.plant_protein_gene(rng, plen_range=(pl, pl), flank=(30, 400))
        t = ref.task(g, q)
        kw = {}
        if i % 7 == 3:
            kw.update(a_left=int(rng.integers(0, 10)), a_right=len(q) - int(rng.integers(0, 10)),
                      b_left=int(rng.integers(0, 40)), b_right=len(g) - int(rng.integers(0, 40)))
        if kw:
            t.set(**kw)
        lw, up = t.stripe31(p["sh"])
        ex = t.export_p()
        ex.update(lw=lw, up=up)
        o = O.lsp_h(p, ex)
        if o["unsupported"]:
            unsup += 1
            t.close()
            continue
        m, nn = ex["a_right"] - ex["a_left"], ex["b_right"] - ex["b_left"]
        udh += 2.0 * m * (nn + 3 * m) >= p["MaxVmfSpace"]
        if verbose:
            print("case", i, pl, kw, lw, up, flush=True)
        r = t.lsp_p(lw, up)
        ok = r["score"] == o["score"] and np.array_equal(r["skl"], o["skl"])
        if not ok:
            bad += 1
            print("MISMATCH", i, pl, kw, r["score"], o["score"], len(r["skl"]), len(o["skl"]))
            if verbose:
                print(r["skl"].tolist(), o["skl"].tolist(), sep="\n")
        t.close()
    print(f"{opts}: {n} problems ({udh} through the Hirschberg pass, {unsup} unsupported), {bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
