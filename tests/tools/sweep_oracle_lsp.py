#!/usr/bin/env python
"""Build-container tool: pins the driver restatements (oracle/spaln_oracle.c so_lsp,
oracle/spaln_oracle_h.c so_lsp_h) against the unmodified reference's Aln2s1::lspS_ng
(src/fwd2s1.cc:1801-1897) / Aln2h1::lspH_ng (src/fwd2h1.cc:2134-2230) under ANY option string --
in particular the default mode -A0 (hexagonal volume in the dispatch, scalar Hirschberg passes,
blocks banded by the bounds the pass records) at a small -V, where every problem of some size
takes the Hirschberg route and its post-work.
With `cip` as the fifth argument every query is annotated with intron positions (a `;B` block:
Cip_score, src/gsinfo.h:36-139), so the acceptor bonus runs through the kernels and the driver.
usage: sweep_oracle_lsp.py dna|prot [n] [seed] [reference options] [cip]   (one option string per process)"""
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
OPTS = sys.argv[4] if len(sys.argv) > 4 else ("-Q0 -A0 -yX0 -V64K -TDictyost" if PROT else "-Q0 -A0 -S1 -yX0 -V64K -TDictyost")
CIP = len(sys.argv) > 5 and sys.argv[5] == "cip"
FLAGS = [(1, 1, 1, 1), (0, 0, 0, 0), (1, 0, 1, 0), (0, 1, 0, 1), (1, 1, 0, 0), (0, 0, 1, 1), (1, 0, 0, 0)]

ref = R.Reference(OPTS, protein=PROT)
p = ref.params()
if PROT:
    p.update(ref.scalar_p_tables())
rng = np.random.default_rng(SEED)
bad = unsup = 0
for i in range(N):
    kind = i % 3
    if PROT:
        g, q, truth = synth.plant_protein_gene(rng, plen_range=[(8, 40), (30, 150), (120, 400)][kind],
                                               flank=[(20, 200), (30, 400), (40, 300)][kind])
    else:
        g, q, truth = synth.plant_gene(rng, qlen_range=[(8, 60), (40, 300), (250, 900)][kind],
                                       flank=[(10, 200), (20, 500), (50, 800)][kind],
                                       intron_scale=float(rng.choice([0.3, 1.0, 4.0])))
    t = ref.task(g, q)          # (no unrelated pairs: without a path the reference dereferences a null intermediate, src/fwd2s1.cc:1093)
    if i % 4 == 0 and not PROT:
        f = FLAGS[int(rng.integers(0, len(FLAGS)))]
        t.set(a_exgl=f[0], a_exgr=f[1], b_exgl=f[2], b_exgr=f[3])
    if i % 5 == 1 and len(q) > 20:
        t.set(a_left=int(rng.integers(0, 4)), a_right=len(q) - int(rng.integers(0, 4)),
              b_left=int(rng.integers(0, 20)), b_right=len(g) - int(rng.integers(0, 20)))
    cip = None
    if CIP:
        # the true exon boundaries in query coordinates (coding positions for a protein), some of
        # their neighbours and a few random positions, multiplicity 1 .. 3
        step, pos, acc = (3 if PROT else 1), set(), 0
        for (s0, e0) in list(truth)[:-1]:
            acc += e0 - s0
            pos.add(acc + (int(rng.integers(-2, 3)) if rng.random() < 0.3 else 0))
        for _ in range(int(rng.integers(1, 5))):
            pos.add(int(rng.integers(1, max(2, step * len(q)))))
        pos = np.array(sorted(x for x in pos if 0 < x < step * len(q)), np.int32)
        cip = t.set_cip(pos, rng.integers(1, 4, size=len(pos)).astype(np.int32))
    lw, up = (t.stripe31 if PROT else t.stripe)(p["sh"])
    ex = t.export_p() if PROT else t.export()
    ex.update(int53=t.export_int53(), lw=lw, up=up)
    if cip is not None:
        ex["cip"] = cip
    pp = dict(p)
    pp.update(t.export_ng_tables(max(4096, ex["blen"] + 2)))
    o = (O.lsp_h if PROT else O.lsp)(pp, ex, cap=1 << 17)
    if o["unsupported"]:
        unsup += 1
        t.close()
        continue
    r = (t.lsp_p if PROT else t.lsp)(lw, up, cap=1 << 17)
    if r["score"] != o["score"] or not np.array_equal(r["skl"], o["skl"]):
        bad += 1
        if bad < 5:
            print("MISMATCH", i, len(q), len(g), t.info(), r["score"], o["score"], len(r["skl"]), len(o["skl"]))
    t.close()
print(f"{WHAT} {OPTS}{' + Cip_score' if CIP else ''}: {bad} mismatches in {N} problems ({unsup} the oracle calls unsupported)", flush=True)
sys.exit(1 if bad else 0)
