#!/usr/bin/env python
"""GPU-box tool: protein x genome kernel throughput on the bench workload + a parity spot check
against the oracle.  usage: [GSPALN_LIB=...] quick_protein.py [n_problems] [n_check]"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench            # noqa: E402
import golden_io        # noqa: E402
import oracle_harness as O  # noqa: E402
from spaln_b200 import EngineH  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
ncheck = int(sys.argv[2]) if len(sys.argv) > 2 else 12
for fixture in ("prot_A2_global", "prot_A2_local"):
    prm, _ = golden_io.load_protein(fixture)
    raw = bench.make_protein_workload(n, 20251017)
    bench.host_cells_h(raw)
    probs = bench.to_problems_h(raw)
    cells = sum(r["cells"] for r in raw)
    eng = EngineH(prm, device=0)
    eng.upload(probs)
    eng.run()
    ks = []
    for _ in range(3):
        eng.run()
        ks.append(eng.timing().kernel_ms)
    res = eng.download()
    bad = 0
    for i in range(ncheck):
        o = O.forward_h1_wip(prm, raw[i])
        if o["score"] != res[i].score or not np.array_equal(o["skl"], res[i].skl):
            bad += 1
    eng.upload(probs, kind=1)
    eng.run()
    eng.run()
    so_ms = eng.timing().kernel_ms
    print(f"{fixture}: {n} problems {cells / 1e9:.2f} Gcells  trace {np.mean(ks):.1f} ms = "
          f"{cells / np.mean(ks) / 1e6:.1f} GCUPS   score-only {so_ms:.1f} ms = {cells / so_ms / 1e6:.1f} GCUPS   "
          f"parity mismatches {bad}/{ncheck}", flush=True)
    eng.close()
