#!/usr/bin/env python
"""GPU-box tool: DNA forward kernel throughput on the config-2 workload + parity spot check
against the oracle.  usage: [GSPALN_LIB=...] quick_dna.py [n_queries] [n_check]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench            # noqa: E402
import oracle_harness as O  # noqa: E402
from spaln_b200 import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ncheck = int(sys.argv[2]) if len(sys.argv) > 2 else 6
prm = bench.load_params()
raw = bench.make_workload(n, 20251017)
bench.host_cells(raw)
probs = bench.to_problems(raw)
cells = sum(r["cells"] for r in raw)
eng = Engine(prm, device=0)
eng.upload(probs)
eng.run()
ks = []
for _ in range(3):
    eng.run()
    ks.append(eng.timing().kernel_ms)
res = eng.download()
bad = sum(1 for r in res if r.status != 0)
for i in range(ncheck):
    t = dict(raw[i])
    t["a"] = np.concatenate([[0], raw[i]["a"], [0]]).astype(np.uint8)
    t["b"] = np.concatenate([[0], raw[i]["b"], [0]]).astype(np.uint8)
    t.update(a_exgl=1, a_exgr=1, b_exgl=1, b_exgr=1)
    o = O.forward_wip(prm, t)
    if o["score"] != res[i].score or not np.array_equal(o["skl"], res[i].skl):
        bad += 1
eng.upload(probs, kind=1)
eng.run()
eng.run()
so = eng.timing().kernel_ms
print(f"{n} queries {cells / 1e9:.1f} Gcells  trace {np.mean(ks):.1f} ms = {cells / np.mean(ks) / 1e6:.1f} GCUPS   "
      f"score-only {so:.1f} ms = {cells / so / 1e6:.1f} GCUPS   bad {bad}", flush=True)
eng.close()
