#!/usr/bin/env python
"""GPU-box tool: throughput of the exact-ILD kernels (forwardS_ng with path records, scorealoneS_ng)
on config-2 problems.  usage: quick_xild.py [n_queries] [take]   (take: keep the `take` problems
closest to the median size; a third argument "rows,cols" crops every problem to that many query
rows / genome columns -- a short kernel for a profiler run)"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench            # noqa: E402
from spaln_b200 import Engine, capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
prm = bench.load_params()
raw = bench.make_workload(n, 20251017)
bench.host_cells(raw)
if len(sys.argv) > 2:
    raw.sort(key=lambda r: r["cells"])
    k = int(sys.argv[2])
    raw = raw[(n - k) // 2: (n - k) // 2 + k]
    n = k
if len(sys.argv) > 3:
    from spaln_b200 import workload
    rows, colsn = (int(x) for x in sys.argv[3].split(","))
    for r in raw:
        r["a_right"] = min(r["a_right"], rows)
        r["b_right"] = min(r["b_right"], colsn)
        r["lw"], r["up"] = workload.stripe(r["a_left"], r["a_right"], r["b_left"], r["b_right"], 100)
    bench.host_cells(raw)
probs = bench.to_problems(raw)
cells = sum(r["cells"] for r in raw)
eng = Engine(prm, device=0)
for kind, name in ((capi.FORWARD_NG, "forwardS_ng"), (capi.SCOREALONE_NG, "scorealoneS_ng")):
    eng.upload(probs, kind=kind)
    eng.run()
    ks = []
    for _ in range(2):
        eng.run()
        ks.append(eng.timing().kernel_ms)
    res = eng.download()
    bad = sum(1 for r in res if r.status != 0)
    print(f"{name}: {n} queries {cells / 1e9:.2f} Gcells  {np.mean(ks):.1f} ms = {cells / np.mean(ks) / 1e6:.2f} GCUPS  "
          f"status!=0 {bad}", flush=True)
eng.close()
