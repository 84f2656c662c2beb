#!/usr/bin/env python
"""GPU-box tool: throughput of the splice-signal scan kernel on a genome-sized segment
(resident input, CUDA-event time of the kernel), next to the C oracle on one host core.
usage: quick_scan.py [megabases] [repeats]"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import golden_io                    # noqa: E402
import oracle_harness as O          # noqa: E402
from spaln_b200 import ExinonScan   # noqa: E402

MB = int(sys.argv[1]) if len(sys.argv) > 1 else 100
REP = int(sys.argv[2]) if len(sys.argv) > 2 else 5
prm, _ = golden_io.load("dna_A2_global")
rng = np.random.default_rng(20251017)
n = MB * 1_000_000
codes = rng.choice(np.array([2, 3, 5, 9], np.uint8), size=n, p=[0.295, 0.205, 0.205, 0.295])
sc = ExinonScan(prm, device=0)
sc.upload(codes)
ms = []
for _ in range(REP + 2):
    sc.run()
    ms.append(sc.timing()["kernel_ms"])
ms = ms[2:]
t0 = time.perf_counter()
got = sc.scan(codes)
e2e = time.perf_counter() - t0
k = min(n, 10_000_000)
t0 = time.perf_counter()
o = O.exinon_scan(prm, codes[:k])
cpu = time.perf_counter() - t0
ok = all(np.array_equal(x[:k - 64], y[:k - 64]) for x, y in zip(got, (o["sig5"], o["sig3"], o["int53"])))
best = float(np.median(ms))
print(json.dumps({"megabases": MB, "kernel_ms": best, "gnt_per_s": n / best / 1e6,
                  "algorithmic_GBps": 7.0 * n / best / 1e6, "e2e_ms_host_buffers": e2e * 1e3,
                  "h2d_d2h_bytes": 7 * n, "cpu_port_1core_mnt_per_s": k / cpu / 1e6,
                  "parity_on_cpu_sample": bool(ok)}))
