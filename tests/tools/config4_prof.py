#!/usr/bin/env python
"""GPU-box tool: where the time of config-4 shaped problems goes (trace-back kernel alone, Hirschberg
pass alone, whole driver; global vs local parameters).  usage: config4_prof.py [n_problems]"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench                        # noqa: E402
import golden_io                    # noqa: E402
from spaln_b200 import Engine, workload   # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rng = np.random.default_rng(20251017 + 4)
raw = [workload.config2_problem(rng, qlen_range=(1500, 3500), intron_scale=20.0) for _ in range(N)]
bench.host_cells(raw)
cells = np.array([r["cells"] for r in raw], float)
print(f"{N} problems, cells total {cells.sum():.3e} mean {cells.mean():.3e} max {cells.max():.3e}", flush=True)
P = bench.to_problems(raw)
for name in ("dna_A2_global", "dna_A2_local"):
    prm, _ = golden_io.load(name)
    eng = Engine(prm, 0)
    eng.upload(P[:64]); eng.run()
    eng.upload(P)
    eng.run(); eng.run()
    t = eng.timing()
    print(f"{name}: forwardS1_wip kernel {t.kernel_ms:.1f} ms -> {cells.sum() / t.kernel_ms / 1e6:.1f} GCUPS", flush=True)
    # the largest problem alone: what one warp does when it has an SM to itself
    big = int(np.argmax(cells))
    eng.upload([P[big]]); eng.run(); eng.run()
    t = eng.timing()
    print(f"{name}: largest problem alone ({cells[big]:.3e} cells) {t.kernel_ms:.1f} ms -> "
          f"{cells[big] / t.kernel_ms / 1e6:.3f} GCUPS per warp", flush=True)
    eng.close()
    eng = Engine(prm, 0)
    for p in P:
        m = p.a_right - p.a_left
        p.n_imd = max(1, min(int(round((2.0 * m * 2 / 12) ** (1 / 3))) - 1, m // 16))
    from spaln_b200 import capi
    eng.upload(P, capi.HIRSCHBERG_WIP)
    eng.run(); eng.run()
    t = eng.timing()
    print(f"{name}: hirschbergS1_wip kernel {t.kernel_ms:.1f} ms -> {cells.sum() / t.kernel_ms / 1e6:.1f} GCUPS", flush=True)
    t0 = time.perf_counter()
    eng.lspS_ng(P, max_vmf_space=32 << 20, sh=100, alg=2)
    dt = time.perf_counter() - t0
    t = eng.timing()
    print(f"{name}: lspS_ng total {1e3 * dt:.1f} ms, kernels {t.kernel_ms:.1f} ms, device cells {t.cells:.3e}", flush=True)
    eng.close()
