#!/usr/bin/env python
"""GPU-box tool: the pieces of one e2e step of bench.py at N = 1 (descriptor marshalling, the C-ABI
submit with host buffers, corner extraction, hit records)"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402
from spaln_b200 import Engine, shard, workload  # noqa: E402

prm = bench.load_params()
G = workload.config2_global(10000, bench.SEED, procs=16)
eng = Engine(prm, 0)
part = np.arange(10000)
qlen = G["lens"][:, 0]
prev = None
for i in range(6):
    t0 = time.perf_counter()
    packed = eng.pack_global(G, part, reuse=prev)
    t1 = time.perf_counter()
    eng.submit_packed(packed)
    t2 = time.perf_counter()
    n_skl = np.minimum(packed.res["n_skl"][:packed.n], np.diff(packed.off)).astype(np.int64)
    lens_c = np.repeat(np.cumsum(n_skl) - n_skl, n_skl)
    src = np.repeat(packed.off[:-1], n_skl) + (np.arange(int(n_skl.sum())) - lens_c)
    corners = packed.skl[src]
    t3 = time.perf_counter()
    hits = shard.make_hits(part, packed.scores, n_skl, qlen, corners, min_intron=int(prm["llmt"]))
    got = shard.gather_hit_records(hits, corners, 0, None)
    t4 = time.perf_counter()
    tm = eng.timing()
    prev = packed
    print(f"step {i}: pack_global {1e3 * (t1 - t0):.1f} ms, submit_packed {1e3 * (t2 - t1):.1f} (kernel span {tm.kernel_ms:.1f}), "
          f"corners {1e3 * (t3 - t2):.1f}, hits + gather {1e3 * (t4 - t3):.1f}", flush=True)
