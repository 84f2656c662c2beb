#!/usr/bin/env python
"""GPU-box tool: run-to-run spread of the end-to-end submit (host buffers -> results) on config 2"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import bench
from spaln_b200 import Engine
prm = bench.load_params()
raw = bench.make_workload(10000, 20251017)
P = bench.to_problems(raw)
eng = Engine(prm, 0)
eng.upload(P); eng.run()
packed = eng.pack(P)
eng.submit(P[:200])
for i in range(8):
    t0 = time.perf_counter()
    eng.submit_packed(packed)
    dt = time.perf_counter() - t0
    tm = eng.timing()
    print(f"submit {i}: total {1e3 * dt:.1f} ms, kernel span {tm.kernel_ms:.1f} ms, h2d(first chunk) {tm.h2d_ms:.2f}, d2h {tm.d2h_ms:.2f}", flush=True)
