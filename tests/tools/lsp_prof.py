#!/usr/bin/env python
"""GPU-box tool: phase times of the lsp driver on config 2 (set GSPALN_LSP_DEBUG=1)"""
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
eng.lspS_ng(P[:200])
pk = eng.pack(P)
for _ in range(3):
    t0 = time.perf_counter()
    eng.lsp_packed(pk, max_vmf_space=32 * 1024 * 1024, sh=100, alg=2)
    print("lsp_packed total ms", 1e3 * (time.perf_counter() - t0), flush=True)
