#!/usr/bin/env python
"""GPU-box tool: the reference's default mode -A0 for protein queries on config-3 shaped problems:
Aln2h1::lspH_ng with alg 0 (forwardH_ng / hirschbergH_ng kernels) next to -A2 (forwardH1_wip).
usage: quick_a0_protein.py [n_problems]"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import golden_io  # noqa: E402
from spaln_b200 import EngineH, ProblemH, capi, workload  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
rng = np.random.default_rng(20251017 + 3)
raw = []
for _ in range(n):
    r = workload.protein_problem(rng, plen_range=(300, 800), flank=(500, 5000), sh=100)
    r["int53"] = workload.synthetic_int53(workload.encode_dna(r["genome"]))
    raw.append(r)
lib = capi.load()
import ctypes as C  # noqa: E402
cells = 0
for r in raw:
    t = capi.GspalnHTask()
    t.a_left, t.a_right, t.b_left, t.b_right, t.lw, t.up = r["a_left"], r["a_right"], r["b_left"], r["b_right"], r["lw"], r["up"]
    cells += int(lib.gspaln_h_task_cells(C.byref(t)))
P = [ProblemH.from_export(r, r["lw"], r["up"]) for r in raw]
for name, alg in (("prot_A0_udh", 0), ("prot_A2_global", 2)):
    prm, _ = golden_io.load_protein(name)
    eng = EngineH(prm, device=0)
    eng.lspH_ng(P[:32], max_vmf_space=32 << 20, sh=100, alg=alg)
    t0 = time.perf_counter()
    res = eng.lspH_ng(P, max_vmf_space=32 << 20, sh=100, alg=alg)
    dt = time.perf_counter() - t0
    bad = sum(1 for r in res if r.status != 0)
    print(f"-A{alg}: {n} proteins, {cells / 1e9:.2f} Gcells, {1e3 * dt:.0f} ms wall (host buffers, Python marshalling included) = "
          f"{n / dt:.0f} queries/s, {cells / dt / 1e9:.2f} GCUPS; kernels {eng.timing().kernel_ms:.0f} ms; status != 0: {bad}", flush=True)
    eng.close()
