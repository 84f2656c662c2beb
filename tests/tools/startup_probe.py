#!/usr/bin/env python
"""GPU-box tool: where the first call of a fresh process spends its time (CUDA context, engine
creation, exact-ILD tables, first submit, second submit)."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
t0 = time.perf_counter()
import numpy as np  # noqa: E402
import golden_io  # noqa: E402
from spaln_b200 import Engine, Problem, capi  # noqa: E402
t1 = time.perf_counter()
prm, probs = golden_io.load("dna_A2_udh")
lib = capi.load()
t2 = time.perf_counter()
import ctypes as C  # noqa: E402
cudart = C.CDLL("libcudart.so")
cudart.cudaFree(0)
t3 = time.perf_counter()
eng = Engine(prm, device=0)
t4 = time.perf_counter()
P = [Problem.from_export(pb, pb["lw"], pb["up"]) for pb in probs[:4]]
r = eng.lspS_ng(P, max_vmf_space=32 << 20, sh=100)
t5 = time.perf_counter()
r = eng.lspS_ng(P, max_vmf_space=32 << 20, sh=100)
t6 = time.perf_counter()
r = eng.forwardS_ng(P)
t7 = time.perf_counter()
r = eng.forwardS_ng(P)
t8 = time.perf_counter()
print(f"imports {t1 - t0:.2f} s, load lib {t2 - t1:.2f}, CUDA context {t3 - t2:.2f}, Engine() {t4 - t3:.2f}, "
      f"first lsp {1e3 * (t5 - t4):.1f} ms, second lsp {1e3 * (t6 - t5):.1f} ms, first exact {1e3 * (t7 - t6):.1f} ms, "
      f"second exact {1e3 * (t8 - t7):.1f} ms")
