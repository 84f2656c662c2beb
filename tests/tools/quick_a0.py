#!/usr/bin/env python
"""GPU-box tool: the -A0 leg of bench.py alone (Aln2s1::lspS_ng with alg 0 on config-2 problems).
usage: [GSPALN_NG_REC_EIGHTHS=k] [GSPALN_LSP_DEBUG=1] quick_a0.py [queries (the leg takes a tenth)]"""
import json
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402

args = types.SimpleNamespace(queries=int(sys.argv[1]) if len(sys.argv) > 1 else 10000)
out = bench.a0_leg(args, 16, with_cpu=False)
print(json.dumps({k: out[k] for k in ("queries", "queries_per_s", "gcups_root_cells", "kernel_ms", "total_ms",
                                      "launches", "status_nonzero")}))
