#!/usr/bin/env python
"""GPU-box tool: the config-4 leg of bench.py alone (mRNA x loci with 20x introns, -LS, lspS_ng at
-V 32 MiB: multi-intermediate Hirschberg route + block re-alignments).
usage: [GSPALN_LSP_DEBUG=1] quick_config4.py [queries (the leg takes a fifth)]"""
import json
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402

args = types.SimpleNamespace(queries=int(sys.argv[1]) if len(sys.argv) > 1 else 10000)
out = bench.config4_leg(args, 16, with_cpu=False)
print(json.dumps({k: out[k] for k in ("queries", "queries_per_s", "gcups_root_cells", "kernel_ms", "total_ms",
                                      "launches", "device_cells", "status_nonzero")}))
