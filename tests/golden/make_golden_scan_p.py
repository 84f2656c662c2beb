#!/usr/bin/env python
"""Generate tests/golden/scan_p.npz from the UNMODIFIED reference (oracle/_ref): parameters of the
protein-side scan (Exinon::intron53_p: four PSSMs, coding potential table, factors) and, for a few
TRON segments, the SGPT6 table + INT53 array the reference builds.
    python tests/golden/make_golden_scan_p.py"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE.parent / "tools"))

if __name__ == "__main__":
    import ref_harness as R
    from spaln_b200 import workload as synth
    from sweep_oracle_scan_p import scan_params_p

    ref = R.Reference("-Q0 -A2 -yX0 -TDictyost", protein=True)
    rng = np.random.default_rng(31)
    out = {}
    n = 0
    for i in range(8):
        g, q, _ = synth.plant_protein_gene(rng, plen_range=(30, 250), flank=[(30, 60), (100, 400), (300, 1500)][i % 3])
        if i % 2 == 0:
            k = int(rng.integers(10, len(g) - 12))
            g = g[:k] + "NNRY"[: int(rng.integers(1, 5))] + g[k + 4:]
        t = ref.task(g, q)
        ex = t.export_p()
        if n == 0:
            for k2, v in scan_params_p(ref, t).items():
                out["prm_" + k2] = np.asarray(v)
            out["prm_sig53tab"] = t.export_ng_tables(64)["sig53tab"]
        out[f"s{n}_tron"] = ex["b"]             # at(-1 .. len)
        out[f"s{n}_sgpt6"] = ex["sgpt6"]
        out[f"s{n}_int53"] = t.export_int53()
        n += 1
        t.close()
    out["n"] = np.int32(n)
    np.savez_compressed(HERE / "scan_p.npz", **out)
    print("scan_p.npz:", n, "segments,", (HERE / "scan_p.npz").stat().st_size // 1024, "KiB")
