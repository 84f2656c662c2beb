#!/usr/bin/env python
"""Build-container tool: pins the protein-side scan restatement (so_exinon_scan_p: intron53_p,
calcScr_3, PSSMs on tron codes) against the SGPT6 tables of the unmodified reference.
usage: sweep_oracle_scan_p.py [n] [seed] [options]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_harness as O          # noqa: E402
import ref_harness as R             # noqa: E402
from spaln_b200 import workload as synth    # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 60
SEED = int(sys.argv[2]) if len(sys.argv) > 2 else 3
OPTS = sys.argv[3] if len(sys.argv) > 3 else "-Q0 -A2 -yX0 -TDictyost"
NAMES = ["sig5", "sig3", "sigS", "sigT", "sigE", "sigI", "phs5", "phs3"]


def scan_params_p(ref, t):
    p = {}
    for w, name in ((0, "pat5"), (1, "pat3"), (2, "patI"), (3, "patT")):
        pm = ref.patmat(w)
        if pm is None:
            continue
        p[name + "_meta"] = np.array([pm["rows"], pm["cols"], pm["offset"], pm["nalpha"], pm["morder"]], np.int32)
        p[name + "_f"] = np.array([pm["tonic"], pm["min_elem"]], np.float32)
        p[name + "_mtx"] = pm["mtx"]
    f = t.scan_factors()
    fp = t.scan_factors_p()
    assert ref.patmat(4) is None and not fp["exonpot"] and not fp["intnpot"] and f["many"] == 1
    p["scan_f"] = np.array([f["fS"], f["sss"]], np.float32)
    p["scan_fp"] = np.array([fp["fact"], fp["z"], fp["bti"], fp["o"]], np.float32)
    p["any"] = f["any"]
    p["sig53tab"] = t.export_ng_tables(64)["sig53tab"]
    p["codepot"] = ref.codepot()
    return p


if __name__ == "__main__":
    ref = R.Reference(OPTS, protein=True)
    rng = np.random.default_rng(SEED)
    bad = 0
    for i in range(N):
        g, q, _ = synth.plant_protein_gene(rng, plen_range=(30, 200), flank=[(30, 60), (60, 300), (200, 900)][i % 3])
        if i % 4 == 0:
            k = int(rng.integers(10, len(g) - 12))
            g = g[:k] + "NNRY"[: int(rng.integers(1, 5))] + g[k + 4:]
        t = ref.task(g, q)
        ex = t.export_p()
        o = O.exinon_scan_p(scan_params_p(ref, t), ex["b"][1:-1])
        L = len(g)
        got, want = o["sgpt6"], ex["sgpt6"]
        for c, nm in enumerate(NAMES):
            # entries that depend on the INT53 halves the reference never writes (uninitialised
            # memory: dinc3 / cano3 of column 0, dinc5 / cano5 of column len - 1) are not reproducible
            lo, hi = {"sig3": (1, L), "sig5": (0, L - 1), "phs3": (2, L), "phs5": (0, L - 2)}.get(nm, (0, L))
            d = np.nonzero(got[lo:hi, c] != want[lo:hi, c])[0] + lo
            if len(d):
                bad += 1
                if bad < 8:
                    print("MISMATCH", i, L, nm, len(d), d[:6], got[d[:6], c], want[d[:6], c])
        t.close()
    print(f"{OPTS}: {bad} mismatching columns in {N} segments")
    sys.exit(1 if bad else 0)
