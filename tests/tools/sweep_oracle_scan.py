#!/usr/bin/env python
"""Build-container tool: pins the splice-signal scan restatement (oracle/spaln_oracle_scan.c) against
the Exinon tables of the unmodified reference.  usage: sweep_oracle_scan.py [n] [seed] [options]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_harness as O          # noqa: E402
import ref_harness as R             # noqa: E402
from spaln_b200 import workload as synth    # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
SEED = int(sys.argv[2]) if len(sys.argv) > 2 else 3
OPTS = sys.argv[3] if len(sys.argv) > 3 else "-Q0 -A2 -S1 -yX0 -TDictyost"


def scan_params(ref, t):
    p = {}
    for w, name in ((0, "pat5"), (1, "pat3")):
        pm = ref.patmat(w)
        p[name + "_meta"] = np.array([pm["rows"], pm["cols"], pm["offset"], pm["nalpha"], pm["morder"]], np.int32)
        p[name + "_f"] = np.array([pm["tonic"], pm["min_elem"]], np.float32)
        p[name + "_mtx"] = pm["mtx"]
    f = t.scan_factors()
    assert f["cmpc"] == 0 and f["many"] == 1
    p["scan_f"] = np.array([f["fS"], f["sss"]], np.float32)
    p["any"] = f["any"]
    p["sig53tab"] = t.export_ng_tables(64)["sig53tab"]
    return p


if __name__ == "__main__":
    ref = R.Reference(OPTS)
    rng = np.random.default_rng(SEED)
    bad = 0
    for i in range(N):
        g, q, _ = synth.plant_gene(rng, qlen_range=(20, 400), flank=[(0, 5), (10, 60), (100, 900)][i % 3])
        if i % 4 == 0 and len(g) > 60:      # ambiguity codes
            k = int(rng.integers(0, len(g) - 8))
            g = g[:k] + "NRYN"[: int(rng.integers(1, 5))] + g[k + 4:]
        t = ref.task(g, q)
        ex = t.export()
        i53 = t.export_int53()
        o = O.exinon_scan(scan_params(ref, t), ex["b"][1:-1])
        L = len(g)
        # the reference never writes dinc3 of column 0 nor dinc5 of column len - 1 (uninitialised
        # INT53 entries): sig3[0] and sig5[len - 1] are not reproducible
        ok = (np.array_equal(o["sig5"][:L - 1], ex["sig5"][:L - 1]) and np.array_equal(o["sig3"][1:L], ex["sig3"][1:L])
              and np.array_equal(o["int53"][0:L - 1] & 0x0f0f, i53[0:L - 1] & 0x0f0f)
              and np.array_equal(o["int53"][1:L + 1] & 0xf0f0, i53[1:L + 1] & 0xf0f0))
        if not ok:
            bad += 1
            if bad < 4:
                d5 = np.nonzero(o["sig5"][:L - 1] != ex["sig5"][:L - 1])[0]
                d3 = np.nonzero(o["sig3"][1:L] != ex["sig3"][1:L])[0] + 1
                print("MISMATCH", i, L, d5[:6], d3[:6])
        t.close()
    print(f"{OPTS}: {bad} mismatches in {N} segments")
    sys.exit(1 if bad else 0)
