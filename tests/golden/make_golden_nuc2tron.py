#!/usr/bin/env python
"""Generate tests/golden/nuc2tron.npz from the UNMODIFIED reference (oracle/_ref): residue codes of
DNA segments as the reference reads them (DNA set-up) and their tron codes after Seq::nuc2tron
(protein set-up), plus the genetic code table.  Two child processes: the reference keeps one
option string per process.   python tests/golden/make_golden_nuc2tron.py"""
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))


def segments():
    rng = np.random.default_rng(77)
    out = ["ACGTAACCGGTTRACGTYACGTNNACGTMACGTKACGTSACGTWACGTBACGTDACGTHACGTVACGT" * 2, "TAATAGTGA" * 8, "ACGTTGCAAGTCCGATGCATGCAAGTCGATCGATGCTAGCTAGCATCGATCGACT"]
    for n in (50, 333, 1000, 4097):
        s = "".join(rng.choice(list("ACGT"), size=n))
        k = int(rng.integers(5, n - 5))
        out.append(s[:k] + "NRY" + s[k + 3:])
        out.append(s)
    return out


def child(mode):
    import ref_harness as R
    ref = R.Reference("-Q0 -A2 -yX0 -TDictyost" if mode == "prot" else "-Q0 -A2 -S1 -yX0 -TDictyost",
                      protein=(mode == "prot"))
    res = {}
    for i, g in enumerate(segments()):
        t = ref.task(g, "MKVLAAGIVGLLLAQW" if mode == "prot" else "ACGTACGTACGTAAAA")
        res[f"s{i}"] = (t.export_p() if mode == "prot" else t.export())["b"]
        t.close()
    if mode == "prot":
        res["gencode"] = ref.gencode()
    np.savez(HERE / f"_nuc2tron_{mode}.npz", **res)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
        sys.exit(0)
    for mode in ("dna", "prot"):
        subprocess.run([sys.executable, __file__, mode], check=True)
    d = np.load(HERE / "_nuc2tron_dna.npz")
    p = np.load(HERE / "_nuc2tron_prot.npz")
    out = {"gencode": p["gencode"], "n": np.int32(len(segments()))}
    for i in range(len(segments())):
        out[f"dna{i}"] = d[f"s{i}"]         # at(-1 .. len)
        out[f"tron{i}"] = p[f"s{i}"]
    np.savez_compressed(HERE / "nuc2tron.npz", **out)
    (HERE / "_nuc2tron_dna.npz").unlink()
    (HERE / "_nuc2tron_prot.npz").unlink()
    print("nuc2tron.npz:", len(segments()), "segments")
