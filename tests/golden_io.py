"""Reader of tests/golden/*.npz (written by tests/golden/make_golden.py)."""
from pathlib import Path

import numpy as np

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"
GOLDEN_NAMES = ["dna_A2_global", "dna_A2_local", "dna_A3_global", "dna_A2_tetrapod"]
DAGP_NAMES = ["dna_A2_dagp"]        # double affine gaps (-yl3)
UDH_NAMES = ["dna_A2_udh", "dna_A2_udh_local", "dna_A6_udh_recursive"]
A0_NAMES = ["dna_A0_udh", "dna_A0_udh_local", "dna_A0_udh_dagp"]   # -A0: scalar kernels + hirschbergS_ng
SUDH_KEYS = tuple(f"sudh{nn}_{k}" for nn in (1, 2, 5) for k in ("nim", "intvl", "score", "cpos", "ranges"))
CIP_NAMES = ["dna_A2_cip"]          # queries annotated with intron positions (Cip_score)
PROTEIN_CIP_NAMES = ["prot_A2_cip"]
PROTEIN_A0_NAMES = ["prot_A0_udh", "prot_A0_udh_local"]     # -A0: forwardH_ng + hirschbergH_ng
GEOM_KEYS = ["a_left", "a_right", "b_left", "b_right", "a_exgl", "a_exgr", "b_exgl", "b_exgr",
             "lw", "up"]


def load(name):
    z = np.load(GOLDEN_DIR / f"{name}.npz")
    prm = {}
    for k in z.files:
        if k.startswith("prm_"):
            v = z[k]
            prm[k[4:]] = v if v.ndim else int(v)
    probs = []
    for i in range(int(z["n"])):
        pre = f"p{i}_"
        d = {"a": z[pre + "a"], "b": z[pre + "b"], "sig5": z[pre + "sig5"], "sig3": z[pre + "sig3"],
             "score": int(z[pre + "score"]), "skl": z[pre + "skl"],
             "score_only": int(z[pre + "score_only"]), "tag": str(z[pre + "tag"])}
        d.update({k: int(v) for k, v in zip(GEOM_KEYS, z[pre + "geom"])})
        for k in ("lsp_score", "lsp_skl", "udh_nim", "udh_score", "udh_cpos", "udh_ranges",
                  "int53", "ng_score", "ng_skl", "ng_score_only", "cip") + SUDH_KEYS:
            if pre + k in z.files:
                v = z[pre + k]
                d[k] = v if v.ndim else int(v)
        probs.append(d)
    return prm, probs


PROTEIN_NAMES = ["prot_A2_global", "prot_A2_local"]
PROTEIN_UDH_NAMES = ["prot_A2_udh", "prot_A2_udh_local", "prot_A6_udh_recursive"]
GEOM_KEYS_P = GEOM_KEYS + ["blen", "alen"]


def load_protein(name):
    z = np.load(GOLDEN_DIR / f"{name}.npz")
    prm = {}
    for k in z.files:
        if k.startswith("prm_"):
            v = z[k]
            prm[k[4:]] = v if v.ndim else int(v)
    probs = []
    for i in range(int(z["n"])):
        pre = f"p{i}_"
        d = {"a": z[pre + "a"], "b": z[pre + "b"], "sgpt6": z[pre + "sgpt6"],
             "score": int(z[pre + "score"]), "skl": z[pre + "skl"],
             "score_only": int(z[pre + "score_only"]), "tag": str(z[pre + "tag"])}
        d.update({k: int(v) for k, v in zip(GEOM_KEYS_P, z[pre + "geom"])})
        d.setdefault("alen", len(d["a"]) - 2)
        for k in ("lsp_score", "lsp_skl", "udh_nim", "udh_score", "udh_cpos", "udh_ranges",
                  "int53", "ng_score", "ng_skl", "cip") + SUDH_KEYS:
            if pre + k in z.files:
                v = z[pre + k]
                d[k] = v if v.ndim else int(v)
        probs.append(d)
    return prm, probs
