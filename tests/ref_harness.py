"""ctypes front-end to oracle/_ref/libspaln_ref.so (the UNMODIFIED reference
compiled by oracle/Makefile).  TEST INFRASTRUCTURE ONLY: used by tests/, by
scripts that generate tests/golden/*, and by bench.py's cpu_baseline /
--impl reference legs.  Never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF_DIR = ROOT / "oracle" / "_ref"
REF_SO = REF_DIR / "libspaln_ref.so"

PARAM_FIELDS = [
    "DvsP", "Noll", "Vab", "Vthr", "BasicGOP", "BasicGEP", "LongGOP", "LongGEP",
    "codonk1", "GapWI", "llmt", "mu", "rlmt", "minl", "maxl", "ild_mode", "nquant",
    "alg", "lcl", "lsg", "any", "qck", "sh", "ubh", "scale", "avmch", "simdim",
    "MaxVmfSpace", "GapPenalty1", "nelem",
]


def available() -> bool:
    return REF_SO.exists() and (REF_DIR / "table" / "gnm2tab").exists()


def write_fasta(path, name, seq):
    with open(path, "w") as f:
        f.write(f">{name}\n")
        for i in range(0, len(seq), 60):
            f.write(seq[i:i + 60] + "\n")


_COMP = str.maketrans("ACGTacgtNn", "TGCAtgcaNn")


def revcomp(s: str) -> str:
    return s.translate(_COMP)[::-1]


_DROPIN = None


def dropin_lib():
    """oracle/_ref/libspaln_dropin.so: the adapter of include/ compiled against the reference
    (links libgspaln; separate from libspaln_ref.so so that the CPU arm never maps the product)"""
    global _DROPIN
    if _DROPIN is None:
        D = C.CDLL(str(REF_DIR / "libspaln_dropin.so"))
        V = C.c_void_p
        D.dropin_s1_adapter.argtypes = [V, V, C.c_int, C.c_int, C.c_int, C.c_int, V, V, C.c_int]
        D.dropin_h1_adapter.argtypes = [V, V, C.c_int, C.c_int, C.c_int, C.c_int, V, V, C.c_int]
        D.dropin_s1_adapter_lsp.argtypes = [V, V, C.c_int, C.c_int, C.c_int, V, V, V, V, C.c_int]
        D.dropin_h1_adapter_lsp.argtypes = [V, V, C.c_int, C.c_int, C.c_int, V, V, V, V, V, C.c_int]
        _DROPIN = D
    return _DROPIN


class Reference:
    """One process-wide reference set-up (the reference keeps its parameters
    in globals, so a process can hold exactly one option string)."""

    _instance = None

    def __init__(self, opts: str = "-Q0 -A2 -S1 -yX0 -TDictyost", protein: bool = False):
        if Reference._instance is not None:
            raise RuntimeError("reference already set up in this process")
        if not available():
            raise RuntimeError("oracle/_ref not built (run make -C oracle ref)")
        os.environ["ALN_TAB"] = str(REF_DIR / "table")
        os.environ.setdefault("ALN_DBS", str(REF_DIR / "seqdb"))
        self.lib = C.CDLL(str(REF_SO))
        L = self.lib
        L.ref_setup.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
        L.ref_task_new.restype = C.c_void_p
        L.ref_task_new.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.ref_task_free.argtypes = [C.c_void_p]
        L.ref_task_info.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_task_set.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_task_export.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.ref_task_inject.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_task_stripe.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_task_kernel.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_void_p]
        L.ref_task_lsp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_void_p]
        for fn in ("ref_pwd", "ref_task_seqs", "ref_task_int53_ptr", "ref_task_sig53tab_ptr"):
            getattr(L, fn).restype = C.c_void_p
        for fn in ("ref_task_seqs", "ref_task_int53_ptr", "ref_task_sig53tab_ptr"):
            getattr(L, fn).argtypes = [C.c_void_p]
        L.ref_task_export_p.argtypes = [C.c_void_p] * 4
        L.ref_get_params_p.argtypes = [C.c_void_p, C.c_int]
        L.ref_task_inject_p.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_task_udh_p.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_task_lsp_p.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_task_kernel_p.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_void_p]
        L.ref_task_stripe31.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_get_params.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.ref_task_export_int53.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_task_export_ng_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_get_patmat.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_task_scan_factors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_task_scalar.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        self.tmp = tempfile.TemporaryDirectory(prefix="spaln_ref_")
        g = os.path.join(self.tmp.name, "g0.fa")
        q = os.path.join(self.tmp.name, "q0.fa")
        write_fasta(g, "g0", "ACGTTGCAAGTCCGATGCATGCAAGTCGATCGATGCTAGCTAGCATCGATCGACTAGCTAGCAT" * 4)
        write_fasta(q, "q0", "MKVLAAGIVGLLLAQWPSEFDHRNTYCMKVLAAGIVGLLLAQWPSEFDHRNTYC" if protein
                    else "ACGTTGCAAGTCCGATGCATGCAAGTCGATCGATGCTAGC")
        rc = L.ref_setup(opts.encode(), g.encode(), q.encode())
        if rc < 0:
            raise RuntimeError(f"ref_setup failed: {rc}")
        self.opts = opts
        self._n = 0
        Reference._instance = self

    # ------------------------------------------------------------------
    def params(self) -> dict:
        buf = np.zeros(128, np.int32)
        sim = np.zeros(64 * 64, np.int32)
        k = self.lib.ref_get_params(buf.ctypes.data, 128, sim.ctypes.data, sim.size)
        nf = len(PARAM_FIELDS)
        p = {name: int(buf[i]) for i, name in enumerate(PARAM_FIELDS)}
        q = buf[nf:k].reshape(-1, 2)
        p["quant_len"] = q[:, 0].astype(np.int32).copy()
        p["quant_pen"] = q[:, 1].astype(np.int32).copy()
        d = p["simdim"]
        p["simmtx"] = sim[: d * d].reshape(d, d).copy()
        if p["DvsP"] == 1:
            pb = np.zeros(16, np.int32)
            self.lib.ref_get_params_p(pb.ctypes.data, 16)
            for i, name in enumerate(["GapW1", "GapW2", "GapW3", "GapW3L", "GapE1", "GapE2", "ExtraGOP",
                                      "codonk1", "termk1", "sim_rows", "sim_cols"]):
                p[name] = int(pb[i])
        return p

    def scalar_p_tables(self):
        """split-codon tables of SpJunc::spjseq + aa2nuc, and minl / ExtraGOP / GapW3L / termk1"""
        tabs = np.zeros(257 * 2 + 64 * 2 + 64 * 2 + 26, np.uint8)
        iv = np.zeros(4, np.int32)
        self.lib.ref_get_scalar_p.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.ref_get_scalar_p(tabs.ctypes.data, iv.ctypes.data)
        return {"spj_tabs": tabs, "minl": int(iv[0]), "ExtraGOP": int(iv[1]), "GapW3L": int(iv[2]), "termk1": int(iv[3])}

    def codepot(self):
        """PwdB::codepot table (ExinPot::begin(), dsize() floats) or None"""
        buf = np.zeros(1 << 16, np.float32)
        self.lib.ref_get_codepot.argtypes = [C.c_void_p, C.c_int]
        n = self.lib.ref_get_codepot(buf.ctypes.data, buf.size)
        return buf[:n].copy() if n > 0 else None

    def gencode(self):
        out = np.zeros(64, np.uint8)
        self.lib.ref_get_gencode.argtypes = [C.c_void_p]
        self.lib.ref_get_gencode(out.ctypes.data)
        return out

    def patmat(self, which: int):
        """EijPat::pattern5 (0) / pattern3 (1): the splice-site PSSM the signal scan applies"""
        meta = np.zeros(8, np.int32)
        fmeta = np.zeros(4, np.float32)
        mtx = np.zeros(1 << 16, np.float32)
        n = self.lib.ref_get_patmat(which, meta.ctypes.data, fmeta.ctypes.data, mtx.ctypes.data, mtx.size)
        if n <= 0:
            return None
        return {"rows": int(meta[0]), "cols": int(meta[1]), "offset": int(meta[2]), "nalpha": int(meta[3]),
                "morder": int(meta[4]), "tonic": np.float32(fmeta[0]), "min_elem": np.float32(fmeta[1]),
                "mtx": mtx[:n].copy()}

    def task(self, genome: str, query: str, comrev_query: bool = False) -> "RefTask":
        self._n += 1
        g = os.path.join(self.tmp.name, f"g{self._n}.fa")
        q = os.path.join(self.tmp.name, f"q{self._n}.fa")
        write_fasta(g, f"g{self._n}", genome)
        write_fasta(q, f"q{self._n}", query)
        h = self.lib.ref_task_new(g.encode(), q.encode(), int(comrev_query))
        os.unlink(g)
        os.unlink(q)
        if not h:
            raise RuntimeError("ref_task_new failed")
        return RefTask(self, h)


class RefTask:
    def __init__(self, ref: Reference, h):
        self.ref, self.h = ref, h
        self.lib = ref.lib

    def close(self):
        if self.h:
            self.lib.ref_task_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict:
        b = np.zeros(10, np.int32)
        self.lib.ref_task_info(self.h, b.ctypes.data)
        keys = ["alen", "blen", "a_left", "a_right", "b_left", "b_right",
                "a_exgl", "a_exgr", "b_exgl", "b_exgr"]
        return {k: int(v) for k, v in zip(keys, b)}

    def set(self, **kw):
        d = self.info()
        d.update(kw)
        keys = ["alen", "blen", "a_left", "a_right", "b_left", "b_right",
                "a_exgl", "a_exgr", "b_exgl", "b_exgr"]
        b = np.array([d[k] for k in keys], np.int32)
        self.lib.ref_task_set(self.h, b.ctypes.data)

    def export(self) -> dict:
        i = self.info()
        a = np.zeros(i["alen"] + 2, np.uint8)
        b = np.zeros(i["blen"] + 2, np.uint8)
        s5 = np.zeros(i["blen"] + 2, np.int16)
        s3 = np.zeros(i["blen"] + 2, np.int16)
        self.lib.ref_task_export(self.h, a.ctypes.data, b.ctypes.data,
                                 s5.ctypes.data, s3.ctypes.data)
        i.update(a=a, b=b, sig5=s5, sig3=s3)
        return i

    def export_p(self) -> dict:
        """protein x genome problem: aa codes, tron codes, SGPT6 table"""
        i = self.info()
        a = np.zeros(i["alen"] + 2, np.uint8)
        b = np.zeros(i["blen"] + 2, np.uint8)
        g = np.zeros((i["blen"] + 2, 8), np.int16)
        self.lib.ref_task_export_p(self.h, a.ctypes.data, b.ctypes.data, g.ctypes.data)
        i.update(a=a, b=b, sgpt6=g)
        return i

    def inject_p(self, sgpt6):
        g = np.ascontiguousarray(sgpt6, np.int16)
        assert g.shape == (self.info()["blen"] + 2, 8)
        self.lib.ref_task_inject_p(self.h, g.ctypes.data)

    def stripe31(self, sh: int):
        b = np.zeros(3, np.int32)
        self.lib.ref_task_stripe31(self.h, sh, b.ctypes.data)
        return int(b[0]), int(b[1])

    def kernel_p(self, lw, up, kind=0, cap=1 << 16):
        score = C.c_int(0)
        secs = C.c_double(0)
        skl = np.zeros((cap, 2), np.int32)
        n = self.lib.ref_task_kernel_p(self.h, lw, up, kind, C.byref(score), skl.ctypes.data, cap,
                                       C.byref(secs))
        return {"score": score.value, "skl": skl[:n].copy(), "seconds": secs.value}

    def udh_p(self, lw, up, n_imd):
        """SimdAln2h1::hirschbergH1_wip; the narrowed ranges are reported, then restored"""
        score = C.c_int(0)
        secs = C.c_double(0)
        cpos = np.zeros((n_imd + 1, 10), np.int32)
        before = self.info()
        self.lib.ref_task_udh_p(self.h, lw, up, n_imd, C.byref(score), cpos.ctypes.data, C.byref(secs))
        after = self.info()
        self.set(**before)
        return {"score": score.value, "cpos": cpos, "seconds": secs.value,
                "ranges": [after["a_left"], after["a_right"], after["b_left"], after["b_right"]]}

    def lsp_p(self, lw, up, cap=1 << 16):
        """Aln2h1::lspH_ng (trace-back vs Hirschberg dispatch + post-work), raw Mfile corners"""
        score = C.c_int(0)
        secs = C.c_double(0)
        skl = np.zeros((cap, 2), np.int32)
        n = self.lib.ref_task_lsp_p(self.h, lw, up, C.byref(score), skl.ctypes.data, cap, C.byref(secs))
        return {"score": score.value, "skl": skl[:n].copy(), "seconds": secs.value}

    def export_int53(self):
        """INT53 nibbles per column (dinc5 | dinc3 << 4 | cano5 << 8 | cano3 << 12)"""
        out = np.zeros(self.info()["blen"] + 2, np.uint16)
        self.lib.ref_task_export_int53(self.h, out.ctypes.data)
        return out

    def export_ng_tables(self, n_pen: int):
        """sig53tab (544 shorts) and IntronPenalty::Penalty(0 .. n_pen - 1)"""
        tab = np.zeros(544, np.int16)
        pen = np.zeros(n_pen, np.int16)
        misc = np.zeros(4, np.int32)
        rc = self.lib.ref_task_export_ng_tables(self.h, tab.ctypes.data, pen.ctypes.data, n_pen,
                                                misc.ctypes.data)
        if rc:
            raise RuntimeError("no intron tables in this set-up")
        return {"sig53tab": tab, "penalty": pen, "intpot": int(misc[0])}

    def scan_factors(self):
        f = np.zeros(4, np.float32)
        i = np.zeros(4, np.int32)
        self.lib.ref_task_scan_factors(self.h, f.ctypes.data, i.ctypes.data)
        return {"fS": np.float32(f[0]), "sss": np.float32(f[1]), "tonic5": np.float32(f[2]),
                "tonic3": np.float32(f[3]), "any": int(i[0]), "cmpc": int(i[1]), "many": int(i[2])}

    def scan_factors_p(self):
        """protein-side scan (Exinon::intron53_p): factors and which potentials exist"""
        f = np.zeros(8, np.float32)
        i = np.zeros(8, np.int32)
        self.lib.ref_task_scan_factors_p.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.ref_task_scan_factors_p(self.h, f.ctypes.data, i.ctypes.data)
        return {"fact": np.float32(f[0]), "z": np.float32(f[1]), "Z": np.float32(f[2]), "bti": np.float32(f[3]),
                "bp_factor": np.float32(f[4]), "o": np.float32(f[5]), "tonicB": np.float32(f[6]),
                "codepot": int(i[0]), "ndata": int(i[1]), "dsize": int(i[2]), "exonpot": int(i[3]),
                "intnpot": int(i[4]), "DvsP": int(i[5]), "maxb3d": int(i[6])}

    def scalar_p(self, lw, up, cap=1 << 16):
        """Aln2h1::trcbkalignH_ng forced onto its scalar branch (forwardH_ng + Vmf)"""
        self.lib.ref_task_scalar_p.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        score = C.c_int(0)
        skl = np.zeros((cap, 2), np.int32)
        n = self.lib.ref_task_scalar_p(self.h, lw, up, C.byref(score), skl.ctypes.data, cap)
        return {"score": score.value, "skl": skl[:n].copy()}

    def set_cip(self, pos, num=None):
        """annotate the query with intron positions (a `;B` / `;b` block, src/gsinfo.h:76-126);
        returns Cip_score::cip_score(c) for every position c the kernels may ask for"""
        pos = np.ascontiguousarray(pos, np.int32)
        num = np.ascontiguousarray(num if num is not None else np.ones(len(pos)), np.int32)
        self.lib.ref_task_set_cip.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        self.lib.ref_task_set_cip.restype = None
        self.lib.ref_task_set_cip(self.h, pos.ctypes.data, num.ctypes.data, len(pos))
        n = 3 * (int(self.info()["alen"]) + 2)
        tab = np.zeros(n, np.int32)
        self.lib.ref_task_cip_table.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        self.lib.ref_task_cip_table.restype = None
        self.lib.ref_task_cip_table(self.h, tab.ctypes.data, n)
        return tab

    def scorealone(self, lw, up):
        """Aln2s1::scorealoneS_ng (scalar score-only kernel)"""
        self.lib.ref_task_scorealone.argtypes = [C.c_void_p, C.c_int, C.c_int]
        return int(self.lib.ref_task_scorealone(self.h, lw, up))

    def scalar(self, lw, up, cap=1 << 16):
        """Aln2s1::trcbkalignS_ng forced onto its scalar branch (forwardS_ng + Vmf), raw Mfile corners"""
        score = C.c_int(0)
        secs = C.c_double(0)
        skl = np.zeros((cap, 2), np.int32)
        n = self.lib.ref_task_scalar(self.h, lw, up, C.byref(score), skl.ctypes.data, cap, C.byref(secs))
        return {"score": score.value, "skl": skl[:n].copy(), "seconds": secs.value}

    def inject(self, sig5, sig3):
        s5 = np.ascontiguousarray(sig5, np.int16)
        s3 = np.ascontiguousarray(sig3, np.int16)
        self.lib.ref_task_inject(self.h, s5.ctypes.data, s3.ctypes.data)

    def stripe(self, sh: int):
        b = np.zeros(3, np.int32)
        self.lib.ref_task_stripe(self.h, sh, b.ctypes.data)
        return int(b[0]), int(b[1])

    def kernel(self, lw, up, kind=0, n_imd=0, mode=2, cap=1 << 16):
        score = C.c_int(0)
        secs = C.c_double(0)
        skl = np.zeros((cap, 2), np.int32)
        cpos = np.zeros((n_imd + 1, 10), np.int32)
        before = self.info()
        n = self.lib.ref_task_kernel(self.h, lw, up, kind, n_imd, mode,
                                     C.byref(score), skl.ctypes.data, cap,
                                     cpos.ctypes.data, C.byref(secs))
        after = self.info()
        if kind == 2:       # the Hirschberg pass narrows the Seq ranges: report, then restore
            self.set(**before)
        return {"score": score.value, "skl": skl[:n].copy(), "cpos": cpos,
                "seconds": secs.value,
                "ranges": [after["a_left"], after["a_right"], after["b_left"], after["b_right"]]}

    def scalar_udh(self, lw, up, n_imd, intvl):
        """Aln2s1::hirschbergS_ng (scalar Hirschberg pass, `-A0`); the narrowed ranges are reported,
        then restored"""
        score = C.c_int(0)
        cpos = np.zeros((n_imd + 1, 10), np.int32)
        before = self.info()
        self.lib.ref_task_scalar_udh.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        self.lib.ref_task_scalar_udh(self.h, lw, up, n_imd, intvl, C.byref(score), cpos.ctypes.data)
        after = self.info()
        self.set(**before)
        return {"score": score.value, "cpos": cpos,
                "ranges": [after["a_left"], after["a_right"], after["b_left"], after["b_right"]]}

    def scalar_udh_p(self, lw, up, n_imd, intvl):
        """Aln2h1::hirschbergH_ng (scalar protein Hirschberg pass, `-A0`); ranges reported, then restored"""
        score = C.c_int(0)
        cpos = np.zeros((n_imd + 1, 10), np.int32)
        before = self.info()
        self.lib.ref_task_scalar_udh_p.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        self.lib.ref_task_scalar_udh_p(self.h, lw, up, n_imd, intvl, C.byref(score), cpos.ctypes.data)
        after = self.info()
        self.set(**before)
        return {"score": score.value, "cpos": cpos,
                "ranges": [after["a_left"], after["a_right"], after["b_left"], after["b_right"]]}

    def adapter(self, lw, up, kind=0, device=0, cap=1 << 16, protein=False):
        """the same problem through include/gspaln_spaln_adapter.hpp (GPU drop-in):
        SpalnEngine::forwardS1_wip / scoreonlyS1_wip, or SpalnEngineH::forwardH1_wip"""
        D = dropin_lib()
        score = C.c_int(0)
        skl = np.zeros((cap, 2), np.int32)
        fn = D.dropin_h1_adapter if protein else D.dropin_s1_adapter
        n = fn(self.lib.ref_task_seqs(self.h), self.lib.ref_pwd(), lw, up, kind, device, C.byref(score),
               skl.ctypes.data, cap)
        return {"score": score.value, "skl": skl[:n].copy()}

    def adapter_lsp(self, lw, up, device=0, cap=1 << 16, protein=False, spj_tabs=None):
        """Aln2s1::lspS_ng / Aln2h1::lspH_ng through include/gspaln_spaln_adapter.hpp (GPU drop-in
        of the driver); returns None if the adapter reports the problem as unsupported"""
        D = dropin_lib()
        score = C.c_int(0)
        skl = np.zeros((cap, 2), np.int32)
        common = (self.lib.ref_task_seqs(self.h), self.lib.ref_pwd(), lw, up, device,
                  self.lib.ref_task_int53_ptr(self.h), self.lib.ref_task_sig53tab_ptr(self.h))
        if protein:
            tabs = np.ascontiguousarray(spj_tabs, np.uint8) if spj_tabs is not None else None
            n = D.dropin_h1_adapter_lsp(*common, tabs.ctypes.data if tabs is not None else None,
                                        C.byref(score), skl.ctypes.data, cap)
        else:
            n = D.dropin_s1_adapter_lsp(*common, C.byref(score), skl.ctypes.data, cap)
        if n < 0:
            return None
        return {"score": score.value, "skl": skl[:n].copy()}

    def lsp(self, lw, up, cap=1 << 16):
        score = C.c_int(0)
        secs = C.c_double(0)
        skl = np.zeros((cap, 2), np.int32)
        n = self.lib.ref_task_lsp(self.h, lw, up, C.byref(score), skl.ctypes.data,
                                  cap, C.byref(secs))
        return {"score": score.value, "skl": skl[:n].copy(), "seconds": secs.value}
