"""ctypes front-end to oracle/libspaln_oracle.so (our plain-C restatement of
the reference algorithm).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_SO = ROOT / "oracle" / "libspaln_oracle.so"
MAXQ = 8


class SoParams(C.Structure):
    _fields_ = [
        ("gop", C.c_int32), ("gep", C.c_int32), ("lgop", C.c_int32), ("lgep", C.c_int32),
        ("noll", C.c_int32), ("ipen", C.c_int32), ("llmt", C.c_int32), ("nquant", C.c_int32),
        ("quant_len", C.c_int32 * MAXQ), ("quant_pen", C.c_int32 * MAXQ),
        ("avmch", C.c_int32), ("local", C.c_int32), ("spj", C.c_int32),
        ("simdim", C.c_int32), ("simmtx", C.c_void_p), ("gappen1", C.c_int32),
        ("codonk1", C.c_int32), ("n_penalty", C.c_int32), ("penalty", C.c_void_p),
        ("sig53tab", C.c_void_p),
    ]


class SoTask(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p), ("sig5", C.c_void_p), ("sig3", C.c_void_p),
        ("a_left", C.c_int32), ("a_right", C.c_int32), ("b_left", C.c_int32), ("b_right", C.c_int32),
        ("a_exgl", C.c_int32), ("a_exgr", C.c_int32), ("b_exgl", C.c_int32), ("b_exgr", C.c_int32),
        ("lw", C.c_int32), ("up", C.c_int32), ("int53", C.c_void_p), ("cip", C.c_void_p),
    ]


def build():
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "oracle"], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not ORACLE_SO.exists():
            build()
        _lib = C.CDLL(str(ORACLE_SO))
        _lib.so_forward_wip.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_void_p]
        _lib.so_scoreonly_wip.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.so_hirschberg_wip.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_void_p]
        _lib.so_lsp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_int, C.c_void_p]
        _lib.so_exinon_scan_n.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.so_exinon_scan_n.restype = None
        _lib.so_trcbk_ng.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib.so_forward_h1_wip.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        _lib.so_hirschberg_h1_wip.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.so_lsp_h.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    return _lib


def make_params(p: dict):
    """p: dict in the layout of tests/ref_harness.py::Reference.params() /
    tests/golden params (see spaln_b200.params)."""
    sp = SoParams()
    sp.gop, sp.gep = p["BasicGOP"], p["BasicGEP"]
    sp.lgop, sp.lgep = p["LongGOP"], p["LongGEP"]
    sp.noll = p["Noll"]
    sp.ipen = p["GapWI"]
    sp.llmt = p["llmt"]
    sp.nquant = p["nquant"]
    for j in range(min(p["nquant"], len(p["quant_len"]))):
        sp.quant_len[j] = int(p["quant_len"][j])
        sp.quant_pen[j] = int(p["quant_pen"][j])
    sp.avmch = p["avmch"]
    sp.local = 1 if (p["lcl"] & 16) else 0
    sp.spj = p.get("spj", 1)
    sp.simdim = p["simdim"]
    sim = np.ascontiguousarray(p["simmtx"], np.int32)
    sp.simmtx = sim.ctypes.data
    sp.gappen1 = p["GapPenalty1"]
    sp._keep = [sim]
    # inputs of the scalar exact-ILD kernel (present in fixtures that carry them)
    sp.codonk1 = int(p.get("codonk1", 2 ** 31 - 1))
    if "penalty" in p and "sig53tab" in p:
        pen = np.ascontiguousarray(p["penalty"], np.int16)
        tab = np.ascontiguousarray(p["sig53tab"], np.int16)
        sp.n_penalty, sp.penalty, sp.sig53tab = len(pen), pen.ctypes.data, tab.ctypes.data
        sp._keep += [pen, tab]
    return sp


def make_task(t: dict):
    """t: dict with a, b (uint8 arrays holding codes for at(-1..len)),
    sig5, sig3 (int16, by column), ranges, flags, lw, up."""
    st = SoTask()
    a = np.ascontiguousarray(t["a"], np.uint8)
    b = np.ascontiguousarray(t["b"], np.uint8)
    s5 = np.ascontiguousarray(t["sig5"], np.int16)
    s3 = np.ascontiguousarray(t["sig3"], np.int16)
    st.a = a.ctypes.data + 1      # exported arrays start at at(-1)
    st.b = b.ctypes.data + 1
    st.sig5 = s5.ctypes.data
    st.sig3 = s3.ctypes.data
    for k in ("a_left", "a_right", "b_left", "b_right", "a_exgl", "a_exgr",
              "b_exgl", "b_exgr", "lw", "up"):
        setattr(st, k, int(t[k]))
    st._keep = [a, b, s5, s3]
    if t.get("int53") is not None:
        i53 = np.ascontiguousarray(t["int53"], np.uint16)
        st.int53 = i53.ctypes.data
        st._keep.append(i53)
    if t.get("cip") is not None:
        cip = np.ascontiguousarray(t["cip"], np.int32)
        st.cip = cip.ctypes.data
        st._keep.append(cip)
    return st


def forward_wip(p: dict, t: dict, cap: int = 1 << 16):
    sp, st = make_params(p), make_task(t)
    score = C.c_int32(0)
    cells = C.c_int64(0)
    skl = np.zeros((cap, 2), np.int32)
    n = lib().so_forward_wip(C.byref(sp), C.byref(st), C.byref(score), skl.ctypes.data,
                             cap, C.byref(cells))
    if n < 0:
        raise RuntimeError(f"so_forward_wip failed: {n}")
    return {"score": score.value, "skl": skl[:n].copy(), "cells": cells.value}


def trcbk_ng(p: dict, t: dict, cap: int = 1 << 16):
    """scalar kernel (Aln2s1::trcbkalignS_ng's scalar branch): score + corners"""
    sp, st = make_params(p), make_task(t)
    score = C.c_int32(0)
    skl = np.zeros((cap, 2), np.int32)
    n = lib().so_trcbk_ng(C.byref(sp), C.byref(st), C.byref(score), skl.ctypes.data, cap)
    if n < 0:
        raise RuntimeError(f"so_trcbk_ng failed: {n}")
    return {"score": score.value, "skl": skl[:n].copy()}


def scorealone_ng(p: dict, t: dict):
    """Aln2s1::scorealoneS_ng (scalar score-only kernel)"""
    sp, st = make_params(p), make_task(t)
    score = C.c_int32(0)
    lib().so_scorealone_ng.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    rc = lib().so_scorealone_ng(C.byref(sp), C.byref(st), C.byref(score))
    if rc < 0:
        raise RuntimeError(f"so_scorealone_ng failed: {rc}")
    return {"score": score.value}


def scoreonly_wip(p: dict, t: dict):
    sp, st = make_params(p), make_task(t)
    score = C.c_int32(0)
    rc = lib().so_scoreonly_wip(C.byref(sp), C.byref(st), C.byref(score))
    if rc < 0:
        raise RuntimeError("so_scoreonly_wip failed")
    return {"score": score.value}


def hirschberg_wip(p: dict, t: dict, n_im: int):
    sp, st = make_params(p), make_task(t)
    score = C.c_int32(0)
    cpos = np.zeros((n_im + 1, 10), np.int32)
    ranges = np.zeros(4, np.int32)
    rc = lib().so_hirschberg_wip(C.byref(sp), C.byref(st), n_im, C.byref(score),
                                 cpos.ctypes.data, ranges.ctypes.data)
    if rc < 0:
        raise RuntimeError(f"so_hirschberg_wip failed: {rc}")
    return {"score": score.value, "cpos": cpos, "ranges": ranges.tolist()}


class SoPatMat(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("offset", C.c_int32), ("nalpha", C.c_int32),
                ("morder", C.c_int32), ("tonic", C.c_float), ("min_elem", C.c_float), ("mtx", C.c_void_p)]


class SoScanParams(C.Structure):
    _fields_ = [("pat5", SoPatMat), ("pat3", SoPatMat), ("fS", C.c_float), ("sss", C.c_float),
                ("any", C.c_int32), ("sig53tab", C.c_void_p)]


def hirschberg_ng(p: dict, t: dict, n_im: int, intvl: int):
    """Aln2s1::hirschbergS_ng (scalar Hirschberg pass of `-A0`)"""
    sp, st = make_params(p), make_task(t)
    score = C.c_int32(0)
    cpos = np.zeros((n_im + 1, 10), np.int32)
    ranges = np.zeros(4, np.int32)
    L = lib()
    L.so_hirschberg_ng.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = L.so_hirschberg_ng(C.byref(sp), C.byref(st), n_im, intvl, C.byref(score), cpos.ctypes.data,
                            ranges.ctypes.data)
    if rc < 0:
        raise RuntimeError(f"so_hirschberg_ng failed ({rc})")
    return {"score": score.value, "cpos": cpos, "ranges": ranges.tolist()}


def make_scan_params(p: dict):
    """p: fixture parameters with the splice PSSMs (pat5_* / pat3_*), scan_f = (fS, sss), sig53tab"""
    sp = SoScanParams()
    keep = []
    for name, pm in (("pat5", sp.pat5), ("pat3", sp.pat3)):
        meta = [int(x) for x in p[name + "_meta"]]
        f = np.asarray(p[name + "_f"], np.float32)
        mtx = np.ascontiguousarray(p[name + "_mtx"], np.float32)
        pm.rows, pm.cols, pm.offset, pm.nalpha, pm.morder = meta
        pm.tonic, pm.min_elem = float(f[0]), float(f[1])
        pm.mtx = mtx.ctypes.data
        keep.append(mtx)
    sf = np.asarray(p["scan_f"], np.float32)
    sp.fS, sp.sss = float(sf[0]), float(sf[1])
    sp.any = int(p["any"])
    tab = np.ascontiguousarray(p["sig53tab"], np.int16)
    sp.sig53tab = tab.ctypes.data
    keep.append(tab)
    sp._keep = keep
    return sp


def exinon_scan(p: dict, codes):
    """Exinon::intron53_c + intron53_n over a whole segment: codes[i] == *Seq::at(i).
    Returns sig5, sig3 (int16) and int53 (uint16) by column n in [0, len + 1]."""
    sp = make_scan_params(p)
    c = np.ascontiguousarray(codes, np.uint8)
    n = len(c)
    buf = np.concatenate([c, np.zeros(4, np.uint8)])       # the kernel never reads past len - 1
    s5 = np.zeros(n + 2, np.int16)
    s3 = np.zeros(n + 2, np.int16)
    i53 = np.zeros(n + 2, np.uint16)
    lib().so_exinon_scan_n(C.byref(sp), buf.ctypes.data, n, s5.ctypes.data, s3.ctypes.data, i53.ctypes.data)
    return {"sig5": s5, "sig3": s3, "int53": i53}


class SoScanParamsP(C.Structure):
    _fields_ = [("base", SoScanParams), ("patI", SoPatMat), ("patT", SoPatMat), ("codepot", C.c_void_p),
                ("ndata", C.c_int32), ("cp_order", C.c_int32), ("fact", C.c_float), ("z", C.c_float),
                ("bti", C.c_float), ("o", C.c_float)]


def exinon_scan_p(p: dict, tron):
    """Exinon::intron53_p over a whole TRON segment (tron[i] == *Seq::at(i)).  p: as for exinon_scan
    plus patI_* / patT_*, codepot (flat [ndata][3] floats), scan_fp = (fact, z, bti, o).
    Returns the SGPT6 table as (len + 2, 8) int16 and int53."""
    sp = SoScanParamsP()
    base = make_scan_params(p)
    sp.base = base
    keep = [base]
    for name, pm in (("patI", sp.patI), ("patT", sp.patT)):
        if p.get(name + "_mtx") is None:
            continue
        meta = [int(x) for x in p[name + "_meta"]]
        f = np.asarray(p[name + "_f"], np.float32)
        mtx = np.ascontiguousarray(p[name + "_mtx"], np.float32)
        pm.rows, pm.cols, pm.offset, pm.nalpha, pm.morder = meta
        pm.tonic, pm.min_elem = float(f[0]), float(f[1])
        pm.mtx = mtx.ctypes.data
        keep.append(mtx)
    if p.get("codepot") is not None:
        cp = np.ascontiguousarray(p["codepot"], np.float32)
        sp.codepot = cp.ctypes.data
        sp.ndata = cp.size // 3
        sp.cp_order = int(round(np.log(sp.ndata) / np.log(4))) - 1
        keep.append(cp)
    fp = np.asarray(p["scan_fp"], np.float32)
    sp.fact, sp.z, sp.bti, sp.o = [float(x) for x in fp]
    c = np.ascontiguousarray(tron, np.uint8)
    n = len(c)
    buf = np.concatenate([c, np.zeros(8, np.uint8)])
    out = np.zeros((n + 2, 8), np.int16)
    i53 = np.zeros(n + 2, np.uint16)
    lib().so_exinon_scan_p.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib().so_exinon_scan_p.restype = None
    lib().so_exinon_scan_p(C.byref(sp), buf.ctypes.data, n, out.ctypes.data, i53.ctypes.data)
    return {"sgpt6": out, "int53": i53}


def nuc2tron(gencode, codes_with_ends):
    """codes_with_ends: at(-1 .. len) (len + 2 bytes); returns the tron codes of at(0 .. len - 1)"""
    c = np.ascontiguousarray(codes_with_ends, np.uint8)
    g = np.ascontiguousarray(gencode, np.uint8)
    n = len(c) - 2
    out = np.zeros(max(n, 0), np.uint8)
    lib().so_nuc2tron.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib().so_nuc2tron.restype = None
    if n > 0:
        lib().so_nuc2tron(g.ctypes.data, c.ctypes.data + 1, n, out.ctypes.data)
    return out


class SoLspOpts(C.Structure):
    _fields_ = [("max_vmf_space", C.c_int32), ("sh", C.c_int32), ("ubh", C.c_int32),
                ("alg", C.c_int32)]


def lsp(p: dict, t: dict, cap: int = 1 << 16, max_vmf_space=None):
    sp, st = make_params(p), make_task(t)
    o = SoLspOpts(int(max_vmf_space if max_vmf_space is not None else p["MaxVmfSpace"]),
                  int(p["sh"]), int(p["ubh"]), int(p["alg"]))
    score = C.c_int32(0)
    unsup = C.c_int(0)
    skl = np.zeros((cap, 2), np.int32)
    n = lib().so_lsp(C.byref(sp), C.byref(st), C.byref(o), C.byref(score), skl.ctypes.data, cap,
                     C.byref(unsup))
    return {"score": score.value, "skl": skl[:min(n, cap)].copy(), "unsupported": bool(unsup.value)}


# ---------------------------------------------------------------------------
# protein x genome
# ---------------------------------------------------------------------------
class SoParamsH(C.Structure):
    _fields_ = [("gop", C.c_int32), ("gep", C.c_int32), ("lgep", C.c_int32), ("codonk1", C.c_int32),
                ("gw1", C.c_int32), ("gw2", C.c_int32), ("gw3", C.c_int32),
                ("ipen", C.c_int32), ("llmt", C.c_int32), ("nquant", C.c_int32),
                ("quant_len", C.c_int32 * MAXQ), ("quant_pen", C.c_int32 * MAXQ),
                ("avmch", C.c_int32), ("local", C.c_int32), ("lcl", C.c_int32), ("spj", C.c_int32),
                ("simdim", C.c_int32), ("simmtx", C.c_void_p),
                ("lgop", C.c_int32), ("gape1", C.c_int32), ("gape2", C.c_int32), ("ng", C.c_void_p)]


class SoTaskH(C.Structure):
    _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("sgpt6", C.c_void_p), ("b_len", C.c_int32),
                ("a_left", C.c_int32), ("a_right", C.c_int32), ("b_left", C.c_int32), ("b_right", C.c_int32),
                ("a_exgl", C.c_int32), ("a_exgr", C.c_int32), ("b_exgl", C.c_int32), ("b_exgr", C.c_int32),
                ("lw", C.c_int32), ("up", C.c_int32), ("a_len", C.c_int32), ("cip", C.c_void_p)]


def _params_h(p: dict):
    sp = SoParamsH()
    sp.gop, sp.gep, sp.lgep, sp.codonk1 = p["BasicGOP"], p["BasicGEP"], p["LongGEP"], p["codonk1"]
    sp.gw1, sp.gw2, sp.gw3 = p["GapW1"], p["GapW2"], p["GapW3"]
    sp.ipen, sp.llmt, sp.nquant = p["GapWI"], p["llmt"], p["nquant"]
    for j in range(min(p["nquant"], len(p["quant_len"]))):      # (-A0 fixtures carry no quantile table)
        sp.quant_len[j] = int(p["quant_len"][j])
        sp.quant_pen[j] = int(p["quant_pen"][j])
    sp.avmch = p["avmch"]
    sp.local = 1 if (p["lcl"] & 16) else 0
    sp.lcl = p["lcl"]
    sp.spj = p.get("spj", 1)
    sp.simdim = p["simdim"]
    sim = np.ascontiguousarray(p["simmtx"], np.int32)
    sp.simmtx = sim.ctypes.data
    sp.lgop, sp.gape1, sp.gape2 = int(p["LongGOP"]), int(p["GapE1"]), int(p["GapE2"])
    sp._keep = sim
    return sp


def _task_h(t: dict):
    st = SoTaskH()
    a = np.ascontiguousarray(t["a"], np.uint8)
    b = np.ascontiguousarray(t["b"], np.uint8)
    g = np.ascontiguousarray(t["sgpt6"], np.int16)
    st.a, st.b, st.sgpt6 = a.ctypes.data + 1, b.ctypes.data + 1, g.ctypes.data
    st.b_len = int(t["blen"])
    st.a_len = int(t.get("alen", len(a) - 2))
    for k in ("a_left", "a_right", "b_left", "b_right", "a_exgl", "a_exgr", "b_exgl", "b_exgr", "lw", "up"):
        setattr(st, k, int(t[k]))
    st._keep = (a, b, g)
    if t.get("cip") is not None:
        cip = np.ascontiguousarray(t["cip"], np.int32)
        st.cip = cip.ctypes.data
        st._keep = (a, b, g, cip)
    return st


def forward_h1_wip(p: dict, t: dict, want_trace=True, cap: int = 1 << 16):
    sp, st = _params_h(p), _task_h(t)
    score = C.c_int32(0)
    skl = np.zeros((cap, 2), np.int32)
    n = lib().so_forward_h1_wip(C.byref(sp), C.byref(st), int(want_trace), C.byref(score),
                                skl.ctypes.data, cap)
    if n < 0:
        raise RuntimeError(f"so_forward_h1_wip failed: {n}")
    return {"score": score.value, "skl": skl[:n].copy()}


def hirschberg_h1_wip(p: dict, t: dict, n_im: int):
    sp, st = _params_h(p), _task_h(t)
    score = C.c_int32(0)
    cpos = np.zeros((n_im + 1, 10), np.int32)
    ranges = np.zeros(4, np.int32)
    rc = lib().so_hirschberg_h1_wip(C.byref(sp), C.byref(st), n_im, C.byref(score),
                                    cpos.ctypes.data, ranges.ctypes.data)
    if rc < 0:
        raise RuntimeError(f"so_hirschberg_h1_wip failed: {rc}")
    return {"score": score.value, "cpos": cpos, "ranges": ranges.tolist()}


def lsp_h(p: dict, t: dict, cap: int = 1 << 16, max_vmf_space=None):
    """Aln2h1::lspH_ng restatement: score + corner list in Mfile order"""
    sp, st = _params_h(p), _task_h(t)
    keep = None
    if all(p.get(k) is not None for k in ("penalty", "sig53tab", "spj_tabs")) and t.get("int53") is not None:
        keep = _ng_h(p, t)          # blocks with fewer than 8 rows: scalar kernel
        sp.ng = C.addressof(keep[0])
    o = SoLspOpts(int(max_vmf_space if max_vmf_space is not None else p["MaxVmfSpace"]),
                  int(p["sh"]), int(p["ubh"]), int(p["alg"]))
    score = C.c_int32(0)
    unsup = C.c_int(0)
    skl = np.zeros((cap, 2), np.int32)
    n = lib().so_lsp_h(C.byref(sp), C.byref(st), C.byref(o), C.byref(score), skl.ctypes.data, cap,
                       C.byref(unsup))
    return {"score": score.value, "skl": skl[:min(n, cap)].copy(), "unsupported": bool(unsup.value)}


class SoNgH(C.Structure):
    _fields_ = [("penalty", C.c_void_p), ("n_penalty", C.c_int32), ("sig53tab", C.c_void_p),
                ("int53", C.c_void_p), ("spj_tabs", C.c_void_p), ("minl", C.c_int32),
                ("extragop", C.c_int32), ("gw3l", C.c_int32), ("noll", C.c_int32)]


def _ng_h(p: dict, t: dict):
    x = SoNgH()
    pen = np.ascontiguousarray(p["penalty"], np.int16)
    tab = np.ascontiguousarray(p["sig53tab"], np.int16)
    i53 = np.ascontiguousarray(t["int53"], np.uint16)
    spj = np.ascontiguousarray(p["spj_tabs"], np.uint8)
    x.penalty, x.n_penalty, x.sig53tab = pen.ctypes.data, len(pen), tab.ctypes.data
    x.int53, x.spj_tabs = i53.ctypes.data, spj.ctypes.data
    x.minl, x.extragop, x.gw3l, x.noll = int(p["minl"]), int(p["ExtraGOP"]), int(p["GapW3L"]), int(p["Noll"])
    return x, pen, tab, i53, spj


def trcbk_h_ng(p: dict, t: dict, cap: int = 1 << 16):
    """scalar protein kernel (Aln2h1::trcbkalignH_ng's scalar branch): score + corners.  p carries
    penalty, sig53tab, spj_tabs, minl, ExtraGOP, GapW3L; t carries int53"""
    sp, st = _params_h(p), _task_h(t)
    keep = _ng_h(p, t)
    x = keep[0]
    score = C.c_int32(0)
    skl = np.zeros((cap, 2), np.int32)
    lib().so_trcbk_h_ng.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    n = lib().so_trcbk_h_ng(C.byref(sp), C.byref(x), C.byref(st), C.byref(score), skl.ctypes.data, cap)
    if n < 0:
        raise RuntimeError(f"so_trcbk_h_ng failed: {n}")
    return {"score": score.value, "skl": skl[:n].copy()}


def hirschberg_h_ng(p: dict, t: dict, n_im: int, intvl: int):
    """Aln2h1::hirschbergH_ng (scalar protein Hirschberg pass of `-A0`)"""
    sp, st = _params_h(p), _task_h(t)
    keep = _ng_h(p, t)
    x = keep[0]
    score = C.c_int32(0)
    cpos = np.zeros((n_im + 1, 10), np.int32)
    ranges = np.zeros(4, np.int32)
    L = lib()
    L.so_hirschberg_h_ng.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
    rc = L.so_hirschberg_h_ng(C.byref(sp), C.byref(x), C.byref(st), n_im, intvl, C.byref(score),
                              cpos.ctypes.data, ranges.ctypes.data)
    if rc < 0:
        raise RuntimeError(f"so_hirschberg_h_ng failed ({rc})")
    return {"score": score.value, "cpos": cpos, "ranges": ranges.tolist()}
