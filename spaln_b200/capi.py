"""ctypes binding of the C-ABI in include/gspaln.h (libgspaln.so).

This is plumbing only: every DP cell is computed by the sm_100a kernels in
spaln_b200/csrc.  There is no CPU fallback -- if the shared library is missing
or no CUDA device is usable, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
# GSPALN_LIB: alternative build of the same library (kernel experiments)
LIB_PATH = Path(os.environ.get("GSPALN_LIB", PKG / "libgspaln.so"))

MAXQUANT = 8
MAXDIM = 32

FORWARD_WIP = 0
SCOREONLY_WIP = 1
HIRSCHBERG_WIP = 2
FORWARD_NG = 3          # scalar exact-ILD kernel (Aln2s1::forwardS_ng + Vmf trace-back)
SCOREALONE_NG = 4       # scalar score-only kernel (Aln2s1::scorealoneS_ng)
HIRSCHBERG_NG = 5       # scalar Hirschberg pass (Aln2s1::hirschbergS_ng)
END_OF_ULK = 2 ** 31 - 1 - 2

EXPORTS = [
    "gspaln_create", "gspaln_destroy", "gspaln_submit", "gspaln_upload", "gspaln_run",
    "gspaln_download", "gspaln_get_timing", "gspaln_last_error", "gspaln_device_count",
    "gspaln_version", "gspaln_task_cells", "gspaln_lsp", "gspaln_set_ng_tables",
    "gspaln_h_create", "gspaln_h_destroy", "gspaln_h_submit", "gspaln_h_upload", "gspaln_h_run",
    "gspaln_h_download", "gspaln_h_get_timing", "gspaln_h_last_error", "gspaln_h_task_cells",
    "gspaln_h_lsp", "gspaln_h_set_ng_tables",
    "gspaln_queue_create", "gspaln_queue_submit", "gspaln_queue_stats", "gspaln_queue_destroy",
    "gspaln_queue_submit_lsp", "gspaln_h_queue_create", "gspaln_h_queue_submit",
    "gspaln_h_queue_submit_lsp", "gspaln_h_queue_stats", "gspaln_h_queue_destroy",
    "gspaln_scan_create", "gspaln_scan_destroy", "gspaln_exinon_scan", "gspaln_scan_upload",
    "gspaln_scan_run", "gspaln_scan_download", "gspaln_scan_get_timing", "gspaln_scan_last_error",
    "gspaln_nuc2tron", "gspaln_scan_create_p", "gspaln_exinon_scan_p",
]


class GspalnParams(C.Structure):
    _fields_ = [
        ("gop", C.c_int32), ("gep", C.c_int32), ("lgop", C.c_int32), ("lgep", C.c_int32),
        ("noll", C.c_int32), ("ipen", C.c_int32), ("llmt", C.c_int32), ("nquant", C.c_int32),
        ("quant_len", C.c_int32 * MAXQUANT), ("quant_pen", C.c_int32 * MAXQUANT),
        ("avmch", C.c_int32), ("local", C.c_int32), ("spj", C.c_int32),
        ("simdim", C.c_int32), ("gappen1", C.c_int32),
        ("simmtx", C.c_int32 * (MAXDIM * MAXDIM)),
    ]


class GspalnTask(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("a", C.c_void_p), ("b", C.c_void_p), ("sig5", C.c_void_p), ("sig3", C.c_void_p),
        ("a_left", C.c_int32), ("a_right", C.c_int32), ("b_left", C.c_int32), ("b_right", C.c_int32),
        ("a_exgl", C.c_int32), ("a_exgr", C.c_int32), ("b_exgl", C.c_int32), ("b_exgr", C.c_int32),
        ("lw", C.c_int32), ("up", C.c_int32), ("skl_cap", C.c_int32), ("n_imd", C.c_int32),
        ("int53", C.c_void_p), ("cip", C.c_void_p),
    ]


class GspalnResult(C.Structure):
    _fields_ = [
        ("score", C.c_int32), ("status", C.c_int32), ("n_skl", C.c_int32), ("reserved", C.c_int32),
        ("cells", C.c_int64), ("skl", C.c_void_p),
        ("ranges", C.c_int32 * 4), ("cpos", C.c_void_p),
    ]


class GspalnLspOpts(C.Structure):
    _fields_ = [("max_vmf_space", C.c_int32), ("sh", C.c_int32), ("ubh", C.c_int32), ("alg", C.c_int32)]


class GspalnTiming(C.Structure):
    _fields_ = [
        ("h2d_ms", C.c_float), ("kernel_ms", C.c_float), ("d2h_ms", C.c_float),
        ("launches", C.c_int32), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("trace_bytes", C.c_int64), ("cells", C.c_int64),
    ]


class GspalnPatMat(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("offset", C.c_int32), ("nalpha", C.c_int32),
                ("morder", C.c_int32), ("tonic", C.c_float), ("min_elem", C.c_float), ("mtx", C.c_void_p)]


class GspalnScanParams(C.Structure):
    _fields_ = [("pat5", GspalnPatMat), ("pat3", GspalnPatMat), ("fS", C.c_float), ("sss", C.c_float),
                ("any", C.c_int32), ("sig53tab", C.c_int16 * 32)]


class GspalnScanParamsP(C.Structure):
    _fields_ = [("base", GspalnScanParams), ("patI", GspalnPatMat), ("patT", GspalnPatMat),
                ("codepot", C.c_void_p), ("ndata", C.c_int32), ("cp_order", C.c_int32),
                ("fact", C.c_float), ("z", C.c_float), ("bti", C.c_float), ("o", C.c_float)]


class GspalnHParams(C.Structure):
    _fields_ = [
        ("gop", C.c_int32), ("gep", C.c_int32), ("lgep", C.c_int32), ("codonk1", C.c_int32),
        ("gw1", C.c_int32), ("gw2", C.c_int32), ("gw3", C.c_int32),
        ("ipen", C.c_int32), ("llmt", C.c_int32), ("nquant", C.c_int32),
        ("quant_len", C.c_int32 * MAXQUANT), ("quant_pen", C.c_int32 * MAXQUANT),
        ("avmch", C.c_int32), ("lcl", C.c_int32), ("spj", C.c_int32), ("simdim", C.c_int32),
        ("simmtx", C.c_int32 * (MAXDIM * MAXDIM)),
        ("lgop", C.c_int32), ("gape1", C.c_int32), ("gape2", C.c_int32),
    ]


class GspalnHTask(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("a", C.c_void_p), ("b", C.c_void_p), ("sg", C.c_void_p),
        ("b_len", C.c_int32),
        ("a_left", C.c_int32), ("a_right", C.c_int32), ("b_left", C.c_int32), ("b_right", C.c_int32),
        ("a_exgl", C.c_int32), ("a_exgr", C.c_int32), ("b_exgl", C.c_int32), ("b_exgr", C.c_int32),
        ("lw", C.c_int32), ("up", C.c_int32), ("skl_cap", C.c_int32), ("n_imd", C.c_int32),
        ("a_len", C.c_int32), ("int53", C.c_void_p), ("cip", C.c_void_p),
    ]


# SGPT6 (src/codepot.h:34-43): 6 shorts + 2 chars
SGPT6_DTYPE = np.dtype([("sig5", "<i2"), ("sig3", "<i2"), ("sigS", "<i2"), ("sigT", "<i2"),
                        ("sigE", "<i2"), ("sigI", "<i2"), ("phs5", "i1"), ("phs3", "i1")])
assert SGPT6_DTYPE.itemsize == 14


def sgpt6_from_table(tab) -> np.ndarray:
    """(n, 8) int16 table (sig5, sig3, sigS, sigT, sigE, sigI, phs5, phs3) -> SGPT6 records"""
    tab = np.asarray(tab)
    out = np.zeros(len(tab), SGPT6_DTYPE)
    for i, k in enumerate(("sig5", "sig3", "sigS", "sigT", "sigE", "sigI", "phs5", "phs3")):
        out[k] = tab[:, i]
    return out


_lib = None


def load():
    """Load libgspaln.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the DP engine)")
    lib = C.CDLL(str(LIB_PATH))
    lib.gspaln_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(GspalnParams), C.c_int]
    lib.gspaln_create.restype = C.c_int
    lib.gspaln_destroy.argtypes = [C.c_void_p]
    lib.gspaln_destroy.restype = None
    lib.gspaln_set_ng_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
    lib.gspaln_submit.argtypes = [C.c_void_p, C.POINTER(GspalnTask), C.c_int, C.POINTER(GspalnResult)]
    lib.gspaln_upload.argtypes = [C.c_void_p, C.POINTER(GspalnTask), C.c_int]
    lib.gspaln_run.argtypes = [C.c_void_p]
    lib.gspaln_download.argtypes = [C.c_void_p, C.POINTER(GspalnResult)]
    lib.gspaln_lsp.argtypes = [C.c_void_p, C.POINTER(GspalnTask), C.c_int, C.POINTER(GspalnLspOpts),
                               C.POINTER(GspalnResult)]
    lib.gspaln_get_timing.argtypes = [C.c_void_p, C.POINTER(GspalnTiming)]
    lib.gspaln_last_error.argtypes = [C.c_void_p]
    lib.gspaln_last_error.restype = C.c_char_p
    lib.gspaln_device_count.restype = C.c_int
    lib.gspaln_version.restype = C.c_char_p
    lib.gspaln_task_cells.argtypes = [C.POINTER(GspalnTask)]
    lib.gspaln_task_cells.restype = C.c_int64
    lib.gspaln_queue_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int]
    lib.gspaln_queue_submit.argtypes = [C.c_void_p, C.POINTER(GspalnTask), C.POINTER(GspalnResult)]
    lib.gspaln_queue_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.gspaln_queue_destroy.argtypes = [C.c_void_p]
    lib.gspaln_queue_destroy.restype = None
    lib.gspaln_scan_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(GspalnScanParams), C.c_int]
    lib.gspaln_scan_destroy.argtypes = [C.c_void_p]
    lib.gspaln_scan_destroy.restype = None
    lib.gspaln_exinon_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gspaln_scan_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    lib.gspaln_scan_run.argtypes = [C.c_void_p]
    lib.gspaln_scan_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gspaln_scan_get_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                           C.POINTER(C.c_float)]
    lib.gspaln_scan_last_error.argtypes = [C.c_void_p]
    lib.gspaln_scan_last_error.restype = C.c_char_p
    lib.gspaln_scan_create_p.argtypes = [C.POINTER(C.c_void_p), C.POINTER(GspalnScanParamsP), C.c_int]
    lib.gspaln_exinon_scan_p.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.gspaln_nuc2tron.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_float)]
    lib.gspaln_h_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(GspalnHParams), C.c_int]
    lib.gspaln_h_destroy.argtypes = [C.c_void_p]
    lib.gspaln_h_destroy.restype = None
    lib.gspaln_h_set_ng_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                           C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    lib.gspaln_h_submit.argtypes = [C.c_void_p, C.POINTER(GspalnHTask), C.c_int, C.POINTER(GspalnResult)]
    lib.gspaln_h_upload.argtypes = [C.c_void_p, C.POINTER(GspalnHTask), C.c_int]
    lib.gspaln_h_run.argtypes = [C.c_void_p]
    lib.gspaln_h_download.argtypes = [C.c_void_p, C.POINTER(GspalnResult)]
    lib.gspaln_h_get_timing.argtypes = [C.c_void_p, C.POINTER(GspalnTiming)]
    lib.gspaln_h_last_error.argtypes = [C.c_void_p]
    lib.gspaln_h_last_error.restype = C.c_char_p
    lib.gspaln_h_lsp.argtypes = [C.c_void_p, C.POINTER(GspalnHTask), C.c_int, C.POINTER(GspalnLspOpts),
                                 C.POINTER(GspalnResult)]
    lib.gspaln_h_task_cells.argtypes = [C.POINTER(GspalnHTask)]
    lib.gspaln_h_task_cells.restype = C.c_int64
    _lib = lib
    return lib


def make_params(p: dict) -> GspalnParams:
    """p uses the reference's own names (PwdB / IntronPrm / algmode fields)."""
    gp = GspalnParams()
    gp.gop, gp.gep = int(p["BasicGOP"]), int(p["BasicGEP"])
    gp.lgop, gp.lgep = int(p["LongGOP"]), int(p["LongGEP"])
    gp.noll = int(p["Noll"])
    gp.ipen = int(p["GapWI"])
    gp.llmt = int(p["llmt"])
    # under -A0 / -A1 the reference does not build the quantised intron penalty of the `_wip` kernels
    # (IntronPenalty::qm); such parameter sets drive the exact-ILD kernels only and get one flat bin
    gp.nquant = min(int(p["nquant"]), len(p["quant_len"]))
    for j in range(gp.nquant):
        gp.quant_len[j] = int(p["quant_len"][j])
        gp.quant_pen[j] = int(p["quant_pen"][j])
    if gp.nquant == 0:
        gp.nquant, gp.quant_len[0], gp.quant_pen[0] = 1, 0, 0
    gp.avmch = int(p["avmch"])
    gp.local = 1 if (int(p["lcl"]) & 16) else 0
    gp.spj = int(p.get("spj", 1))
    d = int(p["simdim"])
    gp.simdim = d
    gp.gappen1 = int(p["GapPenalty1"])
    sim = np.asarray(p["simmtx"], np.int32).reshape(d, d)
    flat = sim.ravel()
    for i in range(d * d):
        gp.simmtx[i] = int(flat[i])
    return gp


def make_scan_params(p: dict):
    """splice-signal scan parameters: PSSMs pat5_* / pat3_* (meta = rows, cols, offset, nalpha,
    morder; f = tonic, min_elem; mtx), scan_f = (Exinon::fS, alprm2.sss), any, sig53tab.
    Returns (struct, arrays to keep alive)."""
    sp = GspalnScanParams()
    keep = []
    for name, pm in (("pat5", sp.pat5), ("pat3", sp.pat3)):
        if p.get(name + "_mtx") is None:
            continue
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
    tab = np.asarray(p["sig53tab"], np.int16)
    for i in range(32):
        sp.sig53tab[i] = int(tab[i])
    return sp, keep


def make_scan_params_p(p: dict):
    """protein-side scan parameters: make_scan_params keys + patI_* / patT_*, codepot (flat
    [ndata][3]), scan_fp = (Exinon::fact, alprm2.z, alprm2.bti, alprm2.o)"""
    sp = GspalnScanParamsP()
    base, keep = make_scan_params(p)
    sp.base = base
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
    return sp, keep


def make_h_params(p: dict) -> GspalnHParams:
    """protein x genome parameter set; p uses the reference's names (PwdB / IntronPrm / algmode)"""
    gp = GspalnHParams()
    gp.gop, gp.gep = int(p["BasicGOP"]), int(p["BasicGEP"])
    gp.lgep, gp.codonk1 = int(p["LongGEP"]), int(p["codonk1"])
    gp.gw1, gp.gw2, gp.gw3 = int(p["GapW1"]), int(p["GapW2"]), int(p["GapW3"])
    gp.ipen, gp.llmt = int(p["GapWI"]), int(p["llmt"])
    gp.nquant = min(int(p["nquant"]), len(p["quant_len"]))      # (-A0 / -A1 parameter sets: see the DNA twin)
    for j in range(gp.nquant):
        gp.quant_len[j] = int(p["quant_len"][j])
        gp.quant_pen[j] = int(p["quant_pen"][j])
    if gp.nquant == 0:
        gp.nquant, gp.quant_len[0], gp.quant_pen[0] = 1, 0, 0
    gp.avmch = int(p["avmch"])
    gp.lcl = int(p["lcl"])
    gp.spj = int(p.get("spj", 1))
    d = int(p["simdim"])
    gp.simdim = d
    flat = np.asarray(p["simmtx"], np.int32).reshape(d, d).ravel()
    for i in range(d * d):
        gp.simmtx[i] = int(flat[i])
    gp.lgop = int(p.get("LongGOP", 0))
    gp.gape1, gp.gape2 = int(p.get("GapE1", 0)), int(p.get("GapE2", 0))
    return gp
