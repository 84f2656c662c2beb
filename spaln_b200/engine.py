"""Host-side mirror of the reference's DP call surface for the spliced-DP path.

Reference surface (C++): `SimdAln2s1(seqs, pwd, wdw, spjcs, cip, mode)` followed
by `forwardS1_wip(mfd)` / `scoreonlyS1_wip()` (src/fwd2s1_simd.h:184-196), one
problem per call.  Here a `Problem` carries exactly the inputs that constructor
reads (query / genome codes and ranges, end-gap flags, Exinon splice-signal
table, band window) and `Engine.forwardS1_wip(problems)` /
`Engine.scoreonlyS1_wip(problems)` run a whole batch on the GPU through the
C-ABI (include/gspaln.h).  Results are the reference's: the VTYPE score and the
(m, n) corner list appended to the caller's Mfile.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import capi


@dataclass
class Problem:
    a: np.ndarray           # uint8 query codes, a[i] == *Seq::at(i)
    b: np.ndarray           # uint8 genome codes
    sig5: np.ndarray        # int16, Exinon::data_n[n].sig5 by column n (len >= b_right + 1)
    sig3: np.ndarray        # int16
    a_left: int
    a_right: int
    b_left: int
    b_right: int
    lw: int
    up: int
    a_exgl: int = 1
    a_exgr: int = 1
    b_exgl: int = 1
    b_exgr: int = 1
    skl_cap: int = 0        # 0: a_len + b_len + 8
    n_imd: int = 0          # hirschbergS1_wip only: number of intermediate rows
    int53: np.ndarray = None    # uint16 Exinon::int53[n] nibbles by column (scalar kernel only)
    cip: np.ndarray = None      # int32 Cip_score::cip_score(m) by query position 0 .. a_right (optional)

    @staticmethod
    def from_export(ex: dict, lw: int, up: int) -> "Problem":
        """ex: tests/ref_harness.py::RefTask.export() layout (arrays start at at(-1))."""
        return Problem(int53=(np.ascontiguousarray(ex["int53"], np.uint16)
                              if ex.get("int53") is not None else None),
                       cip=(np.ascontiguousarray(ex["cip"], np.int32) if ex.get("cip") is not None else None),
                       a=np.ascontiguousarray(ex["a"][1:], np.uint8),
                       b=np.ascontiguousarray(ex["b"][1:], np.uint8),
                       sig5=np.ascontiguousarray(ex["sig5"], np.int16),
                       sig3=np.ascontiguousarray(ex["sig3"], np.int16),
                       a_left=ex["a_left"], a_right=ex["a_right"],
                       b_left=ex["b_left"], b_right=ex["b_right"], lw=lw, up=up,
                       a_exgl=ex["a_exgl"], a_exgr=ex["a_exgr"],
                       b_exgl=ex["b_exgl"], b_exgr=ex["b_exgr"])


@dataclass
class Result:
    score: int
    status: int
    skl: np.ndarray         # (n, 2) int32 corners, alignment end first
    cells: int
    ranges: tuple = ()      # hirschbergS1_wip: (a_left, a_right, b_left, b_right) afterwards
    cpos: np.ndarray = None  # hirschbergS1_wip: (n_imd + 1, 10) crossing records


@dataclass
class Timing:
    h2d_ms: float = 0.0
    kernel_ms: float = 0.0
    d2h_ms: float = 0.0
    launches: int = 0
    h2d_bytes: int = 0
    d2h_bytes: int = 0
    trace_bytes: int = 0
    cells: int = 0


class EngineError(RuntimeError):
    pass


_RESULT_DTYPE = np.dtype([("score", "<i4"), ("status", "<i4"), ("n_skl", "<i4"), ("reserved", "<i4"),
                          ("cells", "<i8"), ("skl", "<u8"), ("ranges", "<i4", (4,)), ("cpos", "<u8")])
assert _RESULT_DTYPE.itemsize == C.sizeof(capi.GspalnResult)


class PackedBatch:
    """Task descriptors marshalled once (the sequence / signal buffers stay host numpy
    arrays and are packed + copied to the device by every submit) plus a flat result area."""

    def __init__(self, arr, keep, n, reuse=None):
        """reuse: a PackedBatch of an earlier call whose result area (corner pool, result records) is
        taken over if it is large enough -- what a caller that submits batch after batch does instead
        of faulting in tens of MB of fresh pages per batch"""
        self.arr, self.keep, self.n = arr, keep, n
        if isinstance(arr, np.ndarray):     # descriptors built with array operations (pack_global)
            caps = arr["skl_cap"][:n].astype(np.int64)
            self.arr = arr.ctypes.data_as(C.POINTER(capi.GspalnTask))
            self.keep = (arr, keep)
        else:
            caps = np.array([arr[i].skl_cap for i in range(n)], np.int64)
        self.off = np.concatenate([[0], np.cumsum(caps)])
        need = max(int(self.off[-1]), 1)
        if reuse is not None and len(reuse.skl) >= need and len(reuse.res) >= max(n, 1):
            self.skl, self.res = reuse.skl, reuse.res
            self.res[:] = 0
        else:
            self.skl = np.zeros((need, 2), np.int32)
            self.res = np.zeros(max(n, 1), _RESULT_DTYPE)
        self.res["skl"][:n] = self.skl.ctypes.data + 8 * self.off[:-1].astype(np.uint64)
        self.res["skl"][:n][caps == 0] = 0

    @property
    def scores(self):
        return self.res["score"][:self.n]

    @property
    def status(self):
        return self.res["status"][:self.n]

    def corners(self, i):
        k = min(int(self.res["n_skl"][i]), int(self.off[i + 1] - self.off[i]))
        return self.skl[self.off[i]:self.off[i] + k]


class Engine:
    """One engine per (parameter set, device); not thread-safe."""

    def __init__(self, params: dict, device: int = 0):
        self.lib = capi.load()
        raw = isinstance(params, capi.GspalnParams)     # a frozen gspaln_params as the C-ABI takes it
        self._gp = params if raw else capi.make_params(params)
        if raw:
            params = {}
        self._h = C.c_void_p()
        rc = self.lib.gspaln_create(C.byref(self._h), C.byref(self._gp), device)
        if rc != 0:
            raise EngineError(f"gspaln_create failed ({rc}): no usable CUDA device or bad "
                              "parameters; the DP engine has no CPU fallback")
        self.device = device
        self._tasks = None
        self._keep = None
        self._n = 0
        # tables of the exact intron scoring (scalar kernel), when the parameter set carries them
        if params.get("penalty") is not None and params.get("sig53tab") is not None:
            self.set_ng_tables(params["sig53tab"], params["penalty"], int(params.get("codonk1", 2 ** 31 - 1)))

    def set_ng_tables(self, sig53tab, penalty, codonk1):
        """Exinon::sig53tab (544 shorts), IntronPenalty::Penalty(0 .. n - 1), PwdB::codonk1"""
        tab = np.ascontiguousarray(sig53tab, np.int16)
        pen = np.ascontiguousarray(penalty, np.int16)
        if tab.size != 544:
            raise ValueError("sig53tab must hold 544 entries")
        self._check(self.lib.gspaln_set_ng_tables(self._h, tab.ctypes.data, pen.ctypes.data, pen.size,
                                                  int(codonk1)), "gspaln_set_ng_tables")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.gspaln_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def _pack(self, problems, kind):
        n = len(problems)
        arr = (capi.GspalnTask * max(n, 1))()
        keep = []
        for i, p in enumerate(problems):
            a = np.ascontiguousarray(p.a, np.uint8)
            b = np.ascontiguousarray(p.b, np.uint8)
            s5 = np.ascontiguousarray(p.sig5, np.int16)
            s3 = np.ascontiguousarray(p.sig3, np.int16)
            if len(a) < p.a_right or len(b) < p.b_right or len(s5) <= p.b_right or len(s3) <= p.b_right:
                raise ValueError("problem arrays shorter than the stated ranges")
            i53 = None
            if p.int53 is not None:
                i53 = np.ascontiguousarray(p.int53, np.uint16)
                if len(i53) <= p.b_right:
                    raise ValueError("int53 shorter than the stated range")
            cip = None
            if p.cip is not None:
                cip = np.ascontiguousarray(p.cip, np.int32)
                if len(cip) <= p.a_right:
                    raise ValueError("cip shorter than the stated range")
            keep.append((a, b, s5, s3, i53, cip))
            t = arr[i]
            t.kind = kind
            t.a, t.b = a.ctypes.data, b.ctypes.data
            t.sig5, t.sig3 = s5.ctypes.data, s3.ctypes.data
            t.int53 = i53.ctypes.data if i53 is not None else None
            t.cip = cip.ctypes.data if cip is not None else None
            t.a_left, t.a_right, t.b_left, t.b_right = p.a_left, p.a_right, p.b_left, p.b_right
            t.a_exgl, t.a_exgr, t.b_exgl, t.b_exgr = p.a_exgl, p.a_exgr, p.b_exgl, p.b_exgr
            t.lw, t.up = p.lw, p.up
            cap = p.skl_cap or ((p.a_right - p.a_left) + (p.b_right - p.b_left) + 8)
            t.skl_cap = cap if kind in (capi.FORWARD_WIP, capi.FORWARD_NG) else 0
            t.n_imd = int(p.n_imd) if kind in (capi.HIRSCHBERG_WIP, capi.HIRSCHBERG_NG) else 0
        return arr, keep

    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.gspaln_last_error(self._h)
            raise EngineError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def _results(self, n, arr):
        res = (capi.GspalnResult * max(n, 1))()
        bufs = []
        for i in range(n):
            cap = arr[i].skl_cap
            buf = np.zeros((max(cap, 1), 2), np.int32)
            cp = np.zeros((arr[i].n_imd + 1, 10), np.int32) if arr[i].kind in (capi.HIRSCHBERG_WIP, capi.HIRSCHBERG_NG) else None
            bufs.append((buf, cp))
            res[i].skl = buf.ctypes.data if cap > 0 else None
            res[i].cpos = cp.ctypes.data if cp is not None else None
        return res, bufs

    def _collect(self, n, res, bufs, arr):
        out = []
        for i in range(n):
            k = min(res[i].n_skl, arr[i].skl_cap)
            out.append(Result(int(res[i].score), int(res[i].status), bufs[i][0][:max(k, 0)].copy(),
                              int(res[i].cells), tuple(int(x) for x in res[i].ranges), bufs[i][1]))
        return out

    # ---- one-shot (host buffers in, host results out) -------------------
    def submit(self, problems, kind=capi.FORWARD_WIP):
        arr, keep = self._pack(problems, kind)
        n = len(problems)
        res, bufs = self._results(n, arr)
        self._check(self.lib.gspaln_submit(self._h, arr, n, res), "gspaln_submit")
        self._n = n
        return self._collect(n, res, bufs, arr)

    def pack(self, problems, kind=capi.FORWARD_WIP) -> PackedBatch:
        arr, keep = self._pack(problems, kind)
        return PackedBatch(arr, keep, len(problems))

    def submit_packed(self, batch: PackedBatch) -> PackedBatch:
        """host buffers in, host results out (scores / status / corners in `batch`)"""
        res = batch.res.ctypes.data_as(C.POINTER(capi.GspalnResult))
        self._check(self.lib.gspaln_submit(self._h, batch.arr, batch.n, res), "gspaln_submit")
        return batch

    def pack_global(self, g: dict, index, kind=capi.FORWARD_WIP, skl_cap=512, sh=100, reuse=None) -> PackedBatch:
        """Task descriptors of the problems `index` of a flat query set (the layout of
        workload.config2_global: concatenated query / genome codes and per-column tables plus a
        length table), filled with array operations: what a caller that keeps its sequences in
        flat buffers does per job.  Same band (stripe) and flags as workload.global_problems."""
        index = np.asarray(index, np.int64)
        n = len(index)
        off = g.get("_off")
        if off is None:
            lens = g["lens"]
            off = g["_off"] = (np.concatenate([[0], np.cumsum(lens[:, 0])]), np.concatenate([[0], np.cumsum(lens[:, 1])]),
                               np.concatenate([[0], np.cumsum(lens[:, 1] + 2)]))
        la, lb = g["lens"][index, 0], g["lens"][index, 1]
        arr = np.zeros(max(n, 1), np.dtype(capi.GspalnTask))
        t = arr[:n]
        t["kind"] = kind
        t["a"] = g["a"].ctypes.data + off[0][index]
        t["b"] = g["b"].ctypes.data + off[1][index]
        t["sig5"] = g["sig5"].ctypes.data + 2 * off[2][index]
        t["sig3"] = g["sig3"].ctypes.data + 2 * off[2][index]
        t["int53"] = g["int53"].ctypes.data + 2 * off[2][index]
        t["a_right"], t["b_right"] = la, lb
        for k in ("a_exgl", "a_exgr", "b_exgl", "b_exgr"):
            t[k] = 1
        up, lw = lb - la, np.zeros(n, np.int64)
        up, lw = np.maximum(up, lw), np.minimum(up, lw)
        t["up"] = np.minimum(up + sh, lb)
        t["lw"] = np.maximum(lw - sh, -la)
        t["skl_cap"] = skl_cap if kind in (capi.FORWARD_WIP, capi.FORWARD_NG) else 0
        return PackedBatch(arr, g, n, reuse)

    def forwardS1_wip(self, problems):
        return self.submit(problems, capi.FORWARD_WIP)

    def scoreonlyS1_wip(self, problems):
        return self.submit(problems, capi.SCOREONLY_WIP)

    def forwardS_ng(self, problems):
        """Aln2s1::trcbkalignS_ng on its scalar branch (forwardS_ng + Vmf trace-back, exact intron
        scoring; src/fwd2s1.cc:217-444, 1667-1710): score + corners.  Problems carry int53."""
        return self.submit(problems, capi.FORWARD_NG)

    def scorealoneS_ng(self, problems):
        """Aln2s1::scorealoneS_ng (src/fwd2s1.cc:1163-1336): scalar score-only kernel, exact intron
        scoring; problems carry int53"""
        return self.submit(problems, capi.SCOREALONE_NG)

    def HomScoreS_ng(self, problems):
        """HomScoreS_ng's kernel choice (src/fwd2s1.cc:2696-2716) for -A2 / -A3: queries shorter
        than 4 residues go to the scalar score-only kernel, the rest to scoreonlyS1_wip.  (The band
        is the caller's: the reference computes it with stripe(seqs, &wdw, alprm.sh).)"""
        small = [i for i, p in enumerate(problems) if p.a_right - p.a_left < 4]
        big = [i for i, p in enumerate(problems) if p.a_right - p.a_left >= 4]
        out = [None] * len(problems)
        if small:
            for i, r in zip(small, self.submit([problems[i] for i in small], capi.SCOREALONE_NG)):
                out[i] = r
        if big:
            for i, r in zip(big, self.submit([problems[i] for i in big], capi.SCOREONLY_WIP)):
                out[i] = r
        return out

    def hirschbergS1_wip(self, problems):
        """problems carry n_imd; results carry score, ranges and cpos (Dim10 records)"""
        return self.submit(problems, capi.HIRSCHBERG_WIP)

    def hirschbergS_ng(self, problems):
        """Aln2s1::hirschbergS_ng, the scalar Hirschberg pass of `-A0` (src/fwd2s1.cc:764-1104).  Problems
        carry n_imd = the number of intermediate rows lspS_ng asks for before its even-division
        correction, and int53; results carry score, ranges and cpos (entries [8], [9]: the diagonal
        bounds of each block)"""
        return self.submit(problems, capi.HIRSCHBERG_NG)

    def lspS_ng(self, problems, max_vmf_space=32 * 1024 * 1024, sh=100, ubh=0, alg=2):
        """Aln2s1::lspS_ng over a batch (src/fwd2s1.cc:1801-1897): trace-back vs
        multi-intermediate Hirschberg dispatch with the reference's space estimate, block
        re-alignments included.  Returns score + corner list in Mfile order."""
        arr, keep = self._pack(problems, capi.FORWARD_WIP)
        n = len(problems)
        res, bufs = self._results(n, arr)
        o = capi.GspalnLspOpts(int(max_vmf_space), int(sh), int(ubh), int(alg))
        self._check(self.lib.gspaln_lsp(self._h, arr, n, C.byref(o), res), "gspaln_lsp")
        self._n = n
        return self._collect(n, res, bufs, arr)

    def lsp_packed(self, batch: PackedBatch, max_vmf_space=32 * 1024 * 1024, sh=100, ubh=0, alg=2) -> PackedBatch:
        """lspS_ng over task descriptors marshalled once (`pack`): host buffers in, scores / status /
        corners in `batch` -- the driver without the per-call Python marshalling"""
        res = batch.res.ctypes.data_as(C.POINTER(capi.GspalnResult))
        o = capi.GspalnLspOpts(int(max_vmf_space), int(sh), int(ubh), int(alg))
        self._check(self.lib.gspaln_lsp(self._h, batch.arr, batch.n, C.byref(o), res), "gspaln_lsp")
        return batch

    # ---- coalescing queue: the shape of the literal drop-in ---------------
    def queue(self, max_batch=256, max_wait_us=200) -> "SubmitQueue":
        """Thread-safe single-problem submits, coalesced into batches by a dispatcher thread
        (what Spaln's pthread workers would call, one problem at a time)."""
        return SubmitQueue(self, max_batch, max_wait_us)

    # ---- split form (batch resident in HBM) -----------------------------
    def upload(self, problems, kind=capi.FORWARD_WIP):
        arr, keep = self._pack(problems, kind)
        self._check(self.lib.gspaln_upload(self._h, arr, len(problems)), "gspaln_upload")
        self._tasks, self._keep, self._n = arr, keep, len(problems)

    def run(self):
        self._check(self.lib.gspaln_run(self._h), "gspaln_run")

    def download(self):
        n = self._n
        res, bufs = self._results(n, self._tasks)
        self._check(self.lib.gspaln_download(self._h, res), "gspaln_download")
        return self._collect(n, res, bufs, self._tasks)

    def timing(self) -> Timing:
        t = capi.GspalnTiming()
        self.lib.gspaln_get_timing(self._h, C.byref(t))
        return Timing(t.h2d_ms, t.kernel_ms, t.d2h_ms, t.launches, t.h2d_bytes, t.d2h_bytes,
                      t.trace_bytes, t.cells)


class SubmitQueue:
    """gspaln_queue_*: `forwardS1_wip(problem)` may be called from many threads at once"""

    def __init__(self, engine: Engine, max_batch: int, max_wait_us: int):
        self.eng = engine
        self._q = C.c_void_p()
        rc = engine.lib.gspaln_queue_create(C.byref(self._q), engine._h, max_batch, max_wait_us)
        if rc != 0:
            raise EngineError(f"gspaln_queue_create failed ({rc})")

    def submit(self, problem: Problem, kind=capi.FORWARD_WIP) -> Result:
        arr, keep = self.eng._pack([problem], kind)
        res, bufs = self.eng._results(1, arr)
        rc = self.eng.lib.gspaln_queue_submit(self._q, arr, res)      # ctypes releases the GIL
        if rc != 0:
            raise EngineError(f"gspaln_queue_submit failed ({rc})")
        return self.eng._collect(1, res, bufs, arr)[0]

    def forwardS1_wip(self, problem: Problem) -> Result:
        return self.submit(problem, capi.FORWARD_WIP)

    def stats(self):
        t, b = C.c_int64(0), C.c_int64(0)
        self.eng.lib.gspaln_queue_stats(self._q, C.byref(t), C.byref(b))
        return t.value, b.value

    def close(self):
        if self._q and self._q.value:
            self.eng.lib.gspaln_queue_destroy(self._q)
            self._q = C.c_void_p()


# ---------------------------------------------------------------------------
# protein query x genomic segment (SimdAln2h1, src/fwd2h1_simd.h:69-382)
# ---------------------------------------------------------------------------
@dataclass
class ProblemH:
    a: np.ndarray           # uint8 amino-acid codes, a[i] == *Seq::at(i)
    b: np.ndarray           # uint8 tron codes of the genomic segment
    sgpt6: np.ndarray       # SGPT6 records (capi.SGPT6_DTYPE) by column n in [0, b_len + 1]
    b_len: int
    a_left: int
    a_right: int
    b_left: int
    b_right: int
    lw: int
    up: int
    a_exgl: int = 1
    a_exgr: int = 1
    b_exgl: int = 1
    b_exgr: int = 1
    skl_cap: int = 0
    a_len: int = 0          # Seq::len of the query (0: len(a) - 1, arrays carry one pad residue)
    n_imd: int = 0          # hirschbergH1_wip only: number of intermediate rows
    int53: np.ndarray = None    # uint16 Exinon::int53[n] nibbles by column (scalar kernel only)
    cip: np.ndarray = None      # int32 Cip_score::cip_score(c) by coding position 0 .. 3 a_right + 1 (optional)

    @staticmethod
    def from_export(ex: dict, lw: int, up: int) -> "ProblemH":
        """ex: tests/ref_harness.py::RefTask.export_p() layout (arrays start at at(-1))."""
        return ProblemH(int53=(np.ascontiguousarray(ex["int53"], np.uint16)
                               if ex.get("int53") is not None else None),
                        cip=(np.ascontiguousarray(ex["cip"], np.int32) if ex.get("cip") is not None else None),
                        a=np.ascontiguousarray(ex["a"][1:], np.uint8),
                        b=np.ascontiguousarray(ex["b"][1:], np.uint8),
                        sgpt6=capi.sgpt6_from_table(ex["sgpt6"]), b_len=int(ex["blen"]),
                        a_left=ex["a_left"], a_right=ex["a_right"],
                        b_left=ex["b_left"], b_right=ex["b_right"], lw=lw, up=up,
                        a_exgl=ex["a_exgl"], a_exgr=ex["a_exgr"],
                        b_exgl=ex["b_exgl"], b_exgr=ex["b_exgr"],
                        a_len=int(ex.get("alen", len(ex["a"]) - 2)))


class EngineH:
    """Protein x genome engine: `forwardH1_wip(problems)` == one SimdAln2h1 construction +
    forwardH1_wip(mfd) per problem (src/fwd2h1_wip_simd.h:50-336), batched on the GPU."""

    def __init__(self, params: dict, device: int = 0):
        self.lib = capi.load()
        raw = isinstance(params, capi.GspalnHParams)
        self._gp = params if raw else capi.make_h_params(params)
        if raw:
            params = {}
        self._h = C.c_void_p()
        rc = self.lib.gspaln_h_create(C.byref(self._h), C.byref(self._gp), device)
        if rc != 0:
            raise EngineError(f"gspaln_h_create failed ({rc}): no usable CUDA device or bad "
                              "parameters; the DP engine has no CPU fallback")
        self.device = device
        self._tasks = None
        self._keep = None
        self._n = 0
        # tables of the scalar kernel, when the parameter set carries them
        if all(params.get(k) is not None for k in ("penalty", "sig53tab", "spj_tabs")):
            self.set_ng_tables(params["sig53tab"], params["penalty"], params["spj_tabs"], int(params["minl"]),
                               int(params["ExtraGOP"]), int(params["GapW3L"]), int(params["Noll"]))

    def set_ng_tables(self, sig53tab, penalty, spj_tabs, minl, extragop, gw3l, noll):
        """tables of the scalar kernel forwardH_ng (gspaln_h_set_ng_tables)"""
        tab = np.ascontiguousarray(sig53tab, np.int16)
        pen = np.ascontiguousarray(penalty, np.int16)
        spj = np.ascontiguousarray(spj_tabs, np.uint8)
        self._check(self.lib.gspaln_h_set_ng_tables(
            self._h, tab.ctypes.data, pen.ctypes.data, pen.size, spj.ctypes.data, int(minl),
            int(extragop), int(gw3l), int(noll)), "gspaln_h_set_ng_tables")

    def HomScoreH_ng(self, problems):
        """HomScoreH_ng's kernel choice (src/fwd2h1.cc:3293-3312) for -A2 / -A3: queries shorter than
        8 residues go to the scalar kernel forwardH_ng (its score does not depend on the path
        records), the rest to forwardH1_wip without trace-back"""
        small = [i for i, p in enumerate(problems) if p.a_right - p.a_left < 8]
        big = [i for i, p in enumerate(problems) if p.a_right - p.a_left >= 8]
        out = [None] * len(problems)
        if small:
            for i, r in zip(small, self.submit([problems[i] for i in small], capi.FORWARD_NG)):
                out[i] = r
        if big:
            for i, r in zip(big, self.submit([problems[i] for i in big], capi.SCOREONLY_WIP)):
                out[i] = r
        return out

    def hirschbergH_ng(self, problems):
        """Aln2h1::hirschbergH_ng, the scalar protein Hirschberg pass of `-A0` (src/fwd2h1.cc:1085-1520).
        Problems carry n_imd (the count before lspH_ng's even-division correction) and int53; results
        carry score, ranges and cpos (entries [8], [9]: the diagonal bounds of each block)"""
        return self.submit(problems, capi.HIRSCHBERG_NG)

    def forwardH_ng(self, problems):
        """Aln2h1::trcbkalignH_ng on its scalar branch (forwardH_ng + Vmf trace-back, exact intron
        scoring; src/fwd2h1.cc:294-617, 1997-2041): score + corners.  Problems carry int53."""
        return self.submit(problems, capi.FORWARD_NG)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.gspaln_h_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _pack(self, problems, kind):
        n = len(problems)
        arr = (capi.GspalnHTask * max(n, 1))()
        keep = []
        for i, p in enumerate(problems):
            a = np.ascontiguousarray(p.a, np.uint8)
            b = np.ascontiguousarray(p.b, np.uint8)
            sg = np.ascontiguousarray(p.sgpt6, capi.SGPT6_DTYPE)
            if len(a) < p.a_right or len(b) < p.b_right or p.b_len < p.b_right or len(sg) < p.b_len + 2:
                raise ValueError("problem arrays shorter than the stated ranges")
            i53 = np.ascontiguousarray(p.int53, np.uint16) if p.int53 is not None else None
            cip = None
            if p.cip is not None:
                cip = np.ascontiguousarray(p.cip, np.int32)
                if len(cip) < 3 * p.a_right + 2:
                    raise ValueError("cip shorter than the stated range")
            keep.append((a, b, sg, i53, cip))
            t = arr[i]
            t.kind = kind
            t.a, t.b, t.sg = a.ctypes.data, b.ctypes.data, sg.ctypes.data
            t.int53 = i53.ctypes.data if i53 is not None else None
            t.cip = cip.ctypes.data if cip is not None else None
            t.b_len = int(p.b_len)
            t.a_left, t.a_right, t.b_left, t.b_right = p.a_left, p.a_right, p.b_left, p.b_right
            t.a_exgl, t.a_exgr, t.b_exgl, t.b_exgr = p.a_exgl, p.a_exgr, p.b_exgl, p.b_exgr
            t.lw, t.up = p.lw, p.up
            t.a_len = int(p.a_len or (len(a) - 1))
            cap = p.skl_cap or ((p.a_right - p.a_left) + (p.b_right - p.b_left) + 8)
            t.skl_cap = cap if kind in (capi.FORWARD_WIP, capi.FORWARD_NG) else 0
            t.n_imd = int(p.n_imd) if kind in (capi.HIRSCHBERG_WIP, capi.HIRSCHBERG_NG) else 0
        return arr, keep

    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.gspaln_h_last_error(self._h)
            raise EngineError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def _results(self, n, arr):
        res = (capi.GspalnResult * max(n, 1))()
        bufs = []
        for i in range(n):
            cap = arr[i].skl_cap
            buf = np.zeros((max(cap, 1), 2), np.int32)
            cp = np.zeros((arr[i].n_imd + 1, 10), np.int32) if arr[i].kind in (capi.HIRSCHBERG_WIP, capi.HIRSCHBERG_NG) else None
            bufs.append((buf, cp))
            res[i].skl = buf.ctypes.data if cap > 0 else None
            res[i].cpos = cp.ctypes.data if cp is not None else None
        return res, bufs

    def _collect(self, n, res, bufs, arr):
        return [Result(int(res[i].score), int(res[i].status),
                       bufs[i][0][:max(min(res[i].n_skl, arr[i].skl_cap), 0)].copy(), int(res[i].cells),
                       tuple(int(x) for x in res[i].ranges), bufs[i][1])
                for i in range(n)]

    def submit(self, problems, kind=capi.FORWARD_WIP):
        arr, keep = self._pack(problems, kind)
        n = len(problems)
        res, bufs = self._results(n, arr)
        self._check(self.lib.gspaln_h_submit(self._h, arr, n, res), "gspaln_h_submit")
        self._n = n
        return self._collect(n, res, bufs, arr)

    def forwardH1_wip(self, problems, trace=True):
        """trace=False == forwardH1_wip(0) as HomScoreH_ng calls it (score only)"""
        return self.submit(problems, capi.FORWARD_WIP if trace else capi.SCOREONLY_WIP)

    def pack(self, problems, kind=capi.FORWARD_WIP) -> PackedBatch:
        arr, keep = self._pack(problems, kind)
        return PackedBatch(arr, keep, len(problems))

    def submit_packed(self, batch: PackedBatch) -> PackedBatch:
        """host buffers in, host results out (scores / status / corners in `batch`)"""
        res = batch.res.ctypes.data_as(C.POINTER(capi.GspalnResult))
        self._check(self.lib.gspaln_h_submit(self._h, batch.arr, batch.n, res), "gspaln_h_submit")
        return batch

    def hirschbergH1_wip(self, problems):
        """problems carry n_imd; results carry score, narrowed ranges and cpos (Dim10 records)"""
        return self.submit(problems, capi.HIRSCHBERG_WIP)

    def lspH_ng(self, problems, max_vmf_space=32 * 1024 * 1024, sh=100, ubh=0, alg=2):
        """Aln2h1::lspH_ng over a batch (src/fwd2h1.cc:2134-2230).  Problems that take the
        Hirschberg route in the reference come back with status GSPALN_ST_UNSUPPORTED (3)."""
        arr, keep = self._pack(problems, capi.FORWARD_WIP)
        n = len(problems)
        res, bufs = self._results(n, arr)
        o = capi.GspalnLspOpts(int(max_vmf_space), int(sh), int(ubh), int(alg))
        self._check(self.lib.gspaln_h_lsp(self._h, arr, n, C.byref(o), res), "gspaln_h_lsp")
        self._n = n
        return self._collect(n, res, bufs, arr)

    def upload(self, problems, kind=capi.FORWARD_WIP):
        arr, keep = self._pack(problems, kind)
        self._check(self.lib.gspaln_h_upload(self._h, arr, len(problems)), "gspaln_h_upload")
        self._tasks, self._keep, self._n = arr, keep, len(problems)

    def run(self):
        self._check(self.lib.gspaln_h_run(self._h), "gspaln_h_run")

    def download(self):
        n = self._n
        res, bufs = self._results(n, self._tasks)
        self._check(self.lib.gspaln_h_download(self._h, res), "gspaln_h_download")
        return self._collect(n, res, bufs, self._tasks)

    def timing(self) -> Timing:
        t = capi.GspalnTiming()
        self.lib.gspaln_h_get_timing(self._h, C.byref(t))
        return Timing(t.h2d_ms, t.kernel_ms, t.d2h_ms, t.launches, t.h2d_bytes, t.d2h_bytes,
                      t.trace_bytes, t.cells)


# ---------------------------------------------------------------------------
# splice-signal scan of a genomic DNA segment (Exinon::intron53_c / intron53_n)
# ---------------------------------------------------------------------------
class ExinonScan:
    """Device mirror of what the reference's `Exinon` constructor computes for a DNA segment
    (src/codepot.cc:357-399, 437-523): per column the 5' / 3' splice signals (SGPT2 sig5, sig3)
    and the INT53 record.  `params` carries the two PSSMs and factors (see capi.make_scan_params)."""

    def __init__(self, params: dict, device: int = 0):
        self.lib = capi.load()
        self._sp, self._keep = capi.make_scan_params(params)
        self._h = C.c_void_p()
        rc = self.lib.gspaln_scan_create(C.byref(self._h), C.byref(self._sp), device)
        if rc != 0:
            raise EngineError(f"gspaln_scan_create failed ({rc}): no usable CUDA device or bad "
                              "parameters; the scan has no CPU fallback")
        self._len = 0

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.gspaln_scan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.gspaln_scan_last_error(self._h)
            raise EngineError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def _out(self, n):
        return np.zeros(n + 2, np.int16), np.zeros(n + 2, np.int16), np.zeros(n + 2, np.uint16)

    def scan(self, codes):
        """codes[i] == *Seq::at(i); returns sig5, sig3 (int16), int53 (uint16) by column 0 .. len + 1"""
        c = np.ascontiguousarray(codes, np.uint8)
        s5, s3, i53 = self._out(len(c))
        self._check(self.lib.gspaln_exinon_scan(self._h, c.ctypes.data, len(c), s5.ctypes.data,
                                                s3.ctypes.data, i53.ctypes.data), "gspaln_exinon_scan")
        self._len = len(c)
        return s5, s3, i53

    def upload(self, codes):
        c = np.ascontiguousarray(codes, np.uint8)
        self._check(self.lib.gspaln_scan_upload(self._h, c.ctypes.data, len(c)), "gspaln_scan_upload")
        self._len = len(c)

    def run(self):
        self._check(self.lib.gspaln_scan_run(self._h), "gspaln_scan_run")

    def download(self):
        s5, s3, i53 = self._out(self._len)
        self._check(self.lib.gspaln_scan_download(self._h, s5.ctypes.data, s3.ctypes.data, i53.ctypes.data),
                    "gspaln_scan_download")
        return s5, s3, i53

    def timing(self):
        a, b, c = C.c_float(0), C.c_float(0), C.c_float(0)
        self.lib.gspaln_scan_get_timing(self._h, C.byref(a), C.byref(b), C.byref(c))
        return {"h2d_ms": a.value, "kernel_ms": b.value, "d2h_ms": c.value}


class ExinonScanP(ExinonScan):
    """Protein-side scan (Exinon::intron53_p over a TRON segment): SGPT6 records + INT53"""

    def __init__(self, params: dict, device: int = 0):
        self.lib = capi.load()
        self._sp, self._keep = capi.make_scan_params_p(params)
        self._h = C.c_void_p()
        rc = self.lib.gspaln_scan_create_p(C.byref(self._h), C.byref(self._sp), device)
        if rc != 0:
            raise EngineError(f"gspaln_scan_create_p failed ({rc}): no usable CUDA device or parameters "
                              "outside this version's limits; the scan has no CPU fallback")
        self._len = 0

    def scan(self, tron):
        """tron[i] == *Seq::at(i); returns (SGPT6 records (capi.SGPT6_DTYPE, len + 2), int53)"""
        c = np.ascontiguousarray(tron, np.uint8)
        sg = np.zeros(len(c) + 2, capi.SGPT6_DTYPE)
        i53 = np.zeros(len(c) + 2, np.uint16)
        self._check(self.lib.gspaln_exinon_scan_p(self._h, c.ctypes.data, len(c), sg.ctypes.data,
                                                  i53.ctypes.data), "gspaln_exinon_scan_p")
        self._len = len(c)
        return sg, i53


def nuc2tron(gencode, codes_with_ends, device: int = 0):
    """Seq::nuc2tron on the device (src/seq.cc:774-798): codes_with_ends = at(-1 .. len) of a DNA
    segment; returns (tron codes of at(0 .. len - 1), kernel ms).  No CPU fallback."""
    lib = capi.load()
    g = np.ascontiguousarray(gencode, np.uint8)
    c = np.ascontiguousarray(codes_with_ends, np.uint8)
    if g.size != 64 or c.size < 2:
        raise ValueError("gencode must hold 64 entries, codes at least the two terminal residues")
    n = c.size - 2
    out = np.zeros(n, np.uint8)
    ms = C.c_float(0)
    rc = lib.gspaln_nuc2tron(device, g.ctypes.data, c.ctypes.data, n, out.ctypes.data, C.byref(ms))
    if rc != 0:
        raise EngineError(f"gspaln_nuc2tron failed ({rc}); there is no CPU fallback")
    return out, ms.value
