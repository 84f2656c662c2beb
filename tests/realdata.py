"""Real-data harness (test infrastructure): runs the reference's own binaries from oracle/_ref on
the sample data that ships with the reference (seqdb/dictdisc*: BASELINE.json configs[0]).

  spaln          the unmodified reference (CPU)
  spaln_gpu      the same program with the INTEGRATION.md patch: lsp*_ng / trcbkalign*_ng /
                 HomScore*_ng routed through libgspaln (include/gspaln_spaln_dropin.hpp)
  spaln_harvest  the same program, stock CPU code, dumping every top-level lsp*_ng call (inputs,
                 frozen parameters, score, corners) to a file

and parses the harvest files into Problem / ProblemH lists for the replay tests.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import struct
import subprocess
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref"
SEQDB = REF / "seqdb"


def available() -> bool:
    return all((REF / f).exists() for f in ("spaln", "spaln_gpu", "spaln_harvest")) and \
        (SEQDB / "dictdisc_g.gf.gz").exists() and (REF / "table" / "gnm2tab").exists()


class Workspace:
    """a scratch ALN_DBS directory holding the formatted genome (spaln -W: .bkn / .bkp block index)"""

    def __init__(self, protein=True, dna=True, threads=None):
        self.tmp = tempfile.TemporaryDirectory(prefix="gspaln_rd_")
        self.dir = Path(self.tmp.name)
        self.threads = threads or min(16, os.cpu_count() or 1)
        self.env = dict(os.environ, ALN_TAB=str(REF / "table"), ALN_DBS=str(self.dir))
        os.symlink(SEQDB / "dictdisc_g.gf.gz", self.dir / "dictdisc_g.gf.gz")
        for flag, on in (("-KP", protein), ("-KD", dna)):
            if on:
                subprocess.run([str(REF / "spaln"), "-W", flag, f"-t{self.threads}", "dictdisc_g.gf.gz"],
                               cwd=self.dir, env=self.env, check=True, stdout=subprocess.DEVNULL,
                               stderr=subprocess.DEVNULL)

    def head_fasta(self, src: Path, n: int) -> Path:
        """the first n records of a FASTA file"""
        out = self.dir / f"{src.stem}_{n}.fa"
        if not out.exists():
            k = 0
            with open(src) as f, open(out, "w") as g:
                for line in f:
                    if line.startswith(">"):
                        k += 1
                        if k > n:
                            break
                    g.write(line)
        return out

    def run(self, binary: str, opts: list, query: Path, harvest: Path = None, timeout=3600, stats=None) -> bytes:
        """stats: a dict that receives the drop-in's call counters (GSPALN_DROPIN_STATS), e.g.
        {"dna": {"lsp": 12, ...}, "protein": {...}}"""
        env = dict(self.env)
        if harvest is not None:
            env["GSPALN_HARVEST_FILE"] = str(harvest)
        if stats is not None:
            env["GSPALN_DROPIN_STATS"] = "1"
        r = subprocess.run([str(REF / binary)] + opts + ["-ddictdisc_g", str(query)], cwd=self.dir, env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
        if r.returncode != 0:
            raise RuntimeError(f"{binary} {' '.join(opts)} failed ({r.returncode}): {r.stderr[-400:].decode(errors='replace')}")
        if stats is not None:
            for ln in r.stderr.decode(errors="replace").splitlines():
                if ln.startswith("gspaln drop-in set-up:"):
                    stats["set-up"] = ln.split(":", 1)[1].strip()
                elif ln.startswith("gspaln drop-in ("):
                    kind = ln[len("gspaln drop-in ("):].split(")")[0]
                    stats[kind] = {k: int(v) for k, v in (x.split("=") for x in ln.split(":", 1)[1].split())}
        return r.stdout

    def close(self):
        self.tmp.cleanup()


# ---------------------------------------------------------------------------------------------
# harvest files (include/gspaln_spaln_dropin.hpp, GSPALN_HARVEST)
# ---------------------------------------------------------------------------------------------
class HarvestParams:
    def __init__(self, protein, payload):
        from spaln_b200 import capi
        self.protein = protein
        P = capi.GspalnHParams if protein else capi.GspalnParams
        at = C.sizeof(P)
        self.params = P.from_buffer_copy(payload[:at])
        self.opts = struct.unpack_from("<4i", payload, at)      # max_vmf_space, sh, ubh, alg
        at += 16
        self.sig53tab = np.frombuffer(payload, np.int16, 544, at).copy()
        at += 1088
        n_pen = struct.unpack_from("<i", payload, at)[0]
        at += 4
        self.penalty = np.frombuffer(payload, np.int16, n_pen, at).copy()
        at += 2 * n_pen
        if protein:
            self.spj_tabs = np.frombuffer(payload, np.uint8, 796, at).copy()
            at += 796
            self.minl, self.extragop, self.gw3l, self.noll = struct.unpack_from("<4i", payload, at)
        else:
            self.codonk1 = struct.unpack_from("<i", payload, at)[0]

    def engine(self, device=0):
        from spaln_b200 import Engine, EngineH
        if self.protein:
            e = EngineH(self.params, device=device)
            e.set_ng_tables(self.sig53tab, self.penalty, self.spj_tabs, self.minl, self.extragop,
                            self.gw3l, self.noll)
        else:
            e = Engine(self.params, device=device)
            e.set_ng_tables(self.sig53tab, self.penalty, self.codonk1)
        return e

    def lsp_kwargs(self):
        return dict(max_vmf_space=self.opts[0], sh=self.opts[1], ubh=self.opts[2], alg=self.opts[3])


def read_harvest(path):
    """yields ('params', HarvestParams) once, then ('call', dict) per harvested lsp*_ng call"""
    from spaln_b200 import capi
    protein = False
    with open(path, "rb") as f:
        while True:
            head = f.read(8)
            if len(head) < 8:
                return
            tag = head[:4].decode()
            n = struct.unpack("<i", head[4:])[0]
            payload = f.read(n)
            if tag in ("PRMS", "PRMH"):
                protein = tag == "PRMH"
                yield "params", HarvestParams(protein, payload)
                continue
            if tag != "CALL":
                continue
            h = struct.unpack_from("<16i", payload, 0)
            a_len, b_len, a_left, a_right, b_left, b_right = h[:6]
            exg = h[6:10]
            lw, up, a0, a1, b0, b1 = h[10:16]
            at = 64
            a = np.zeros(a_len + 1, np.uint8)
            a[a0:a1] = np.frombuffer(payload, np.uint8, a1 - a0, at)
            at += a1 - a0
            b = np.zeros(b_len + 1, np.uint8)
            b[b0:b1] = np.frombuffer(payload, np.uint8, b1 - b0, at)
            at += b1 - b0
            ncol = b1 + 1 - b0 + 1
            call = dict(a=a, b=b, a_len=a_len, b_len=b_len, a_left=a_left, a_right=a_right, b_left=b_left,
                        b_right=b_right, exg=exg, lw=lw, up=up)
            if protein:
                sg = np.zeros(b_len + 2, capi.SGPT6_DTYPE)
                sg[b0:b0 + ncol] = np.frombuffer(payload, capi.SGPT6_DTYPE, ncol, at)
                at += 14 * ncol
                call["sgpt6"] = sg
            else:
                s5 = np.zeros(b_len + 2, np.int16)
                s3 = np.zeros(b_len + 2, np.int16)
                s5[b0:b0 + ncol] = np.frombuffer(payload, np.int16, ncol, at)
                at += 2 * ncol
                s3[b0:b0 + ncol] = np.frombuffer(payload, np.int16, ncol, at)
                at += 2 * ncol
                call["sig5"], call["sig3"] = s5, s3
            i53 = np.zeros(b_len + 2, np.uint16)
            i53[b0:b0 + ncol] = np.frombuffer(payload, np.uint16, ncol, at)
            at += 2 * ncol
            call["int53"] = i53
            score, n_skl = struct.unpack_from("<2i", payload, at)
            at += 8
            call["score"] = score
            call["skl"] = np.frombuffer(payload, np.int32, 2 * n_skl, at).reshape(-1, 2).copy()
            yield "call", call


def to_problem(call, protein):
    from spaln_b200 import Problem, ProblemH
    common = dict(a=call["a"], b=call["b"], a_left=call["a_left"], a_right=call["a_right"],
                  b_left=call["b_left"], b_right=call["b_right"], lw=call["lw"], up=call["up"],
                  a_exgl=call["exg"][0], a_exgr=call["exg"][1], b_exgl=call["exg"][2], b_exgr=call["exg"][3],
                  int53=call["int53"], skl_cap=max(64, len(call["skl"]) + 8))
    if protein:
        return ProblemH(sgpt6=call["sgpt6"], b_len=call["b_len"], a_len=call["a_len"], **common)
    return Problem(sig5=call["sig5"], sig3=call["sig3"], **common)


def call_key(call):
    return (call["a_len"], call["b_len"], call["a_left"], call["a_right"], call["b_left"], call["b_right"],
            call["lw"], call["up"], tuple(call["exg"]), hashlib.sha1(call["a"].tobytes()).hexdigest()[:12])


def digest_calls(records):
    """order-independent digest of (problem geometry, score, corners) over harvested calls
    (the worker threads of the reference write them in any order)"""
    rows = sorted((call_key(c), int(c["score"]), c["skl"].tobytes()) for c in records)
    h = hashlib.sha256()
    for k, s, b in rows:
        h.update(repr(k).encode())
        h.update(struct.pack("<i", s))
        h.update(b)
    return h.hexdigest()
