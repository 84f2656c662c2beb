#!/usr/bin/env python
"""bench.py -- throughput of the spliced-alignment DP hot path.

  python bench.py --gpus N --steps K --warmup W            (our engine, CUDA)
  python bench.py --impl reference --gpus N --steps K ...  (reference CPU SIMD path)

Workload (BASELINE.json configs[1]): synthetic cDNAs of 1-3 kb, each against its
genomic locus (+- 0.5-5 kb flanks), DNA->DNA spliced alignment, band = stripe()
with the default shoulder.  One step = one pass of SimdAln2s1::forwardS1_wip
semantics (score + trace-back corners) over the whole query set.  Metric: GCUPS
(band cells of the scalar reference loop per second, 1e9).

Rank layout: one process per GPU; queries are sharded by rank (independent
problems, no data-path collective), weak scaling (fixed queries per GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

# frozen reference parameters for `-Q0 -A2 -yX0 -TDictyost` as dumped by the
# reference itself into tests/golden/dna_A2_global.npz (prm_* keys)
PARAM_FIXTURE = "dna_A2_global"
REF_OPTS = "-Q0 -A2 -S1 -yX0 -TDictyost"
METRIC = "GCUPS (spliced-DP band cells/s, 1e9) forwardS1_wip, 10k synthetic cDNA 1-3 kb vs genomic loci"
B_CELL = 2.0    # algorithmic bytes per cell: 1 B trace + 16 B per (column x 16-row strip), DESIGN.md


def load_params():
    import golden_io
    prm, _ = golden_io.load(PARAM_FIXTURE)
    return prm


SEED = 20251017
_LETTERS = np.zeros(256, np.uint8)
for _c, _ch in ((2, "A"), (3, "C"), (5, "G"), (9, "T")):
    _LETTERS[_c] = ord(_ch)


def make_workload(n_queries, seed=SEED, first=0):
    """problems [first, first + n) of the seeded global config-2 query set (with the FASTA strings
    the reference arm needs)"""
    from spaln_b200 import workload
    out = []
    for i in range(first, first + n_queries):
        r = workload.config2_problem_seeded(seed, i)
        r["int53"] = workload.synthetic_int53(r["b"])
        out.append(r)
    return out


def with_strings(raw):
    for r in raw:
        if "genome_str" not in r:
            r["genome_str"] = _LETTERS[r["b"]].tobytes().decode()
            r["query_str"] = _LETTERS[r["a"]].tobytes().decode()
    return raw


def to_problems(raw):
    from spaln_b200 import Problem, workload
    # int53 (site classes, as Exinon::intron53_c derives them from the residues): blocks with fewer
    # than 8 query rows that the lsp driver cuts go to the scalar exact-ILD kernel, as in the reference
    return [Problem(int53=r["int53"] if r.get("int53") is not None else workload.synthetic_int53(r["b"]),
                    a=r["a"], b=r["b"], sig5=r["sig5"], sig3=r["sig3"], a_left=r["a_left"],
                    a_right=r["a_right"], b_left=r["b_left"], b_right=r["b_right"], lw=r["lw"],
                    up=r["up"], a_exgl=1, a_exgr=1, b_exgl=1, b_exgr=1, skl_cap=512) for r in raw]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# newest first: the packed kernel the config-2 batch runs on (PK = 8), then the round-1 32-bit kernel
TRAFFIC_FILES = ("r02_traffic.json", "r01_traffic.json")
NCU_FILES = ("r02_ncu_full_dp_wip_pk8_q3000.json", "r02_ncu_full_dp_wip_pk4_q3000.json", "r01_ncu_full_dp_wip_q3000.json")


def measured_traffic_per_cell():
    """DRAM bytes per cell of the DP kernel from the committed `ncu --set full` capture"""
    for name in TRAFFIC_FILES:
        try:
            return float(json.loads((ROOT / "profiles" / name).read_text())["dram_bytes_per_cell"]), name
        except Exception:
            continue
    return None, None


def binding_resource():
    """what binds the DP kernel according to the committed `ncu --set full` capture (the HBM
    roofline is reported as the contract asks, but the kernel is integer-ALU bound)"""
    for name in NCU_FILES:
        try:
            d = json.loads((ROOT / "profiles" / name).read_text())[0]
            return {"resource": "integer ALU pipe (packed int16x2 max / add / select recurrences)",
                    "pipe_alu_pct_of_peak": d["pipe_alu_pct"]["value"], "pipe_fma_pct_of_peak": d["pipe_fma_pct"]["value"],
                    "issue_slots_busy_pct": d["issue_slots_busy_pct"]["value"],
                    "warp_instructions_per_cell": d["warp_instructions"]["value"] / d["cells"],
                    "thread_instructions_per_cell": d["warp_instructions"]["value"] * d["active_threads_per_warp_inst"]["value"] / d["cells"],
                    "source": "profiles/" + name}
        except Exception:
            continue
    return None


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


# ---------------------------------------------------------------------------
# reference CPU path (oracle/_ref): the one place bench.py executes oracle/
# ---------------------------------------------------------------------------
def reference_run(raw, steps, warmup, threads, lsp=False):
    """Times SimdAln2s1::forwardS1_wip of the unmodified reference (AVX2 build)
    on `raw` problems with `threads` host threads.  Returns (gcups per step list,
    cells per step, results of last step)."""
    import ref_harness as R
    if not R.available():
        return None
    ref = R.Reference._instance or R.Reference(REF_OPTS)
    tasks = []
    for r in raw:
        t = ref.task(r["genome_str"], r["query_str"])
        ex = t.export()
        assert np.array_equal(ex["a"][1:-1], r["a"]) and np.array_equal(ex["b"][1:-1], r["b"])
        t.inject(r["sig5"], r["sig3"])
        tasks.append(t)
    cells = sum(int(r["cells"]) for r in raw)
    out = [None] * len(tasks)
    # shared work queue, largest problems first: every host thread stays busy until the sample is
    # done (the ctypes call releases the GIL), as the reference's own pthread queue does
    order = sorted(range(len(tasks)), key=lambda i: -int(raw[i]["cells"]))
    nxt = [0]
    lock = threading.Lock()

    def work():
        while True:
            with lock:
                k = nxt[0]
                nxt[0] += 1
            if k >= len(order):
                return
            i = order[k]
            if lsp:     # the whole driver: trace-back vs UDH dispatch at the default -V
                out[i] = tasks[i].lsp(raw[i]["lw"], raw[i]["up"], cap=4096)
            else:
                out[i] = tasks[i].kernel(raw[i]["lw"], raw[i]["up"], 0, cap=4096)

    times = []
    for s in range(warmup + steps):
        nxt[0] = 0
        t0 = time.perf_counter()
        th = [threading.Thread(target=work) for _ in range(threads)]
        [x.start() for x in th]
        [x.join() for x in th]
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    for t in tasks:
        t.close()
    return times, cells, out


# ---------------------------------------------------------------------------
# protein x genome leg (BASELINE.json configs[2] shape: proteins of 300-800 aa against their
# genomic loci); secondary report next to the headline config-2 numbers
# ---------------------------------------------------------------------------
PROT_FIXTURE = "prot_A2_global"
PROT_REF_OPTS = "-Q0 -A2 -yX0 -TDictyost"
B_CELL_H = 3.5  # 2 B trace + (8 B band r/w + 16 B column record) per (column x 16-row strip), DESIGN.md


def make_protein_workload(n, seed):
    from spaln_b200 import workload
    rng = np.random.default_rng(seed)
    return [workload.protein_problem(rng, plen_range=(300, 800), flank=(500, 5000), sh=100)
            for _ in range(n)]


def to_problems_h(raw):
    from spaln_b200 import ProblemH
    out = [ProblemH.from_export(r, r["lw"], r["up"]) for r in raw]
    for p in out:
        p.skl_cap = 512
    return out


def protein_reference_child(nsample, seed, threads, out_path):
    """runs in its own process (the reference keeps one option string per process): times
    SimdAln2h1::forwardH1_wip on the first `nsample` problems of the protein workload"""
    import ref_harness as R
    if not R.available():
        Path(out_path).write_text(json.dumps({"unavailable": "oracle/_ref not built"}))
        return 0
    import golden_io
    import oracle_harness as O
    raw_all = make_protein_workload(nsample, seed)
    # fhlastH1 of the reference can return a start point right of b_right (its last-column scan
    # moves mx but not maxr, src/fwd2h1_simd.h:764-788) and then walks outside its trace buffer
    # (segfault).  Such problems are screened out of the CPU sample with the oracle, which stops
    # there; the GPU arm keeps them (and reports the same lone corner).
    prm, _ = golden_io.load_protein(PROT_FIXTURE)
    keep = [i for i, r in enumerate(raw_all)
            if O.forward_h1_wip(prm, r)["skl"][0][1] <= r["b_right"]]
    raw = [raw_all[i] for i in keep]
    nsample = len(raw)
    ref = R.Reference(PROT_REF_OPTS, protein=True)
    tasks = []
    for r in raw:
        t = ref.task(r["genome"], r["query"])
        ex = t.export_p()
        assert np.array_equal(ex["a"][1:-1], r["a"][1:-1]) and np.array_equal(ex["b"], r["b"])
        t.inject_p(r["sgpt6"])
        tasks.append(t)
    out = [None] * nsample

    def work(tid):
        for i in range(tid, nsample, threads):
            out[i] = tasks[i].kernel_p(raw[i]["lw"], raw[i]["up"], 0, cap=4096)

    times = []
    for s in range(2):
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
        [x.start() for x in th]
        [x.join() for x in th]
        times.append(time.perf_counter() - t0)
    Path(out_path).write_text(json.dumps({
        "seconds": times[-1], "index": keep, "scores": [int(o["score"]) for o in out],
        "skl": [o["skl"].tolist() for o in out]}))
    return 0


def host_cells_h(raw):
    import ctypes as C
    from spaln_b200 import capi
    lib = capi.load()
    for r in raw:
        t = capi.GspalnHTask()
        t.a_left, t.a_right, t.b_left, t.b_right = r["a_left"], r["a_right"], r["b_left"], r["b_right"]
        t.lw, t.up = r["lw"], r["up"]
        r["cells"] = int(lib.gspaln_h_task_cells(C.byref(t)))


def protein_leg(args, local_rank, rank, ncores, barrier, with_cpu):
    from spaln_b200 import EngineH
    import golden_io
    prm, _ = golden_io.load_protein(PROT_FIXTURE)
    seed = 20251017 + 104729 * rank
    raw = make_protein_workload(args.protein_queries, seed)
    host_cells_h(raw)
    probs = to_problems_h(raw)
    cells = sum(r["cells"] for r in raw)
    eng = EngineH(prm, device=local_rank)
    eng.upload(probs)
    eng.run()
    barrier()
    ks = []
    for _ in range(2):
        eng.run()
        ks.append(eng.timing().kernel_ms)
    res = eng.download()
    k_ms = float(np.mean(ks))
    packed = eng.pack(probs)                # task descriptors (metadata) marshalled once
    eng.submit(probs[: max(1, len(probs) // 50)])
    barrier()
    t0 = time.perf_counter()
    # host numpy buffers -> derive column records + pinned pack -> H2D -> kernel (+ walk) -> D2H
    eng.submit_packed(packed)
    barrier()
    e2e_s = time.perf_counter() - t0
    tm = eng.timing()
    peak, _ = measured_peak()
    leg = {"note": "SimdAln2h1::forwardH1_wip semantics (score + corners), proteins 300-800 aa x genomic "
                   "locus (+-0.5-5 kb), band = stripe31(sh=100), synthetic SGPT6 table; this rank",
           "queries": args.protein_queries, "cells": cells, "kernel_ms": k_ms,
           "gcups": cells / (k_ms * 1e-3) / 1e9, "queries_per_s": args.protein_queries / (k_ms * 1e-3),
           "e2e_gcups": cells / e2e_s / 1e9, "e2e_ms": 1e3 * e2e_s,
           "h2d_bytes": int(tm.h2d_bytes), "d2h_bytes": int(tm.d2h_bytes),
           "e2e_phases_ms": {"h2d_first_chunk": tm.h2d_ms, "kernel_span": tm.kernel_ms, "d2h": tm.d2h_ms},
           "status_nonzero": sum(1 for r in res if r.status != 0),
           "roofline": {"bound": "hbm", "bytes_per_cell": B_CELL_H, "unit": "GB/s",
                        "achieved": cells * B_CELL_H / (k_ms * 1e-3) / 1e9, "peak": peak,
                        "frac": cells * B_CELL_H / (k_ms * 1e-3) / 1e9 / peak, "kernel": "dp_h1_kernel<true>"}}
    if with_cpu:
        nsample = min(args.cpu_sample, len(raw))
        tmp = ROOT / "gpurun_out"
        tmp.mkdir(exist_ok=True)
        outp = tmp / f"_prot_ref_{os.getpid()}.json"
        try:
            subprocess.run([sys.executable, str(ROOT / "bench.py"), "--leg", "protein-cpu", "--cpu-sample",
                            str(nsample), "--leg-seed", str(seed), "--leg-out", str(outp)],
                           check=True, timeout=900)
            r = json.loads(outp.read_text())
            outp.unlink()
        except Exception as e:      # the report line must still be printed
            r = {"unavailable": repr(e)}
        if "seconds" in r:
            idx = r["index"]
            sc = sum(raw[i]["cells"] for i in idx)
            mism = sum(1 for k, i in enumerate(idx)
                       if r["scores"][k] != res[i].score or
                       not np.array_equal(np.array(r["skl"][k], np.int32).reshape(-1, 2), res[i].skl))
            leg["cpu_baseline"] = {"value": sc / r["seconds"] / 1e9, "unit": "GCUPS", "cores": ncores,
                                   "kind": "reference",
                                   "sample": f"{len(idx)} of the first {nsample} problems ({sc / 1e6:.0f} Mcells; the rest "
                                             "crash the reference: start point outside its trace buffer), "
                                             f"SimdAln2h1::forwardH1_wip AVX2 build, {ncores} threads",
                                   "parity_mismatches_on_sample": mism}
        else:
            leg["cpu_baseline"] = r
    eng.close()
    return leg


def scan_leg(prm, with_cpu):
    """the splice-signal scan that fills the Exinon tables (sig5, sig3, INT53)
    of a genomic segment, on a 100 Mb synthetic genome (the genome size of BASELINE config 2).
    Kernel time from CUDA events with the segment resident; e2e with host buffers; roofline against
    HBM with 7 algorithmic bytes per position (1 B residue in, 2 + 2 + 2 B out)."""
    from spaln_b200 import ExinonScan
    n = 100_000_000
    rng = np.random.default_rng(20251017)
    codes = rng.choice(np.array([2, 3, 5, 9], np.uint8), size=n, p=[0.295, 0.205, 0.205, 0.295])
    sc = ExinonScan(prm, device=0)
    sc.upload(codes)
    ms = []
    for i in range(6):
        sc.run()
        if i >= 3:
            ms.append(sc.timing()["kernel_ms"])
    t0 = time.perf_counter()
    got = sc.scan(codes)
    e2e_s = time.perf_counter() - t0
    k_ms = float(np.mean(ms))
    peak, peak_kind = measured_peak()
    out = {"note": "Exinon::intron53_c + intron53_n (PatMat::calcPatMat, two Markov-order-2 PSSMs) over a "
                   "100 Mb synthetic genome, one launch; bit-identical shorts",
           "positions": n, "kernel_ms": k_ms, "gnt_per_s": n / k_ms / 1e6,
           "e2e_ms": 1e3 * e2e_s, "e2e_gnt_per_s": n / e2e_s / 1e9, "h2d_bytes": n, "d2h_bytes": 6 * (n + 2),
           "roofline": {"bound": "hbm", "bytes_per_position": 7.0, "achieved": 7.0 * n / k_ms / 1e6,
                        "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": 7.0 * n / k_ms / 1e6 / peak,
                        "kernel": "exinon_scan_fast_kernel<8, 18>",
                        "note": "issue-slot bound: 2 + 8 + 18 dependent fp32 table look-ups per position "
                                "(reference operation order), see DESIGN.md"}}
    # the protein-side preparation of the same genome: Seq::nuc2tron (1 B in, 1 B out per position)
    from spaln_b200 import nuc2tron
    z = np.load(ROOT / "tests" / "golden" / "nuc2tron.npz")
    ends = np.concatenate([[0], codes, [0]]).astype(np.uint8)
    tron, t_ms = nuc2tron(z["gencode"], ends)
    out["nuc2tron"] = {"note": "Seq::nuc2tron over the same 100 Mb, kernel time with the segment resident",
                       "kernel_ms": t_ms, "gnt_per_s": n / t_ms / 1e6,
                       "roofline": {"bound": "hbm", "bytes_per_position": 2.0, "achieved": 2.0 * n / t_ms / 1e6,
                                    "peak": peak, "unit": "GB/s", "frac": 2.0 * n / t_ms / 1e6 / peak,
                                    "kernel": "nuc2tron_kernel"}}
    # ... and the protein-side scan of the tron segment (Exinon::intron53_p): SGPT6 records + INT53
    from spaln_b200 import ExinonScanP
    zp = np.load(ROOT / "tests" / "golden" / "scan_p.npz")
    prm_p = {k[4:]: (zp[k] if zp[k].ndim else zp[k].item()) for k in zp.files if k.startswith("prm_")}
    scp = ExinonScanP(prm_p, device=0)
    t0 = time.perf_counter()
    sg, i53p = scp.scan(tron)
    p_e2e = time.perf_counter() - t0
    p_ms = scp.timing()["kernel_ms"]
    scp.close()
    out["protein_scan"] = {"note": "Exinon::intron53_p over the 100 Mb tron segment (4 PSSMs, 5th-order coding "
                                   "potential, phases): generic one-thread-per-column kernel",
                           "kernel_ms": p_ms, "gnt_per_s": n / p_ms / 1e6, "e2e_ms": 1e3 * p_e2e,
                           "roofline": {"bound": "hbm", "bytes_per_position": 17.0, "achieved": 17.0 * n / p_ms / 1e6,
                                        "peak": peak, "unit": "GB/s", "frac": 17.0 * n / p_ms / 1e6 / peak,
                                        "kernel": "exinon_scan_p_kernel"}}
    if with_cpu:
        sys.path.insert(0, str(ROOT / "tests"))
        import oracle_harness
        kp = 5_000_000
        t0 = time.perf_counter()
        op = oracle_harness.exinon_scan_p(prm_p, tron[:kp])
        cpu_p = time.perf_counter() - t0
        names = ("sig5", "sig3", "sigS", "sigT", "sigE", "sigI", "phs5", "phs3")
        okp = all(np.array_equal(sg[nm][: kp - 64].astype(np.int16), op["sgpt6"][: kp - 64, c])
                  for c, nm in enumerate(names))
        out["protein_scan"]["cpu_baseline"] = {"value": kp / cpu_p / 1e9, "unit": "Gnt/s", "cores": 1, "kind": "port",
                                               "sample": "first 5 Mb", "parity_on_sample": bool(okp)}
        kk = 20_000_000
        t0 = time.perf_counter()
        ot = oracle_harness.nuc2tron(z["gencode"], ends[: kk + 2])
        out["nuc2tron"]["cpu_baseline"] = {"value": kk / (time.perf_counter() - t0) / 1e9, "unit": "Gnt/s",
                                           "cores": 1, "kind": "port", "sample": "first 20 Mb",
                                           "parity_on_sample": bool(np.array_equal(ot[:-1], tron[: kk - 1]))}
        k = 10_000_000
        t0 = time.perf_counter()
        o = oracle_harness.exinon_scan(prm, codes[:k])
        cpu_s = time.perf_counter() - t0
        ok = all(np.array_equal(x[:k - 64], y[:k - 64]) for x, y in zip(got, (o["sig5"], o["sig3"], o["int53"])))
        out["cpu_baseline"] = {"value": k / cpu_s / 1e9, "unit": "Gnt/s", "cores": 1, "kind": "port",
                               "sample": "first 10 Mb of the genome through oracle/spaln_oracle_scan.c",
                               "parity_on_sample": bool(ok)}
    sc.close()
    return out


def config4_leg(args, ncores, with_cpu):
    """BASELINE config 4 shape on one GPU: mRNA of ~2.5 kb against loci with 20x longer introns
    (tens of kb), local mode (-LS), through the driver (Aln2s1::lspS_ng) at the default -V: every
    problem takes the multi-intermediate Hirschberg route + block re-alignments."""
    import golden_io
    from spaln_b200 import Engine, Problem, workload
    prm, _ = golden_io.load("dna_A2_local")
    rng = np.random.default_rng(20251017 + 4)
    nq = max(16, args.queries // 5)
    raw = [workload.config2_problem(rng, qlen_range=(1500, 3500), intron_scale=20.0) for _ in range(nq)]
    problems = to_problems(raw)
    host_cells(raw)
    cells = sum(r["cells"] for r in raw)
    eng = Engine(prm, device=0)
    opts = dict(max_vmf_space=32 * 1024 * 1024, sh=int(prm["sh"]), alg=2)
    pk = eng.pack(problems)                                 # task descriptors marshalled once, as in `e2e`
    eng.lsp_packed(pk, **opts)                              # one whole untimed pass: the grow-only pools get their size
    t0 = time.perf_counter()
    eng.lsp_packed(pk, **opts)
    dt = time.perf_counter() - t0
    tm = eng.timing()
    from spaln_b200 import Result
    res = [Result(int(pk.scores[i]), int(pk.status[i]), pk.corners(i).copy(), 0) for i in range(nq)]
    out = {"note": "config-4 shaped problems (mRNA 1.5-3.5 kb, introns x20: loci of tens of kb), -LS, "
                   "Aln2s1::lspS_ng at -V 32 MiB (Hirschberg route), wall clock with host buffers",
           "queries": nq, "root_cells": cells, "queries_per_s": nq / dt, "gcups_root_cells": cells / dt / 1e9,
           "kernel_ms": tm.kernel_ms, "total_ms": 1e3 * dt, "launches": tm.launches,
           "device_cells": tm.cells, "status_nonzero": sum(1 for r in res if r.status != 0),
           "mean_locus_nt": float(np.mean([r["b_right"] for r in raw]))}
    eng.close()
    if with_cpu:
        import ref_harness
        k = min(8, nq)
        child = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--leg", "config4-ref", "--leg-seed", str(k),
                                "--queries", str(args.queries)], capture_output=True, text=True)
        try:
            c = json.loads(child.stdout.strip().splitlines()[-1])
            mism = sum(1 for i in range(k) if res[i].status != 0 or c["scores"][i] != res[i].score or
                       not np.array_equal(np.array(c["skl"][i], np.int32).reshape(-1, 2), res[i].skl))
            out["cpu_baseline"] = {"queries_per_s": c["queries_per_s"], "gcups_root_cells": c["gcups"],
                                   "cores": c["cores"], "kind": "reference",
                                   "sample": f"first {k} problems, Aln2s1::lspS_ng of the AVX2 build, -LS, "
                                             f"{c['cores']} threads (own process: the reference keeps one option "
                                             "string per process)",
                                   "parity_mismatches_on_sample": mism}
        except Exception as e:      # the reference arm is a reported baseline, never the product
            out["cpu_baseline"] = {"value": None, "error": f"{type(e).__name__}: {child.stderr[-200:]}"}
    return out


def a0_leg(args, ncores, with_cpu):
    """The reference's DEFAULT mode (-A0: exact intron-length scoring, scalar kernels) on config-2
    problems through the driver: Aln2s1::lspS_ng with algmode.alg == 0 at -V 32 MiB -- exact-ILD
    trace-backs (forwardS_ng), the scalar Hirschberg pass (hirschbergS_ng) for the problems above the
    space limit, blocks banded by the pass.  Wall clock with host buffers."""
    import golden_io
    from spaln_b200 import Engine, Result
    prm, _ = golden_io.load("dna_A0_udh")
    nq = max(64, args.queries // 10)
    raw = make_workload(nq, SEED + 7)
    problems = to_problems(raw)
    host_cells(raw)
    cells = sum(r["cells"] for r in raw)
    eng = Engine(prm, device=0)
    opts = dict(max_vmf_space=32 * 1024 * 1024, sh=int(prm["sh"]), alg=0)
    pk = eng.pack(problems)
    eng.lsp_packed(pk, **opts)                              # one whole untimed pass: the grow-only pools get their size
    t0 = time.perf_counter()
    eng.lsp_packed(pk, **opts)
    dt = time.perf_counter() - t0
    tm = eng.timing()
    res = [Result(int(pk.scores[i]), int(pk.status[i]), pk.corners(i).copy(), 0) for i in range(nq)]
    out = {"note": "the reference's default mode -A0 on config-2 problems: Aln2s1::lspS_ng (alg 0) at -V 32 MiB -- "
                   "exact-ILD kernels dp_xild_kernel / dp_xudh_kernel, wall clock with host buffers",
           "queries": nq, "root_cells": cells, "queries_per_s": nq / dt, "gcups_root_cells": cells / dt / 1e9,
           "kernel_ms": tm.kernel_ms, "total_ms": 1e3 * dt, "launches": tm.launches, "device_cells": tm.cells,
           "status_nonzero": sum(1 for r in res if r.status != 0)}
    eng.close()
    if with_cpu:
        k = min(nq, 2 * ncores)
        child = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--leg", "a0-ref", "--leg-seed", str(k),
                                "--queries", str(args.queries)], capture_output=True, text=True)
        try:
            c = json.loads(child.stdout.strip().splitlines()[-1])
            mism = sum(1 for i in range(k) if res[i].status != 0 or c["scores"][i] != res[i].score or
                       not np.array_equal(np.array(c["skl"][i], np.int32).reshape(-1, 2), res[i].skl))
            out["cpu_baseline"] = {"queries_per_s": c["queries_per_s"], "gcups_root_cells": c["gcups"],
                                   "cores": c["cores"], "kind": "reference",
                                   "sample": f"first {k} problems, Aln2s1::lspS_ng of the AVX2 build at -A0 "
                                             f"(scalar forwardS_ng / hirschbergS_ng), {c['cores']} threads, own process",
                                   "parity_mismatches_on_sample": mism}
        except Exception as e:      # the reference arm is a reported baseline, never the product
            out["cpu_baseline"] = {"value": None, "error": f"{type(e).__name__}: {child.stderr[-200:]}"}
    return out


def a0_reference_child(args):
    """child process: the reference's own lspS_ng at -A0 on the first k problems of the -A0 leg"""
    import threading
    import ref_harness
    k = args.leg_seed
    nq = max(64, args.queries // 10)
    raw = with_strings(make_workload(nq, SEED + 7))[:k]
    ref = ref_harness.Reference("-Q0 -A0 -S1 -yX0 -TDictyost")
    tasks = []
    for r in raw:
        t = ref.task(r["genome_str"], r["query_str"])
        t.inject(r["sig5"], r["sig3"])
        tasks.append(t)
    ncores = os.cpu_count() or 1
    outs = [None] * k
    nxt = [0]
    lock = threading.Lock()

    def work():
        while True:
            with lock:
                i = nxt[0]
                nxt[0] += 1
            if i >= k:
                return
            outs[i] = tasks[i].lsp(raw[i]["lw"], raw[i]["up"], cap=1 << 17)

    t0 = time.perf_counter()
    th = [threading.Thread(target=work) for _ in range(min(ncores, k))]
    [x.start() for x in th]
    [x.join() for x in th]
    dt = time.perf_counter() - t0
    host_cells(raw)
    cells = sum(r["cells"] for r in raw)
    print(json.dumps({"queries_per_s": k / dt, "gcups": cells / dt / 1e9, "cores": min(ncores, k),
                      "scores": [o["score"] for o in outs], "skl": [o["skl"].tolist() for o in outs]}))
    return 0


def config4_reference_child(args):
    """child process: the reference's own lspS_ng (-LS) on the first k config-4 problems"""
    import threading
    import ref_harness
    from spaln_b200 import workload
    k = args.leg_seed
    rng = np.random.default_rng(20251017 + 4)
    nq = max(16, args.queries // 5)
    raw = [workload.config2_problem(rng, qlen_range=(1500, 3500), intron_scale=20.0) for _ in range(nq)][:k]
    ref = ref_harness.Reference("-Q0 -A2 -S1 -yX0 -LS -TDictyost")
    tasks = []
    for r in raw:
        t = ref.task(r["genome_str"], r["query_str"])
        t.inject(r["sig5"], r["sig3"])
        tasks.append(t)
    ncores = os.cpu_count() or 1
    outs = [None] * k
    nxt = [0]
    lock = threading.Lock()

    def work():
        while True:
            with lock:
                i = nxt[0]
                nxt[0] += 1
            if i >= k:
                return
            outs[i] = tasks[i].lsp(raw[i]["lw"], raw[i]["up"], cap=1 << 17)

    t0 = time.perf_counter()
    th = [threading.Thread(target=work) for _ in range(min(ncores, k))]
    [x.start() for x in th]
    [x.join() for x in th]
    dt = time.perf_counter() - t0
    host_cells(raw)
    cells = sum(r["cells"] for r in raw)
    print(json.dumps({"queries_per_s": k / dt, "gcups": cells / dt / 1e9, "cores": min(ncores, k),
                      "scores": [o["score"] for o in outs], "skl": [o["skl"].tolist() for o in outs]}))
    return 0


def config5_sweep(args, prm, rank, local_rank, world):
    """BASELINE config 5: band width x query length sweep (256 - 64k DP cells per task), trace-back
    kernel (forwardS1_wip semantics), `--sweep-tasks` tasks per point and rank.  Per point: kernel
    GCUPS with the tasks resident (CUDA events), end-to-end GCUPS with host buffers, HBM roofline
    fraction.  Tasks: random query of m residues against a locus that contains it with 0 / 1 / 2
    planted GT..AG introns; band = stripe() with the shoulder that makes the band W diagonals wide."""
    import ctypes as C
    import torch
    from spaln_b200 import Engine, Problem, capi, workload
    rng = np.random.default_rng(SEED + 5 + 1000 * rank)
    eng = Engine(prm, device=local_rank)
    peak, peak_kind = measured_peak()
    points = []
    distinct = 1024
    for W in (16, 32, 64, 128, 256, 512, 1024):
        for m in (16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
            if not (256 <= m * W <= 65536):
                continue
            if args.sweep_only and args.sweep_only != f"{W},{m}":
                continue
            probs = []
            for k in range(distinct):
                nin = k % 3 if (W >= 64 and m >= 32) else 0
                intr = [int(x) for x in rng.integers(20, max(21, W // 4), size=nin)]
                q = workload.random_dna(rng, m)
                cuts = np.sort(rng.choice(np.arange(4, m - 4), size=nin, replace=False)) if nin else []
                parts, pos = [], 0
                for c, il in zip(cuts, intr):
                    it = workload.random_dna(rng, il)
                    it[:2] = np.frombuffer(b"GT", np.uint8)
                    it[-2:] = np.frombuffer(b"AG", np.uint8)
                    parts += [q[pos:c], it]
                    pos = c
                parts.append(q[pos:])
                gseq = np.concatenate(parts)
                a, b = workload.DNA_CODE[q], workload.DNA_CODE[gseq]
                s5, s3 = workload.synthetic_signals(b, rng)
                sh = max(1, (W - (len(b) - len(a))) // 2)
                lw, up = workload.stripe(0, len(a), 0, len(b), sh)
                probs.append(Problem(a=a, b=b, sig5=s5, sig3=s3, a_left=0, a_right=len(a), b_left=0,
                                     b_right=len(b), lw=lw, up=up, skl_cap=24))
            arr, keep = eng._pack(probs, capi.FORWARD_WIP)
            reps = max(1, args.sweep_tasks // distinct)
            n = distinct * reps
            big = (capi.GspalnTask * n)()
            C.memmove(big, arr, C.sizeof(capi.GspalnTask) * distinct)
            for r in range(1, reps):        # the same host buffers, independent device problems
                C.memmove(C.byref(big, r * distinct * C.sizeof(capi.GspalnTask)), arr,
                          C.sizeof(capi.GspalnTask) * distinct)
            cells = sum(int(eng.lib.gspaln_task_cells(C.byref(arr[i]))) for i in range(distinct)) * reps
            from spaln_b200.engine import PackedBatch
            batch = PackedBatch(big, keep, n)
            eng._check(eng.lib.gspaln_upload(eng._h, big, n), "gspaln_upload")
            ks = []
            for i in range(4):
                eng.run()
                if i:
                    ks.append(eng.timing().kernel_ms)
            eng.submit_packed(batch)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            eng.submit_packed(batch)
            e2e_s = time.perf_counter() - t0
            bad = int(np.count_nonzero(batch.status))
            k_ms = float(np.mean(ks))
            gc = cells / (k_ms * 1e-3) / 1e9
            points.append({"W": W, "m": m, "tasks": n, "cells_per_task": cells // n, "kernel_ms": k_ms,
                           "gcups": gc, "e2e_gcups": cells / e2e_s / 1e9, "tasks_per_s": n / (k_ms * 1e-3),
                           "roofline_frac": gc * B_CELL / peak, "status_nonzero": bad})
    eng.close()
    return points


def _shape_chunk(args):
    """worker: problems [lo, hi) step `stride` of a config-3 / config-4 job (seeded per problem)"""
    cfg, seed, idx = args
    from spaln_b200 import workload
    out = []
    for i in idx:
        rng = np.random.default_rng([seed, cfg, int(i)])
        if cfg == 3:
            r = workload.protein_problem(rng, plen_range=(300, 800), flank=(500, 5000), sh=100)
            r.pop("genome"); r.pop("query")
        else:
            r = workload.config2_problem(rng, qlen_range=(1500, 3500), intron_scale=20.0)
            r["int53"] = workload.synthetic_int53(r["b"])
            r.pop("genome_str"); r.pop("query_str")
        r["index"] = int(i)
        out.append(r)
    return out


def shape_job(args, cfg, rank, local_rank, world, ncores):
    """BASELINE configs 3 / 4 at their stated scale as ONE query-sharded job over N GPUs (weak
    scaling: 100 000 proteins resp. 200 000 mRNAs over 8 GPUs = 12 500 / 25 000 per GPU by default):
    problem i of the job belongs to rank i mod N and is built there from its own seed (the loci of a
    3 Gb genome do not fit one host buffer here, so nothing is broadcast in this mode); every rank
    runs its share -- config 3: SimdAln2h1::forwardH1_wip semantics (score + corners), config 4:
    Aln2s1::lspS_ng at the default -V with -LS, i.e. the multi-intermediate Hirschberg route + block
    re-alignments -- and the GeneRecord hit records are gathered on rank 0."""
    import multiprocessing as mp
    per_gpu = args.queries if args.queries != 10000 else (12500 if cfg == 3 else 25000)
    total = per_gpu * world
    mine = np.arange(rank, total, world)
    chunks = [(cfg, SEED, mine[k:k + 128]) for k in range(0, len(mine), 128)]
    procs = max(1, min(ncores // max(world, 1), 32))
    if procs > 1:
        with mp.get_context("fork").Pool(procs) as pool:
            raw = [r for part in pool.map(_shape_chunk, chunks) for r in part]
    else:
        raw = [r for c in chunks for r in _shape_chunk(c)]

    import torch
    import torch.distributed as dist
    import golden_io
    from spaln_b200 import Engine, EngineH, shard
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if cfg == 3:
        prm, _ = golden_io.load_protein(PROT_FIXTURE)
        host_cells_h(raw)
        problems = to_problems_h(raw)
        eng = EngineH(prm, device=local_rank)
        bcell, kname = B_CELL_H, "dp_h1_kernel<true>"
    else:
        prm, _ = golden_io.load("dna_A2_local")
        host_cells(raw)
        problems = to_problems(raw)
        eng = Engine(prm, device=local_rank)
        bcell, kname = 1.5, "dp_udh_kernel + dp_wip_kernel<true, LOCAL> (lspS_ng levels)"
        lsp_opts = dict(max_vmf_space=32 * 1024 * 1024, sh=int(prm["sh"]), alg=2)
    cells = sum(r["cells"] for r in raw)
    qlen = np.array([r["a_right"] for r in raw], np.int64)

    def run_once():
        """host buffers in -> hit records on rank 0; returns (kernel ms, device cells, phases)"""
        t0 = time.perf_counter()
        packed = eng.pack(problems)
        if cfg == 3:
            eng.submit_packed(packed)
        else:
            eng.lsp_packed(packed, **lsp_opts)
        tm = eng.timing()
        t1 = time.perf_counter()
        if world > 1:
            dist.barrier()
        t2 = time.perf_counter()
        n_skl = np.minimum(packed.res["n_skl"][:packed.n], np.diff(packed.off)).astype(np.int64)
        src = np.repeat(packed.off[:-1], n_skl) + (np.arange(int(n_skl.sum())) - np.repeat(np.cumsum(n_skl) - n_skl, n_skl))
        corners = packed.skl[src]
        hits = shard.make_hits(mine, packed.scores, n_skl, qlen, corners, min_intron=int(prm["llmt"]))
        got = shard.gather_hit_records(hits, corners, 0, dev)
        t3 = time.perf_counter()
        bad = int(np.count_nonzero(packed.status))
        return tm, got, bad, (t1 - t0, t2 - t1, t3 - t2)

    # warm-up: one whole untimed step, so that the grow-only pools (pinned staging of the whole batch,
    # device pools) have their size before the clock starts -- as in the config-2 job
    run_once()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    steps = max(1, min(args.steps, 2))
    kms, ph = 0.0, np.zeros(3)
    for _ in range(steps):
        tm, got, bad, p3 = run_once()
        kms += tm.kernel_ms
        ph += p3
    barrier()
    e2e_s = (time.perf_counter() - t0) / steps
    clocks = sampler.stop()
    kms /= steps
    ph /= steps
    if world > 1:
        t = torch.tensor([kms, e2e_s] + ph.tolist(), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kms, e2e_s = float(t[0]), float(t[1])
        ph = np.array(t.tolist()[2:])
        c = torch.tensor([cells, bad], dtype=torch.int64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        cells_total, bad = int(c[0]), int(c[1])
    else:
        cells_total = cells
    if rank == 0:
        assert got is not None and len(got[0]) == total
        peak, peak_kind = measured_peak()
        name = ("config3: proteins 300-800 aa x genomic locus (+-0.5-5 kb), forwardH1_wip semantics, aa x nt cells"
                if cfg == 3 else
                "config4: mRNA 1.5-3.5 kb x locus with 20x introns (tens of kb), -LS, lspS_ng at -V 32 MiB "
                "(Hirschberg route + block re-alignments), root cells")
        emit({"metric": f"GCUPS ({name})", "value": cells_total / (kms * 1e-3) / 1e9, "unit": "GCUPS",
              "n_gpus": world, "steps": steps, "warmup": 1, "ms_per_step": kms, "higher_is_better": True,
              "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
              "config": {"workload": name, "queries_per_gpu": per_gpu, "queries_total": total,
                         "parallelism": f"ONE job: problem i -> rank i mod {world}, GeneRecord hit records gathered on rank 0"},
              "queries_per_s": total / e2e_s, "clocks": clocks, "status_errors": bad,
              "e2e": {"value": cells_total / e2e_s / 1e9, "unit": "GCUPS", "queries_per_s": total / e2e_s,
                      "h2d_bytes_per_step": int(tm.h2d_bytes), "d2h_bytes_per_step": int(tm.d2h_bytes),
                      "hit_records_on_rank0": total,
                      "phases_ms": {"marshal+pack+h2d+kernels+d2h": 1e3 * ph[0], "wait_for_slowest_rank": 1e3 * ph[1],
                                    "hit_records+gather": 1e3 * ph[2], "total": 1e3 * e2e_s}},
              "gpu_launches": int(tm.launches) * steps,
              "roofline": {"bound": "hbm", "bytes_per_cell": bcell, "unit": "GB/s", "peak": peak, "peak_kind": peak_kind,
                           "achieved": cells / (kms * 1e-3) / 1e9 * bcell, "frac": cells / (kms * 1e-3) / 1e9 * bcell / peak,
                           "kernel": kname, "traffic": None}})
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def host_cells(raw):
    import ctypes as C
    from spaln_b200 import capi
    lib = capi.load()
    for r in raw:
        t = capi.GspalnTask()
        t.a_left, t.a_right, t.b_left, t.b_right = r["a_left"], r["a_right"], r["b_left"], r["b_right"]
        t.lw, t.up = r["lw"], r["up"]
        r["cells"] = int(lib.gspaln_task_cells(C.byref(t)))


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL: "NCCL version ...") write to fd 1; the contract is ONE JSON line on
    stdout, so everything else is sent to stderr and the line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gspaln", choices=["gspaln", "reference"])
    ap.add_argument("--queries", type=int, default=10000, help="queries per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=32, help="problems in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--protein-queries", type=int, default=3000,
                    help="problems of the protein x genome leg per GPU (0 = skip the leg)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="2: the headline workload (default); 3 / 4: the protein and the long-intron mRNA "
                         "jobs at their stated scale; 5: band-width x query-length sweep")
    ap.add_argument("--sweep-tasks", type=int, default=102400, help="tasks per sweep point and rank")
    ap.add_argument("--sweep-only", default="", help="W,m: run this point of the config-5 sweep alone (profiling)")
    ap.add_argument("--leg", default="", help=argparse.SUPPRESS)
    ap.add_argument("--leg-seed", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--leg-out", default="", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.leg == "protein-cpu":
        return protein_reference_child(args.cpu_sample, args.leg_seed, os.cpu_count() or 1, args.leg_out)
    if args.leg == "config4-ref":
        return config4_reference_child(args)
    if args.leg == "a0-ref":
        return a0_reference_child(args)
    _quiet_stdout()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = args.gpus
    prm = load_params()
    ncores = os.cpu_count() or 1
    workload_name = (f"config2: {args.queries} synthetic cDNA 1-3 kb x genomic locus (+-0.5-5 kb), "
                     "DNA spliced, band=stripe(sh=100), -A2 semantics, trace-back kernel (-V raised)")

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        # a bounded sample of the same workload: >= 8 problems per host thread from a shared queue
        nsample = max(args.cpu_sample, 8 * ncores)
        raw = with_strings(make_workload(nsample))
        host_cells(raw)
        r = reference_run(raw, args.steps, max(1, args.warmup), ncores)
        if r is None:
            emit({"impl": "reference", "unavailable": "oracle/_ref not built"})
            return 0
        times, cells, _ = r
        ms = 1e3 * float(np.mean(times))
        val = cells / (ms * 1e-3) / 1e9
        one_raw = with_strings(make_workload(6, first=nsample))
        host_cells(one_raw)
        one = reference_run(one_raw, 1, 0, 1)
        val1 = sum(x["cells"] for x in one_raw) / float(np.mean(one[0])) / 1e9
        line = {
            "metric": METRIC, "value": val, "unit": "GCUPS", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": max(1, args.warmup), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": workload_name, "sample": f"{nsample} problems of the workload per step"},
            "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": ncores, "kind": "reference",
                             "per_core": val / ncores, "one_thread": val1,
                             "sample": f"{nsample} config-2 problems ({cells / 1e6:.0f} Mcells) per step from a "
                                       f"shared queue (largest first), SimdAln2s1::forwardS1_wip AVX2, "
                                       f"{ncores} threads; one_thread = 6 further problems on one thread"},
            "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        emit(line)
        return 0

    if args.config in (3, 4) and args.impl != "reference":
        return shape_job(args, args.config, rank, local_rank, world, ncores)
    if args.config == 5:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        pts = config5_sweep(args, prm, rank, local_rank, world)
        if world > 1:
            # weak scaling: every rank runs its own tasks of each point; a point ends when the slowest rank does
            t = torch.tensor([[p["kernel_ms"], p["tasks"] * p["cells_per_task"] / p["e2e_gcups"] / 1e9] for p in pts],
                             dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            for p, (k, e) in zip(pts, t.tolist()):
                cells = p["tasks"] * p["cells_per_task"] * world
                p.update(kernel_ms=k, gcups=cells / (k * 1e-3) / 1e9, e2e_gcups=cells / e / 1e9,
                         tasks=p["tasks"] * world, tasks_per_s=p["tasks"] * world / (k * 1e-3))
                p["roofline_frac"] = p["gcups"] / world * B_CELL / measured_peak()[0]
            dist.destroy_process_group()
        if rank == 0:
            worst = min(pts, key=lambda p: p["gcups"])
            best = max(pts, key=lambda p: p["gcups"])
            emit({"metric": "GCUPS per (band width W, query length m) point, forwardS1_wip trace-back kernel",
                  "value": float(np.exp(np.mean(np.log([p["gcups"] for p in pts])))), "unit": "GCUPS (geometric mean over points)",
                  "n_gpus": n_gpus, "steps": 3, "warmup": 1, "higher_is_better": True, "scaling": "weak",
                  "vs_baseline": None, "dtype": "int16", "data": "synthetic",
                  "config": {"workload": f"config5: band x length sweep, {args.sweep_tasks} tasks per point and GPU, "
                                         "256 <= m*W <= 65536 cells per task"},
                  "worst_point": worst, "best_point": best, "sweep": pts})
        return 0

    # ------------------------------------------------------------------ our arm
    # ONE job: rank 0 owns the formatted genome (the concatenated loci) and the query set of
    # queries_per_gpu x N problems (weak scaling), generated before CUDA is touched (worker
    # processes are forked)
    from spaln_b200 import shard, workload
    total_q = args.queries * world
    g = workload.config2_global(total_q, SEED, procs=min(ncores, 48)) if rank == 0 else None

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        print("bench.py: no CUDA device; the DP engine has no CPU fallback", file=sys.stderr)
        return 2
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    KEYS = ("a", "b", "lens", "sig5", "sig3", "int53")

    def job_inputs():
        """formatted genome + query set + index (what a run broadcasts once), then the derived
        per-column tables (in a real run each rank computes them with the scan kernels)"""
        bufs = shard.broadcast_buffers([g[k] for k in KEYS] if rank == 0 else [None] * len(KEYS), 0, dev)
        return dict(zip(KEYS, bufs))

    def all_cells(lens):
        """DP cells of every problem of the job (gspaln_task_cells), for the partition"""
        import ctypes as C
        from spaln_b200 import capi
        lib = capi.load()
        t = capi.GspalnTask()
        out = np.empty(len(lens), np.int64)
        for i, (la, lb) in enumerate(lens.tolist()):
            t.a_left, t.a_right, t.b_left, t.b_right = 0, la, 0, lb
            t.lw, t.up = workload.stripe(0, la, 0, lb, 100)
            out[i] = lib.gspaln_task_cells(C.byref(t))
        return out

    from spaln_b200 import Engine
    G = job_inputs()
    cells_all = all_cells(G["lens"])
    mine = shard.lpt_partition(cells_all, world)[rank]
    raw = workload.global_problems(G, mine)
    for r, c in zip(raw, cells_all[mine]):
        r["cells"] = int(c)
    problems = to_problems(raw)
    cells_step = int(cells_all[mine].sum())
    cells_job = int(cells_all.sum())
    eng = Engine(prm, device=local_rank)

    # ---- device-resident throughput (`value`): this rank's share, inputs in HBM
    eng.upload(problems)
    for _ in range(args.warmup):
        eng.run()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    kern_ms = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.run()                                   # blocks until the stream drains
        kern_ms.append(eng.timing().kernel_ms)      # CUDA events on the launching stream
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    tm = eng.timing()
    launches_per_step = tm.launches
    dev_s = sum(kern_ms) * 1e-3
    res = eng.download()
    bad = sum(1 for r in res if r.status != 0)

    # ---- end to end (`e2e`): the whole job through the public API with HOST buffers, every
    # step = broadcast of the formatted genome + query set from rank 0 (N > 1), cell-balanced
    # partition, plan + marshal + pinned pack + H2D + kernels (+ walk) + D2H, hit records
    # (GeneRecord headers + corners) gathered on rank 0
    e2e_steps = max(1, min(args.steps, 5))
    qlen = np.array([r["a_right"] for r in raw], np.int64)

    # rank 0 keeps the formatted genome, the query set and the index in pinned staging buffers
    staged = [torch.from_numpy(g[k]).pin_memory() for k in KEYS[:3]] if (rank == 0 and world > 1) else [None] * 3

    def job_step():
        t_a = time.perf_counter()
        if world > 1:
            # genome + queries + index from rank 0's pinned memory to every GPU's HBM (where the scan /
            # DP kernels of a run consume them)
            shard.broadcast_buffers(staged, 0, dev, to_host=False)
            torch.cuda.synchronize()
        t_b = time.perf_counter()
        part = shard.lpt_partition(cells_all, world)[rank]
        t_c = time.perf_counter()
        # task descriptors marshalled inside the call; the result area of the previous step is reused
        packed = eng.pack_global(G, part, reuse=job_step.prev)
        job_step.prev = packed
        eng.submit_packed(packed)
        t_d = time.perf_counter()
        if world > 1:
            dist.barrier()                          # load imbalance shows up here, not in the gather
        t_w = time.perf_counter()
        n_skl = np.minimum(packed.res["n_skl"][:packed.n], np.diff(packed.off)).astype(np.int64)
        lens_c = np.repeat(np.cumsum(n_skl) - n_skl, n_skl)
        src = np.repeat(packed.off[:-1], n_skl) + (np.arange(int(n_skl.sum())) - lens_c)
        corners = packed.skl[src]
        hits = shard.make_hits(part, packed.scores, n_skl, qlen, corners, min_intron=int(prm["llmt"]))
        got = shard.gather_hit_records(hits, corners, 0, dev)
        t_e = time.perf_counter()
        return packed, got, (t_b - t_a, t_c - t_b, t_d - t_c, t_e - t_w, t_w - t_d)

    job_step.prev = None
    eng.submit(problems[: max(1, len(problems) // 50)])     # warm the pinned/device pools
    job_step()                                              # one untimed full step (host threads, page tables)
    barrier()
    phases = np.zeros(5)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        packed, got, ph = job_step()
        phases += ph
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    phases /= e2e_steps
    tm2 = eng.timing()
    n_hits = len(got[0]) if got is not None else 0
    if rank == 0:
        assert n_hits == total_q and np.array_equal(got[0]["Rid"], np.arange(total_q))
    res2 = eng.forwardS1_wip(problems[: args.cpu_sample])   # sample kept as objects for the parity check

    # ---- the driver path (lspS_ng dispatch at the reference's default -V = 32 MiB):
    # Hirschberg passes + block re-alignments for the larger problems, host in the loop
    lsp_opts = dict(max_vmf_space=32 * 1024 * 1024, sh=int(prm["sh"]), alg=2)
    eng.lspS_ng(problems, **lsp_opts)   # warm-up: pools of this path
    barrier()
    t0 = time.perf_counter()
    packed3 = eng.pack(problems)
    eng.lsp_packed(packed3, **lsp_opts)
    barrier()
    lsp_s = time.perf_counter() - t0
    tm3 = eng.timing()
    lsp_bad = int(np.count_nonzero(packed3.status))
    from spaln_b200 import Result
    res3 = [Result(int(packed3.scores[i]), int(packed3.status[i]), packed3.corners(i).copy(), 0)
            for i in range(min(args.cpu_sample, len(problems)))]      # sample kept for the parity check

    prot = None
    if args.protein_queries > 0:
        prot = protein_leg(args, local_rank, rank, ncores, barrier,
                           with_cpu=(n_gpus == 1 and not args.no_cpu_baseline))

    if world > 1:
        t = torch.tensor([dev_s, wall, e2e_s] + phases.tolist(), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, wall, e2e_s = [float(x) for x in t.tolist()[:3]]
        phases = np.array(t.tolist()[3:])
        c = torch.tensor([cells_step, bad], dtype=torch.int64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        cells_total, bad = int(c[0]), int(c[1])
    else:
        cells_total = cells_step
    assert cells_total == cells_job

    if rank == 0:
        ms_per_step = 1e3 * dev_s / args.steps
        value = cells_total / (ms_per_step * 1e-3) / 1e9
        peak, peak_kind = measured_peak()
        k_ms = float(np.mean(kern_ms))
        achieved = cells_step * B_CELL / (k_ms * 1e-3) / 1e9       # GB/s of this rank's kernel
        line = {
            "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16", "data": "synthetic",
            "config": {"workload": workload_name, "queries_per_gpu": args.queries, "queries_total": total_q,
                       "cells_per_step_per_gpu": cells_step, "cells_per_step": cells_total,
                       "l2_policy": "inputs+trace per step far larger than L2 (no flush needed)",
                       "parallelism": f"ONE job of {total_q} queries: rank 0 broadcasts the formatted genome + "
                                      f"query set (NCCL), LPT partition by gspaln_task_cells over {n_gpus} rank(s), "
                                      "no data-path collective, GeneRecord hit records gathered on rank 0"},
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "queries_per_s": args.queries * n_gpus / (ms_per_step * 1e-3),
            "clocks": clocks,
            "e2e": {"value": cells_total / e2e_s / 1e9, "unit": "GCUPS",
                    "h2d_bytes_per_step": int(tm2.h2d_bytes), "d2h_bytes_per_step": int(tm2.d2h_bytes),
                    "queries_per_s": args.queries * n_gpus / e2e_s, "steps": e2e_steps,
                    "hit_records_on_rank0": n_hits,
                    "phases_ms": {"bcast": 1e3 * phases[0], "partition": 1e3 * phases[1],
                                  "marshal+pack+h2d+kernel+d2h": 1e3 * phases[2],
                                  "hit_records+gather": 1e3 * phases[3],
                                  "wait_for_slowest_rank": 1e3 * phases[4],
                                  "h2d_first_chunk": tm2.h2d_ms, "kernel": tm2.kernel_ms, "d2h": tm2.d2h_ms,
                                  "total": 1e3 * e2e_s},
                    "note": "max over ranks; per-rank bytes are this rank's"},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": (measured_traffic_per_cell()[0] * cells_step
                                     if measured_traffic_per_cell()[0] else None),
                         "traffic_source": f"profiles/{measured_traffic_per_cell()[1]} (ncu dram__bytes per cell x cells of this launch)",
                         "peak_kind": peak_kind,
                         "bytes_per_cell": B_CELL, "kernel": "dp_wip_kernel<TRACE, PK> (packed int16x2)",
                         "kernel_ms": k_ms,
                         "note": "integer-ALU bound DP: see DESIGN.md (HBM roof is not the binding one)",
                         "binding": binding_resource()},
            "status_errors": bad,
            "lsp_path": {"note": "Aln2s1::lspS_ng dispatch at -V 32 MiB (trace-back or multi-intermediate "
                                 "Hirschberg + block re-alignment), wall clock, this rank",
                         "queries_per_s": args.queries / lsp_s, "gcups_root_cells": cells_step / lsp_s / 1e9,
                         "kernel_ms": tm3.kernel_ms, "total_ms": 1e3 * lsp_s, "launches": tm3.launches,
                         "device_cells": tm3.cells, "status_nonzero": lsp_bad},
        }
        if prot is not None:
            line["protein_path"] = prot
        if n_gpus == 1:
            line["scan_path"] = scan_leg(prm, with_cpu=not args.no_cpu_baseline)
            line["config4_path"] = config4_leg(args, ncores, with_cpu=not args.no_cpu_baseline)
            line["a0_path"] = a0_leg(args, ncores, with_cpu=not args.no_cpu_baseline)
        if world > 1:
            line["bcast_ms"] = 1e3 * phases[0]
            line["gather_ms"] = 1e3 * phases[3]
        if n_gpus == 1 and not args.no_cpu_baseline:
            nsample = min(max(args.cpu_sample, 8 * ncores), len(raw))
            sample = with_strings(raw[:nsample])
            npar = min(args.cpu_sample, nsample)    # problems whose results are compared
            r = reference_run(sample, 1, 1, ncores)
            if r is not None:
                times, cells, out = r
                # the reference results must equal ours on the sample
                mism = sum(1 for i in range(npar)
                           if out[i]["score"] != res2[i].score or not np.array_equal(out[i]["skl"], res2[i].skl))
                v = cells / float(np.mean(times)) / 1e9
                line["cpu_baseline"] = {
                    "value": v, "unit": "GCUPS", "cores": ncores, "per_core": v / ncores,
                    "kind": "reference",
                    "sample": f"first {nsample} problems of the step ({cells / 1e6:.0f} Mcells) from a shared "
                              f"queue (largest first), SimdAln2s1::forwardS1_wip AVX2 build, {ncores} threads; "
                              f"results of the first {npar} compared with the GPU's",
                    "parity_mismatches_on_sample": mism}
                # the same sample through the reference's own driver (default -V)
                r = reference_run(sample, 1, 0, ncores, lsp=True)
                times, cells, out = r
                mism = sum(1 for i in range(npar) if res3[i].status != 0 or
                           out[i]["score"] != res3[i].score or not np.array_equal(out[i]["skl"], res3[i].skl))
                line["cpu_baseline"]["lsp_path"] = {
                    "queries_per_s": nsample / float(np.mean(times)),
                    "gcups_root_cells": cells / float(np.mean(times)) / 1e9,
                    "parity_mismatches_on_sample": mism}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "GCUPS", "cores": ncores, "kind": "reference",
                                        "sample": "oracle/_ref not available"}
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
