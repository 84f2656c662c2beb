"""GPU: whole-program, real-data parity (BASELINE.json configs[0] and the cDNA sample).

1. The reference program itself, linked with the drop-in hooks (oracle/_ref/spaln_gpu: every
   lsp*_ng / trcbkalign*_ng / HomScore*_ng call of the run goes through SpalnEngine(H) -> coalescing
   queue -> C-ABI -> CUDA), must print byte-identical output to the stock CPU build on the sample
   data that ships with the reference: exon records / GFF come out of skl_rngS/H_ng
   (src/fwd2s1.cc:446-693, src/fwd2h1.cc:635-...) fed with OUR corner lists.
2. Every top-level lsp*_ng call of those runs, harvested from the stock CPU code
   (oracle/_ref/spaln_harvest), is replayed through gspaln_lsp / gspaln_h_lsp in batches: same
   score, same corner list.  The harvest is also pinned by a committed digest
   (tests/golden/realdata_digest.json).

Sizes: GSPALN_REALDATA_CDNA (default 1500 of the 5999 cDNAs; "all" for the whole file).
"""
import hashlib
import json
import os
from pathlib import Path

import numpy as np
import pytest

import realdata

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
DIGEST = ROOT / "tests" / "golden" / "realdata_digest.json"
PROT_MD5 = "708b9574c28764d8aa19fb8a05fd03cb"     # spaln -Q7 -O0 -A2 -Tdictdisc dictdisc.faa, SURVEY.md 8c


def n_cdna():
    v = os.environ.get("GSPALN_REALDATA_CDNA", "1500")
    return 5999 if v == "all" else int(v)


@pytest.fixture(scope="module")
def ws():
    if not realdata.available():
        pytest.skip("oracle/_ref drop-in binaries or sample data not built (make -C oracle ref dropin)")
    w = realdata.Workspace()
    yield w
    w.close()


def gff_records(out: bytes):
    """feature lines of a GFF3 output without the run-order dependent parts (ID= / Parent= serial
    numbers; '##' pragmas), sorted"""
    rows = []
    for ln in out.decode().splitlines():
        if ln.startswith("#"):
            continue
        f = ln.split("\t")
        if len(f) == 9:
            f[8] = ";".join(x for x in f[8].split(";") if not x.startswith(("ID=", "Parent=")))
        rows.append("\t".join(f))
    return sorted(rows)


def test_protein_sample_gff_identical_through_dropin(ws):
    """config 1: spaln -Q7 -O0 -A2 -Tdictdisc -ddictdisc_g dictdisc.faa"""
    opts = ["-Q7", "-O0", "-A2", "-t1", "-pq", "-Tdictdisc"]
    q = realdata.SEQDB / "dictdisc.faa"
    cpu = ws.run("spaln", opts, q)
    assert hashlib.md5(cpu).hexdigest() == PROT_MD5
    gpu = ws.run("spaln_gpu", opts, q)
    assert gpu == cpu
    # the same with the reference's worker threads feeding the coalescing queue.  The stock
    # program is itself not reproducible under -t8 (record numbering and the ##sequence-region
    # lines follow the thread interleaving), so records are compared without their serial numbers
    gpu8 = ws.run("spaln_gpu", ["-Q7", "-O0", "-A2", "-t8", "-pq", "-Tdictdisc"], q)
    assert gff_records(gpu8) == gff_records(cpu)


PROT_MD5_A0 = "387c8c749d45fccd7fbf6fde01e1071c"  # the same at -A0 (the reference's default mode), SURVEY.md 8c
PROT_MD5_A1 = "0a31a60ecbfba6f22b8c76609721b3ea"  # -A1


@pytest.mark.parametrize("alg,md5", [("-A0", PROT_MD5_A0), ("-A1", PROT_MD5_A1)])
def test_protein_sample_gff_identical_through_dropin_scalar_modes(ws, alg, md5):
    """`-A0`: lspH_ng runs on the device as a whole (forwardH_ng / hirschbergH_ng kernels), as does every
    other trcbkalignH_ng / HomScoreH_ng call.  `-A1`: only blocks with fewer than 8 query rows do."""
    opts = ["-Q7", "-O0", alg, "-t1", "-pq", "-Tdictdisc"]
    q = realdata.SEQDB / "dictdisc.faa"
    cpu = ws.run("spaln", opts, q)
    assert hashlib.md5(cpu).hexdigest() == md5
    st = {}
    assert ws.run("spaln_gpu", opts, q, stats=st) == cpu
    # the device answered: the whole driver at -A0 (61 lspH_ng calls on this sample, as at -A2), the
    # few-row blocks of -A1
    if alg == "-A0":
        assert st["protein"]["lsp"] >= 50, st
    else:
        assert st["protein"]["trcbk_exact"] >= 1 and st["protein"]["lsp"] == 0, st
    assert st["protein"]["trcbk_wip"] == 0, st
    assert st["protein"]["exact_no_tables"] == 0 and st["protein"]["exact_overflow"] == 0, st


@pytest.mark.parametrize("opts", [["-Q7", "-O4", "-S3", "-A2"], ["-Q7", "-O0", "-S3", "-A3"],
                                  ["-Q5", "-O4", "-S3", "-A2", "-LS"], ["-Q7", "-O4", "-S3", "-A0"],
                                  ["-Q7", "-O4", "-S3", "-A1"]])
def test_cdna_sample_identical_through_dropin(ws, opts):
    """cDNA sample (first n queries): exon coordinates (-O4) / GFF (-O0) of the drop-in build equal
    the stock build's, line for line"""
    scalar = "-A0" in opts or "-A1" in opts     # (slow stock modes: fewer queries)
    q = ws.head_fasta(realdata.SEQDB / "dictdisc.cf", min(200 if scalar else 600, n_cdna()))
    full = opts + [f"-t{ws.threads}", "-pq", "-Tdictdisc"]
    cpu = ws.run("spaln", full, q)
    st = {}
    gpu = ws.run("spaln_gpu", full, q, stats=st)
    assert len(cpu.splitlines()) > (150 if scalar else 500)
    if "-A1" not in opts:       # -A0: the driver with the exact-ILD kernels and the scalar Hirschberg pass
        assert st["dna"]["lsp"] >= 100, st
    if "-O0" in opts:
        assert gff_records(gpu) == gff_records(cpu)
    else:
        assert sorted(gpu.splitlines()) == sorted(cpu.splitlines())


def replay(path, device=0, batch=128):
    """harvested lsp*_ng calls through gspaln_lsp / gspaln_h_lsp; returns (n, mismatches, calls)"""
    eng = None
    hp = None
    pend = []
    n = bad = 0
    unsupported = 0
    digests = []

    def flush():
        nonlocal n, bad, unsupported
        if not pend:
            return
        probs = [realdata.to_problem(c, hp.protein) for c in pend]
        res = (eng.lspH_ng if hp.protein else eng.lspS_ng)(probs, **hp.lsp_kwargs())
        for c, r in zip(pend, res):
            n += 1
            if r.status == 3:
                unsupported += 1
            elif r.status != 0 or r.score != c["score"] or not np.array_equal(r.skl, c["skl"]):
                bad += 1
        pend.clear()

    for kind, rec in realdata.read_harvest(path):
        if kind == "params":
            hp = rec
            eng = hp.engine(device)
            continue
        digests.append((realdata.call_key(rec), int(rec["score"]), rec["skl"].tobytes()))
        pend.append(rec)
        if len(pend) >= batch:
            flush()
    flush()
    if eng is not None:
        eng.close()
    return n, bad, unsupported, digests


def digest_of(rows):
    import struct
    h = hashlib.sha256()
    for k, s, b in sorted(rows):
        h.update(repr(k).encode())
        h.update(struct.pack("<i", s))
        h.update(b)
    return h.hexdigest()


def harvest_runs(ws, prot, cq):
    """(tag, spaln options, query file) of the runs whose lsp*_ng calls are harvested"""
    runs = []
    for q in ("-Q7", "-Q6", "-Q5"):       # (-Q4 crashes the stock reference on this sample)
        runs.append((f"prot{q}", [q, "-O0", "-A2", "-t1", "-pq", "-Tdictdisc"], prot))
    runs.append(("prot-Q7A0", ["-Q7", "-O0", "-A0", "-t1", "-pq", "-Tdictdisc"], prot))
    for tag, o in (("Q7", ["-Q7", "-A2"]), ("Q4", ["-Q4", "-A2"]), ("Q7A3", ["-Q7", "-A3"]),
                   ("Q5LS", ["-Q5", "-A2", "-LS"]), ("Q7A0", ["-Q7", "-A0"])):
        runs.append((f"cdna{tag}", o + ["-O4", "-S3", f"-t{ws.threads}", "-pq", "-Tdictdisc"], cq))
    return runs


def test_harvested_lsp_calls_replayed_on_gpu(ws):
    """every lsp*_ng call of the real-data runs (stock CPU code, harvested) == gspaln_lsp /
    gspaln_h_lsp on the same inputs"""
    runs = harvest_runs(ws, realdata.SEQDB / "dictdisc.faa",
                        ws.head_fasta(realdata.SEQDB / "dictdisc.cf", n_cdna()))
    report = {}
    total = 0
    for tag, opts, query in runs:
        hv = ws.dir / f"{tag}.harvest"
        ws.run("spaln_harvest", opts, query, harvest=hv)
        n, bad, unsup, rows = replay(hv)
        hv.unlink()
        report[tag] = {"calls": n, "mismatches": bad, "unsupported": unsup, "digest": digest_of(rows)}
        total += n
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "realdata_replay.json").write_text(json.dumps({"n_cdna": n_cdna(), "runs": report}, indent=1))
    for tag, r in report.items():
        assert r["calls"] > 0, tag
        assert r["mismatches"] == 0 and r["unsupported"] == 0, (tag, r)
    if DIGEST.exists():
        pinned = json.loads(DIGEST.read_text())
        if pinned.get("n_cdna") == n_cdna():
            for tag, r in report.items():
                assert pinned["runs"][tag]["digest"] == r["digest"], f"reference harvest drifted: {tag}"
                assert pinned["runs"][tag]["calls"] == r["calls"], tag
    assert total >= 1000
