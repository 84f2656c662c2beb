"""CPU: the C oracle against the reference on REAL data -- every top-level lsp*_ng call of
`spaln -Q7 -A2` on the sample that ships with the reference (first cDNAs of seqdb/dictdisc.cf, all
proteins of dictdisc.faa), harvested from the stock CPU code by oracle/_ref/spaln_harvest, must
come out of the oracle's drivers (so_lsp / so_lsp_h) with the same score and corner list.
Low-complexity sequence, polyA tails, N runs and real PSSM hits that the synthetic fixtures lack."""
import numpy as np
import pytest

import oracle_harness as O
import realdata


def params_dict(hp):
    p = hp.params
    d = dict(BasicGOP=p.gop, BasicGEP=p.gep, LongGOP=p.lgop, LongGEP=p.lgep, GapWI=p.ipen, llmt=p.llmt,
             nquant=p.nquant, quant_len=list(p.quant_len), quant_pen=list(p.quant_pen), avmch=p.avmch,
             spj=p.spj, simdim=p.simdim, penalty=hp.penalty, sig53tab=hp.sig53tab,
             MaxVmfSpace=hp.opts[0], sh=hp.opts[1], ubh=hp.opts[2], alg=hp.opts[3])
    d["simmtx"] = np.array(list(p.simmtx), np.int32)[: p.simdim * p.simdim]
    if hp.protein:
        d.update(lcl=p.lcl, codonk1=p.codonk1, GapW1=p.gw1, GapW2=p.gw2, GapW3=p.gw3, GapE1=p.gape1,
                 GapE2=p.gape2, spj_tabs=hp.spj_tabs, minl=hp.minl, ExtraGOP=hp.extragop, GapW3L=hp.gw3l,
                 Noll=hp.noll)
    else:
        d.update(lcl=16 if p.local else 0, Noll=p.noll, GapPenalty1=p.gappen1, codonk1=hp.codonk1)
    return d


def oracle_task(c, protein):
    t = dict(a=np.concatenate([[0], c["a"]]).astype(np.uint8), b=np.concatenate([[0], c["b"]]).astype(np.uint8),
             a_left=c["a_left"], a_right=c["a_right"], b_left=c["b_left"], b_right=c["b_right"],
             a_exgl=c["exg"][0], a_exgr=c["exg"][1], b_exgl=c["exg"][2], b_exgr=c["exg"][3],
             lw=c["lw"], up=c["up"], int53=c["int53"])
    if protein:
        sg = c["sgpt6"]
        t["sgpt6"] = np.stack([sg[k].astype(np.int16) for k in
                               ("sig5", "sig3", "sigS", "sigT", "sigE", "sigI", "phs5", "phs3")], axis=1)
        t["blen"], t["alen"] = c["b_len"], c["a_len"]
    else:
        t["sig5"], t["sig3"] = c["sig5"], c["sig3"]
    return t


@pytest.fixture(scope="module")
def ws():
    if not realdata.available():
        pytest.skip("oracle/_ref drop-in binaries or sample data not built (make -C oracle ref dropin)")
    w = realdata.Workspace()
    yield w
    w.close()


@pytest.mark.parametrize("kind", ["cdna", "protein", "cdna_A0", "protein_A0"])
def test_oracle_reproduces_harvested_lsp_calls(ws, kind):
    if kind.startswith("cdna"):
        q = ws.head_fasta(realdata.SEQDB / "dictdisc.cf", 120)
        opts = ["-Q7", "-O4", "-S3", "-A0" if kind == "cdna_A0" else "-A2", f"-t{ws.threads}", "-pq", "-Tdictdisc"]
    else:
        q = realdata.SEQDB / "dictdisc.faa"
        opts = ["-Q7", "-O0", "-A0" if kind == "protein_A0" else "-A2", "-t1", "-pq", "-Tdictdisc"]
    hv = ws.dir / f"oracle_{kind}.harvest"
    ws.run("spaln_harvest", opts, q, harvest=hv)
    prm = None
    n = bad = 0
    for what, rec in realdata.read_harvest(hv):
        if what == "params":
            prm = params_dict(rec)
            protein = rec.protein
            continue
        t = oracle_task(rec, protein)
        o = (O.lsp_h if protein else O.lsp)(prm, t, cap=1 << 15)
        n += 1
        if o["unsupported"] or o["score"] != rec["score"] or not np.array_equal(o["skl"], rec["skl"]):
            bad += 1
    assert n >= 40
    assert bad == 0, f"{bad} of {n} harvested {kind} calls differ"
