#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(oracle/_ref/libspaln_ref.so, built by oracle/Makefile from /root/reference).

Each fixture holds, for one reference option string, the frozen parameters,
the raw inputs of a set of DP problems (query / genome codes, Exinon splice
signal table, ranges, end-gap flags, band) and the reference's outputs for
SimdAln2s1::forwardS1_wip (score + trace-back corners) and scoreonlyS1_wip.

Run in the build container only (needs /root/reference at build time of
oracle/_ref):   python tests/golden/make_golden.py
The reference keeps its parameters in process globals, so every option string
is generated in its own subprocess.
"""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))

CONFIGS = {
    # name: (reference options, seed)
    "dna_A2_global": ("-Q0 -A2 -S1 -yX0 -TDictyost", 11),
    "dna_A2_local": ("-Q0 -A2 -S1 -yX0 -LS -TDictyost", 12),
    "dna_A3_global": ("-Q0 -A3 -S1 -yX0 -TDictyost", 13),
    "dna_A2_tetrapod": ("-Q0 -A2 -S1 -TTetrapod", 14),
    # double affine gap penalty (alprm.ls = 3 -> PwdB::Noll == 3)
    "dna_A2_dagp": ("-Q0 -A2 -S1 -yX0 -yl3 -TDictyost", 18),
    # small -V forces the unidirectional-Hirschberg path on small inputs (SURVEY section 4 pin d)
    "dna_A2_udh": ("-Q0 -A2 -S1 -yX0 -V256K -TDictyost", 15),
    "dna_A2_udh_local": ("-Q0 -A2 -S1 -yX0 -V256K -LS -TDictyost", 16),
    "dna_A6_udh_recursive": ("-Q0 -A6 -S1 -yX0 -V128K -TDictyost", 17),
    # the reference's default mode (-A0): scalar kernels, scalar Hirschberg pass hirschbergS_ng
    "dna_A0_udh": ("-Q0 -A0 -S1 -yX0 -V256K -TDictyost", 41),
    "dna_A0_udh_local": ("-Q0 -A0 -S1 -yX0 -V256K -LS -TDictyost", 42),
    "dna_A0_udh_dagp": ("-Q0 -A0 -S1 -yX0 -yl3 -V256K -TDictyost", 43),
    "prot_A0_udh": ("-Q0 -A0 -yX0 -V128K -TDictyost", 44),
    "prot_A0_udh_local": ("-Q0 -A0 -yX0 -V128K -LS -TDictyost", 45),
    # intron positions annotated on the query (Cip_score, src/gsinfo.h:127-139): read by the
    # exact-ILD kernels only; small -V so that the driver reaches them through block re-alignment
    "dna_A2_cip": ("-Q0 -A2 -S1 -yX0 -V256K -TDictyost", 31),
    "prot_A2_cip": ("-Q0 -A2 -yX0 -V128K -TDictyost", 32),
    # protein query x genomic segment (SimdAln2h1::forwardH1_wip)
    "prot_A2_global": ("-Q0 -A2 -yX0 -TDictyost", 21),
    "prot_A2_local": ("-Q0 -A2 -yX0 -LS -TDictyost", 22),
    # small -V: Aln2h1::lspH_ng takes the Hirschberg route (hirschbergH1_wip + post-work)
    "prot_A2_udh": ("-Q0 -A2 -yX0 -V128K -TDictyost", 23),
    "prot_A2_udh_local": ("-Q0 -A2 -yX0 -V128K -LS -TDictyost", 24),
    "prot_A6_udh_recursive": ("-Q0 -A6 -yX0 -V96K -TDictyost", 25),
}


def annotation(rng, qlen, truth, step):
    """intron positions to annotate a query with (src/gsinfo.h:76-126): the true exon boundaries in
    query coordinates (coding positions for a protein: step 3), a few neighbours of them and some
    random positions, each with a multiplicity 1 .. 3.  Returns (positions, multiplicities), sorted."""
    pos = set()
    if truth is not None and len(truth) > 1:
        acc = 0
        for (s, e) in truth[:-1]:
            acc += e - s
            pos.add(acc)
            if rng.random() < 0.5:
                pos.add(acc + int(rng.integers(-2, 3)))
    for _ in range(int(rng.integers(1, 6))):
        pos.add(int(rng.integers(1, max(2, step * qlen))))
    pos = sorted(x for x in pos if 0 < x < step * qlen)
    return np.array(pos, np.int32), rng.integers(1, 4, size=len(pos)).astype(np.int32)


def gen_protein(name: str):
    import ref_harness as R
    from spaln_b200 import workload as synth

    opts, seed = CONFIGS[name]
    ref = R.Reference(opts, protein=True)
    p = ref.params()
    rng = np.random.default_rng(seed)
    out = {"opts": np.array(opts)}
    for k, v in p.items():
        out["prm_" + k] = np.asarray(v)
    n = 0
    with_cip = "cip" in name
    udh = "udh" in name or with_cip
    scalar_mode = "_A0_" in name
    ng_tables = None

    def add(g, q, tag="", truth=None, **setkw):
        nonlocal n
        t = ref.task(g, q)
        if setkw:
            t.set(**setkw)
        cip = None
        if with_cip:
            cip = t.set_cip(*annotation(rng, len(q), truth, 3))
        lw, up = t.stripe31(p["sh"])
        ex = t.export_p()
        if scalar_mode:
            # -A0 leaves the quantised intron penalty of the `_wip` kernels unset: scalar kernels only
            r = {"score": 0, "skl": np.zeros((0, 2), np.int32)}
            r1 = {"score": 0}
        else:
            r = t.kernel_p(lw, up, 0)
            r1 = t.kernel_p(lw, up, 1)
        pre = f"p{n}_"
        out[pre + "a"] = ex["a"]
        out[pre + "b"] = ex["b"]
        out[pre + "sgpt6"] = ex["sgpt6"]
        out[pre + "geom"] = np.array([ex["a_left"], ex["a_right"], ex["b_left"], ex["b_right"],
                                      ex["a_exgl"], ex["a_exgr"], ex["b_exgl"], ex["b_exgr"],
                                      lw, up, ex["blen"], ex["alen"]], np.int32)
        out[pre + "score"] = np.int32(r["score"])
        out[pre + "skl"] = r["skl"].astype(np.int32)
        out[pre + "score_only"] = np.int32(r1["score"])
        out[pre + "tag"] = np.array(tag)
        # scalar kernel (Aln2h1::trcbkalignH_ng, scalar branch) and its extra inputs
        rn = t.scalar_p(lw, up)
        out[pre + "int53"] = t.export_int53()
        out[pre + "ng_score"] = np.int32(rn["score"])
        out[pre + "ng_skl"] = rn["skl"].astype(np.int32)
        if cip is not None:
            out[pre + "cip"] = cip
        nonlocal ng_tables
        if ng_tables is None:
            ng_tables = t.export_ng_tables(1 << 17)
        if udh:
            # the whole driver (Aln2h1::lspH_ng) and the Hirschberg pass alone
            rl = t.lsp_p(lw, up)
            out[pre + "lsp_score"] = np.int32(rl["score"])
            out[pre + "lsp_skl"] = rl["skl"].astype(np.int32)
            m = ex["a_right"] - ex["a_left"]
            if scalar_mode and not tag.startswith("tiny"):
                for nn in (1, 2, 5):
                    if m < 4 * nn:
                        continue
                    intvl = (m + nn) // (nn + 1)
                    nq = nn - 1 if intvl * nn == m else nn
                    if nq < 1:
                        continue
                    rs = t.scalar_udh_p(lw, up, nq, intvl)
                    out[pre + f"sudh{nn}_nim"] = np.int32(nq)
                    out[pre + f"sudh{nn}_intvl"] = np.int32(intvl)
                    out[pre + f"sudh{nn}_score"] = np.int32(rs["score"])
                    out[pre + f"sudh{nn}_cpos"] = rs["cpos"].astype(np.int32)
                    out[pre + f"sudh{nn}_ranges"] = np.array(rs["ranges"], np.int32)
            if m >= 16 and not scalar_mode:
                n_im = max(1, min(3, m // 16))
                rh = t.udh_p(lw, up, n_im)
                out[pre + "udh_nim"] = np.int32(n_im)
                out[pre + "udh_score"] = np.int32(rh["score"])
                out[pre + "udh_cpos"] = rh["cpos"].astype(np.int32)
                out[pre + "udh_ranges"] = np.array(rh["ranges"], np.int32)
        n += 1
        t.close()

    for i in range(10):
        g, q, tr = synth.plant_protein_gene(rng, plen_range=(20, 200), flank=(40, 250))
        add(g, q, tag="gene", truth=tr)
    if with_cip:
        # a window of rows around an exon junction, all ends global: a post-work block
        for k, (lo, hi) in enumerate(((2, 3), (1, 2), (3, 3), (2, 5), (4, 2), (1, 1))):
            g, q, tr = synth.plant_protein_gene(rng, plen_range=(60, 120), n_exons=3, flank=(30, 90), sub=0)
            j = k % 2
            c = sum(e - s for s, e in tr[: j + 1])      # coding position of the junction
            m = c // 3
            add(g, q, tag=f"few{lo + hi}", truth=tr, a_left=m - lo, a_right=m + hi,
                b_left=tr[j][1] - (c - 3 * (m - lo)), b_right=tr[j + 1][0] + (3 * (m + hi) - c),
                a_exgl=0, a_exgr=0, b_exgl=0, b_exgr=0)
    g, q, _ = synth.plant_protein_gene(rng, plen_range=(80, 150), flank=(60, 200))
    add(g, q, tag="global_left", a_exgl=0, b_exgl=0)
    add(g, q, tag="global_right", a_exgr=0, b_exgr=0)
    add(g, q, tag="global_all", a_exgl=0, a_exgr=0, b_exgl=0, b_exgr=0)
    add(g, q, tag="subrange", a_left=7, a_right=len(q) - 5, b_left=31, b_right=len(g) - 43)
    for pl in (8, 15, 16, 17, 31, 32, 33):
        g2, q2, _ = synth.plant_protein_gene(rng, plen_range=(pl, pl), n_exons=1, flank=(30, 120))
        add(g2, q2, tag=f"tiny{pl}")
    g, q, _ = synth.plant_protein_gene(rng, plen_range=(640, 700), n_exons=3, flank=(40, 80))
    add(g, q, tag="rebase")     # > 544 rows: crosses the int16 re-basing check point
    out["n"] = np.int32(n)
    # inputs of the scalar kernel: Penalty() table, sig53tab, split-codon tables, minl
    out["prm_penalty"] = ng_tables["penalty"]
    out["prm_sig53tab"] = ng_tables["sig53tab"]
    for k, v in ref.scalar_p_tables().items():
        out["prm_" + k] = np.asarray(v)
    path = HERE / f"{name}.npz"
    np.savez_compressed(path, **out)
    print(name, "problems:", n, "->", path, f"{path.stat().st_size / 1024:.0f} KiB")



def gen(name: str):
    import ref_harness as R
    from spaln_b200 import workload as synth

    opts, seed = CONFIGS[name]
    ref = R.Reference(opts)
    p = ref.params()
    rng = np.random.default_rng(seed)
    out = {"opts": np.array(opts)}
    for k, v in p.items():
        out["prm_" + k] = np.asarray(v)
    probs = []
    with_cip = "cip" in name
    udh = "udh" in name or with_cip
    scalar_mode = "_A0_" in name
    ng_tables = None
    scan_f = None

    def add(g, q, comrev=False, tag="", truth=None, **setkw):
        t = ref.task(g, q, comrev)
        if setkw:
            t.set(**setkw)
        cip = None
        if with_cip:
            cip = t.set_cip(*annotation(rng, len(q), None if comrev else truth, 1))
        lw, up = t.stripe(p["sh"])
        ex = t.export()
        if scalar_mode:
            # -A0 leaves the quantised intron penalty of the `_wip` kernels unset: scalar kernels only
            r = {"score": 0, "skl": np.zeros((0, 2), np.int32)}
            r1 = {"score": 0}
        else:
            r = t.kernel(lw, up, 0, cap=1 << 16)
            r1 = t.kernel(lw, up, 1)
        i = len(probs)
        pre = f"p{i}_"
        out[pre + "a"] = ex["a"]
        out[pre + "b"] = ex["b"]
        out[pre + "sig5"] = ex["sig5"]
        out[pre + "sig3"] = ex["sig3"]
        out[pre + "geom"] = np.array([ex["a_left"], ex["a_right"], ex["b_left"], ex["b_right"],
                                      ex["a_exgl"], ex["a_exgr"], ex["b_exgl"], ex["b_exgr"],
                                      lw, up], np.int32)
        out[pre + "score"] = np.int32(r["score"])
        out[pre + "skl"] = r["skl"].astype(np.int32)
        out[pre + "score_only"] = np.int32(r1["score"])
        out[pre + "tag"] = np.array(tag)
        # scalar exact-ILD kernel (Aln2s1::trcbkalignS_ng, scalar branch) and its extra inputs
        rn = t.scalar(lw, up)
        out[pre + "int53"] = t.export_int53()
        out[pre + "ng_score"] = np.int32(rn["score"])
        out[pre + "ng_skl"] = rn["skl"].astype(np.int32)
        out[pre + "ng_score_only"] = np.int32(t.scorealone(lw, up))     # Aln2s1::scorealoneS_ng
        if cip is not None:
            out[pre + "cip"] = cip
        nonlocal ng_tables, scan_f
        scan_f = t.scan_factors()
        if ng_tables is None or len(ng_tables["penalty"]) < ex["blen"] + 2:
            ng_tables = t.export_ng_tables(max(1 << 19, ex["blen"] + 2))
        if udh:
            # the whole driver (Aln2s1::lspS_ng) and the Hirschberg pass alone
            rl = t.lsp(lw, up)
            out[pre + "lsp_score"] = np.int32(rl["score"])
            out[pre + "lsp_skl"] = rl["skl"].astype(np.int32)
            m = ex["a_right"] - ex["a_left"]
            width = up - lw + 3
            mode = 2 if (max(abs(lw), up) + width) < 32767 else 4
            n_im = max(1, min(3, m // 16))
            if scalar_mode and not tag.startswith("tiny"):
                # the scalar pass, with the spacing of the intermediate rows as lspS_ng sets it
                # (unrelated random pairs make the reference dereference a null intermediate,
                # src/fwd2s1.cc:1093: skipped)
                for nn in (1, 2, 5):
                    if m < 4 * nn:
                        continue
                    intvl = (m + nn) // (nn + 1)
                    nq = nn - 1 if intvl * nn == m else nn
                    if nq < 1:
                        continue
                    rs = t.scalar_udh(lw, up, nq, intvl)
                    out[pre + f"sudh{nn}_nim"] = np.int32(nq)
                    out[pre + f"sudh{nn}_intvl"] = np.int32(intvl)
                    out[pre + f"sudh{nn}_score"] = np.int32(rs["score"])
                    out[pre + f"sudh{nn}_cpos"] = rs["cpos"].astype(np.int32)
                    out[pre + f"sudh{nn}_ranges"] = np.array(rs["ranges"], np.int32)
            if m >= 16 and not scalar_mode:
                rh = t.kernel(lw, up, 2, n_imd=n_im, mode=mode)
                out[pre + "udh_nim"] = np.int32(n_im)
                out[pre + "udh_score"] = np.int32(rh["score"])
                out[pre + "udh_cpos"] = rh["cpos"].astype(np.int32)
                out[pre + "udh_ranges"] = np.array(rh["ranges"], np.int32)
        probs.append(i)
        t.close()

    # planted genes, both orientations of the query
    for i in range(10):
        g, q, tr = synth.plant_gene(rng, qlen_range=(60, 900) if udh else (60, 500),
                                    flank=(50, 900) if udh else (50, 300))
        add(g, q, tag="gene", truth=tr)
        if i % 3 == 0:
            add(g, q, comrev=True, tag="gene_rc")
    if with_cip:
        # few-row problems with introns: what the driver hands to the exact-ILD kernel
        # (a window of rows around an exon junction, all ends global: a post-work block)
        for k, (lo, hi) in enumerate(((3, 4), (2, 2), (1, 5), (4, 3), (3, 3), (6, 6), (2, 4), (5, 2))):
            g, q, tr = synth.plant_gene(rng, qlen_range=(120, 300), n_exons=3, flank=(40, 120), sub=0, indel=0)
            j = k % 2
            J = sum(e - s for s, e in tr[: j + 1])
            add(g, q, tag=f"few{lo + hi}", truth=tr, a_left=J - lo, a_right=J + hi,
                b_left=tr[j][1] - lo, b_right=tr[j + 1][0] + hi, a_exgl=0, a_exgr=0, b_exgl=0, b_exgr=0)
    # end-gap variants and sub-ranges (UDH post-work style: all four flags 0)
    g, q, _ = synth.plant_gene(rng, qlen_range=(150, 300), flank=(60, 200))
    add(g, q, tag="global_left", a_exgl=0, b_exgl=0)
    add(g, q, tag="global_right", a_exgr=0, b_exgr=0)
    add(g, q, tag="global_all", a_exgl=0, a_exgr=0, b_exgl=0, b_exgr=0)
    add(g, q, tag="subrange", a_left=17, a_right=len(q) - 9, b_left=33, b_right=len(g) - 41)
    add(g, q, tag="subrange_global", a_left=5, a_right=len(q) - 3, b_left=20, b_right=len(g) - 10,
        a_exgl=0, a_exgr=0, b_exgl=0, b_exgr=0)
    # ragged / tiny shapes: fewer rows than one strip, exact strip multiples
    for ql in (1, 2, 7, 15, 16, 17, 31, 32, 33, 48):
        g2, q2 = synth.random_pair(rng, ql, int(rng.integers(40, 300)))
        add(g2, q2, tag=f"tiny{ql}")
    # one query long enough to cross the int16 re-basing check point (> 1472 rows)
    g, q, _ = synth.plant_gene(rng, qlen_range=(1700, 1900), n_exons=3, flank=(40, 80))
    add(g, q, tag="rebase")
    out["n"] = np.int32(len(probs))
    # IntronPenalty::Penalty(n) table, Exinon::sig53tab, alprm2.Z > 0 (inputs of the scalar kernel)
    out["prm_penalty"] = ng_tables["penalty"]
    out["prm_sig53tab"] = ng_tables["sig53tab"]
    out["prm_intpot"] = np.int32(ng_tables["intpot"])
    # splice PSSMs + factors of the signal scan (Exinon::intron53_n): the fixtures' sig5 / sig3
    # arrays are the reference's output for the fixtures' genome codes
    for w, nm in ((0, "pat5"), (1, "pat3")):
        pm = ref.patmat(w)
        if pm is not None:
            out[f"prm_{nm}_meta"] = np.array([pm["rows"], pm["cols"], pm["offset"], pm["nalpha"], pm["morder"]], np.int32)
            out[f"prm_{nm}_f"] = np.array([pm["tonic"], pm["min_elem"]], np.float32)
            out[f"prm_{nm}_mtx"] = pm["mtx"]
    out["prm_scan_f"] = np.array([scan_f["fS"], scan_f["sss"]], np.float32)
    path = HERE / f"{name}.npz"
    np.savez_compressed(path, **out)
    print(name, "problems:", len(probs), "->", path, f"{path.stat().st_size / 1024:.0f} KiB")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        (gen_protein if sys.argv[1].startswith("prot") else gen)(sys.argv[1])
    else:
        for name in CONFIGS:
            subprocess.run([sys.executable, __file__, name], check=True)
