"""Seeded synthetic spliced-alignment problems (sequence level).

A *gene* is planted in a genomic segment: exons separated by GT...AG introns;
the query is the spliced transcript with substitutions / indels.  Used by the
parity tests (small sizes) and by bench.py (BASELINE.json configs).  Pure
numpy; no reference code involved.
"""
from __future__ import annotations

import numpy as np

ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_dna(rng, n, gc=0.41):
    pr = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
    return ALPHA[rng.choice(4, size=n, p=pr)]


def intron_lengths(rng, k, scale=1.0, lo=40, hi=500000):
    # Frechet-like heavy tail around ~100 nt (Dictyostelium-like when scale=1)
    u = rng.random(k)
    x = 60.0 * scale / np.power(-np.log(u), 1 / 2.5) + 30 * scale
    return np.clip(x.astype(np.int64), lo, hi)


def plant_gene(rng, qlen_range=(300, 900), n_exons=None, flank=(200, 800),
               intron_scale=1.0, sub=0.01, indel=0.002, gc=0.41):
    """returns (genome_segment: str, query: str, truth: list of exon (start,end)
    0-based half-open on the segment)."""
    qlen = int(rng.integers(qlen_range[0], qlen_range[1] + 1))
    if n_exons is None:
        n_exons = 1 + int(rng.poisson(3))
    # split qlen into exon lengths >= 20
    n_exons = max(1, min(n_exons, qlen // 25))
    cuts = np.sort(rng.choice(np.arange(1, qlen // 20), size=n_exons - 1, replace=False)) * 20 \
        if n_exons > 1 else np.array([], np.int64)
    bounds = np.concatenate([[0], cuts, [qlen]])
    exlens = np.diff(bounds)
    introns = intron_lengths(rng, n_exons - 1, intron_scale)
    fl = int(rng.integers(flank[0], flank[1] + 1))
    fr = int(rng.integers(flank[0], flank[1] + 1))
    parts = [random_dna(rng, fl, gc)]
    truth = []
    pos = fl
    mrna = []
    for i, el in enumerate(exlens):
        ex = random_dna(rng, int(el), gc)
        parts.append(ex)
        mrna.append(ex)
        truth.append((pos, pos + int(el)))
        pos += int(el)
        if i < n_exons - 1:
            il = int(introns[i])
            it = random_dna(rng, il, gc)
            it[:2] = np.frombuffer(b"GT", np.uint8)
            it[-2:] = np.frombuffer(b"AG", np.uint8)
            parts.append(it)
            pos += il
    parts.append(random_dna(rng, fr, gc))
    genome = np.concatenate(parts)
    q = np.concatenate(mrna)
    # mutate query
    q = q.copy()
    nsub = rng.binomial(len(q), sub)
    idx = rng.choice(len(q), size=nsub, replace=False)
    q[idx] = ALPHA[rng.integers(0, 4, size=nsub)]
    nind = rng.binomial(len(q), indel)
    for _ in range(nind):
        j = int(rng.integers(1, len(q) - 1))
        if rng.random() < 0.5:
            q = np.delete(q, j)
        else:
            q = np.insert(q, j, ALPHA[rng.integers(0, 4)])
    return genome.tobytes().decode(), q.tobytes().decode(), truth


def random_pair(rng, qlen, glen, gc=0.41):
    return (random_dna(rng, glen, gc).tobytes().decode(),
            random_dna(rng, qlen, gc).tobytes().decode())


# ---------------------------------------------------------------------------
# code-level helpers (no reference needed)
# ---------------------------------------------------------------------------
# residue codes of the reference's DNA alphabet (src/seq.h: A=2, C=3, G=5, T=9)
DNA_CODE = np.zeros(256, np.uint8)
for _ch, _c in (("A", 2), ("C", 3), ("G", 5), ("T", 9), ("N", 16)):
    DNA_CODE[ord(_ch)] = _c
    DNA_CODE[ord(_ch.lower())] = _c


def encode_dna(s: str) -> np.ndarray:
    return DNA_CODE[np.frombuffer(s.encode(), np.uint8)]


def synthetic_signals(bcodes: np.ndarray, rng, scale: float = 1.0):
    """Splice-signal tables shaped like Exinon::data_n (src/codepot.cc:479-523)
    without the PSSM scan: per column n a 5' score driven by the dinucleotide
    at(n), at(n+1) and a 3' score driven by at(n-2), at(n-1), plus noise.
    The magnitudes follow the reference's tables for Dictyostelium (x10 scale):
    GT ~ +40, GC/AT ~ -78, others ~ -250;  AG ~ +27, AC ~ -95, others ~ -260.
    Returns (sig5, sig3) int16 arrays indexed by column n in [0, len + 1]."""
    L = len(bcodes)
    c = np.concatenate([[0], bcodes.astype(np.int64), [0, 0]])     # c[i + 1] = at(i)
    n = np.arange(L + 2)
    x, y = c[np.minimum(n + 1, L + 2)], c[np.minimum(n + 2, L + 2)]   # at(n), at(n+1)
    s5 = np.full(L + 2, -250.0)
    s5[(x == 5) & (y == 9)] = 40.0
    s5[(x == 5) & (y == 3)] = -77.0
    s5[(x == 2) & (y == 9)] = -78.0
    u, v = c[np.maximum(n - 1, 0)], c[n]                            # at(n-2), at(n-1)
    s3 = np.full(L + 2, -260.0)
    s3[(u == 2) & (v == 5)] = 27.0
    s3[(u == 2) & (v == 3)] = -95.0
    s5 = s5 * scale + rng.normal(0, 20.0, L + 2)
    s3 = s3 * scale + rng.normal(0, 20.0, L + 2)
    s5[0] = s3[0] = 0
    s5[L + 1] = s3[L + 1] = 0
    return s5.astype(np.int16), s3.astype(np.int16)


def stripe(a_left, a_right, b_left, b_right, sh=100):
    """band window exactly as `stripe()` (src/aln2.cc:156-176), cmode 0"""
    up = b_right - a_right
    lw = b_left - a_left
    if up < lw:
        up, lw = lw, up
    up += sh
    lw -= sh
    up = min(up, b_right - a_left)
    lw = max(lw, b_left - a_right)
    return lw, up


# ---------------------------------------------------------------------------
# BASELINE.json config 2: synthetic cDNA (1-3 kb) against its genomic locus
# ---------------------------------------------------------------------------
def config2_pair(rng, qlen_range=(1000, 3000), flank=(500, 5000), intron_scale=1.0,
                 sub=0.01, indel=0.002, gc=0.41):
    """One (genome segment, cDNA) pair of the config-2 shape:
    exon lengths ~ LogNormal(ln 150, 0.6) clipped to [30, 2000] drawn until the
    transcript reaches a target length U[1000, 3000]; GT..AG introns from a
    heavy-tailed length model; locus +- U[500, 5000] nt flanks; 1 % substitutions
    and 0.2 % indels in the query.  Returns code arrays (uint8) and exon truth."""
    target = int(rng.integers(qlen_range[0], qlen_range[1] + 1))
    exl = []
    tot = 0
    while tot < target:
        e = int(np.clip(rng.lognormal(np.log(150.0), 0.6), 30, 2000))
        e = min(e, target - tot) if target - tot >= 30 else e
        exl.append(e)
        tot += e
    introns = intron_lengths(rng, len(exl) - 1, intron_scale)
    fl = int(rng.integers(flank[0], flank[1] + 1))
    fr = int(rng.integers(flank[0], flank[1] + 1))
    glen = fl + fr + tot + int(introns.sum())
    genome = random_dna(rng, glen, gc)
    pos = fl
    mrna = []
    truth = []
    for i, e in enumerate(exl):
        mrna.append(genome[pos:pos + e])
        truth.append((pos, pos + e))
        pos += e
        if i < len(exl) - 1:
            il = int(introns[i])
            genome[pos:pos + 2] = np.frombuffer(b"GT", np.uint8)
            genome[pos + il - 2:pos + il] = np.frombuffer(b"AG", np.uint8)
            pos += il
    q = np.concatenate(mrna).copy()
    nsub = rng.binomial(len(q), sub)
    if nsub:
        idx = rng.choice(len(q), size=nsub, replace=False)
        q[idx] = ALPHA[rng.integers(0, 4, size=nsub)]
    nind = rng.binomial(len(q), indel)
    if nind:
        where = np.sort(rng.choice(np.arange(1, len(q) - 1), size=nind, replace=False))
        keep = np.ones(len(q), bool)
        dele = rng.random(nind) < 0.5
        keep[where[dele]] = False
        ins_at = where[~dele]
        q = np.insert(q, ins_at, ALPHA[rng.integers(0, 4, size=len(ins_at))]) if len(ins_at) else q
        if dele.any():
            # positions shift after insertion; recompute a deletion mask on the new array
            shift = np.searchsorted(ins_at, where[dele])
            keep2 = np.ones(len(q), bool)
            keep2[where[dele] + shift] = False
            q = q[keep2]
    return genome, q, truth


def config2_problem(rng, sh=100, **kw):
    """config-2 pair -> raw DP inputs (codes, synthetic splice-signal tables, band)."""
    g, q, truth = config2_pair(rng, **kw)
    a = DNA_CODE[q]
    b = DNA_CODE[g]
    s5, s3 = synthetic_signals(b, rng)
    lw, up = stripe(0, len(a), 0, len(b), sh)
    return {"a": a, "b": b, "sig5": s5, "sig3": s3, "a_left": 0, "a_right": len(a),
            "b_left": 0, "b_right": len(b), "a_exgl": 1, "a_exgr": 1, "b_exgl": 1, "b_exgr": 1,
            "lw": lw, "up": up, "truth": truth,
            "genome_str": g.tobytes().decode(), "query_str": q.tobytes().decode()}


# ---------------------------------------------------------------------------
# protein x genome (BASELINE.json config 3 shape)
# ---------------------------------------------------------------------------
_CODONS = {}
_BASES = "TCAG"
_AAS = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
for _i, _aa in enumerate(_AAS):
    _CODONS.setdefault(_aa, []).append(_BASES[_i // 16] + _BASES[(_i // 4) % 4] + _BASES[_i % 4])
_AA_LETTERS = "ACDEFGHIKLMNPQRSTVWY"


def plant_protein_gene(rng, plen_range=(100, 400), n_exons=None, flank=(100, 600),
                       intron_scale=1.0, sub=0.05, gc=0.41):
    """returns (genome segment, protein query, CDS exon truth).  The CDS (start codon ..
    stop codon) is cut into exons at arbitrary codon phases; introns are GT..AG."""
    plen = int(rng.integers(plen_range[0], plen_range[1] + 1))
    prot = ["M"] + [_AA_LETTERS[i] for i in rng.integers(0, 20, size=plen - 1)]
    cds = "".join(_CODONS[a][int(rng.integers(0, len(_CODONS[a])))] for a in prot) + "TAA"
    L = len(cds)
    if n_exons is None:
        n_exons = 1 + int(rng.poisson(2))
    n_exons = max(1, min(n_exons, L // 40))
    cuts = np.sort(rng.choice(np.arange(20, L - 20), size=n_exons - 1, replace=False)) if n_exons > 1 else []
    bounds = [0] + [int(c) for c in cuts] + [L]
    introns = intron_lengths(rng, n_exons - 1, intron_scale)
    fl = int(rng.integers(flank[0], flank[1] + 1))
    fr = int(rng.integers(flank[0], flank[1] + 1))
    parts = [random_dna(rng, fl, gc).tobytes().decode()]
    truth = []
    pos = fl
    for i in range(n_exons):
        ex = cds[bounds[i]:bounds[i + 1]]
        parts.append(ex)
        truth.append((pos, pos + len(ex)))
        pos += len(ex)
        if i < n_exons - 1:
            il = int(introns[i])
            it = random_dna(rng, il, gc)
            it[:2] = np.frombuffer(b"GT", np.uint8)
            it[-2:] = np.frombuffer(b"AG", np.uint8)
            parts.append(it.tobytes().decode())
            pos += il
    parts.append(random_dna(rng, fr, gc).tobytes().decode())
    q = list(prot)
    nsub = rng.binomial(len(q), sub)
    for j in rng.choice(len(q), size=nsub, replace=False):
        q[int(j)] = _AA_LETTERS[int(rng.integers(0, 20))]
    return "".join(parts), "".join(q), truth
