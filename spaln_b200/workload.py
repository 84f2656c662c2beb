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


def synthetic_int53(bcodes: np.ndarray) -> np.ndarray:
    """Exinon::int53 as Exinon::intron53_c fills it (src/codepot.cc:437-477) for algmode.any == 0
    and one orientation: per column n the rolling dinucleotide codes dinc5 = (at(n), at(n+1)),
    dinc3 = (at(n-2), at(n-1)) (A, C, G, T = 0..3, anything else counts as C, the residue before
    the sequence too) and the site classes cano5 (GT, GC: 3; AT: 2) and cano3 (AG: 3; AC: 2).
    Returns uint16 dinc5 | dinc3 << 4 | cano5 << 8 | cano3 << 12 for n in [0, len + 1]
    (entries the reference leaves unwritten are 0)."""
    L = len(bcodes)
    red = np.full(256, 1, np.int64)
    for code, v in ((2, 0), (3, 1), (5, 2), (9, 3)):
        red[code] = v
    c = red[np.asarray(bcodes, np.uint8)]
    prev = np.concatenate([[1], c[:-1]]) if L else c
    dinc = ((prev << 2) | c) & 15                   # dinc[i]: residues at(i - 1), at(i)
    out = np.zeros(L + 2, np.int64)
    c5 = np.zeros(16, np.int64); c5[[11, 9]] = 3; c5[3] = 2     # GT, GC, AT
    c3 = np.zeros(16, np.int64); c3[2] = 3; c3[1] = 2           # AG, AC
    n5 = np.arange(0, L - 1)                        # dinc5[n] = dinc[n + 1]
    out[n5] |= dinc[n5 + 1] | (c5[dinc[n5 + 1]] << 8)
    n3 = np.arange(1, L + 1)                        # dinc3[n] = dinc[n - 1]
    out[n3] |= (dinc[n3 - 1] << 4) | (c3[dinc[n3 - 1]] << 12)
    return out.astype(np.uint16)


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


def config2_problem_seeded(seed: int, i: int, **kw):
    """problem i of the seeded global config-2 query set: every problem has its own generator
    state, so any rank (or worker process) can build any problem on its own"""
    return config2_problem(np.random.default_rng([int(seed), int(i)]), **kw)


def config2_chunk(args):
    """worker of the global generator: problems [lo, hi) as flat arrays
    (query codes, genome codes, sig5, sig3, int53, lengths)"""
    seed, lo, hi, kw = args
    A, B, S5, S3, I53, lens = [], [], [], [], [], []
    for i in range(lo, hi):
        r = config2_problem_seeded(seed, i, **kw)
        A.append(r["a"]); B.append(r["b"]); S5.append(r["sig5"]); S3.append(r["sig3"])
        I53.append(synthetic_int53(r["b"]))
        lens.append((len(r["a"]), len(r["b"])))
    return (np.concatenate(A), np.concatenate(B), np.concatenate(S5), np.concatenate(S3),
            np.concatenate(I53), np.array(lens, np.int64))


def config2_global(n: int, seed: int, procs: int = 1, **kw):
    """The whole query set of a job as flat buffers: a dict with `a` (query codes), `b` (genome
    codes: the concatenated loci, i.e. the formatted genome of the job), `sig5`, `sig3`, `int53`
    (per locus len + 2 entries) and `lens` ((n, 2): query, locus length).  Problem i is
    `config2_problem_seeded(seed, i)`.  procs > 1 forks worker processes (call before CUDA is
    initialised)."""
    step = 256
    jobs = [(seed, lo, min(n, lo + step), kw) for lo in range(0, n, step)]
    if procs > 1 and len(jobs) > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(procs, len(jobs))) as pool:
            parts = pool.map(config2_chunk, jobs)
    else:
        parts = [config2_chunk(j) for j in jobs]
    keys = ("a", "b", "sig5", "sig3", "int53", "lens")
    return {k: np.concatenate([p[j] for p in parts]) for j, k in enumerate(keys)}


def global_problems(g: dict, index, sh=100):
    """raw problem dicts (views into the flat buffers) of the problems `index` of a global set"""
    lens = g["lens"]
    a_off = np.concatenate([[0], np.cumsum(lens[:, 0])])
    b_off = np.concatenate([[0], np.cumsum(lens[:, 1])])
    t_off = np.concatenate([[0], np.cumsum(lens[:, 1] + 2)])
    out = []
    for i in index:
        i = int(i)
        la, lb = int(lens[i, 0]), int(lens[i, 1])
        lw, up = stripe(0, la, 0, lb, sh)
        out.append({"a": g["a"][a_off[i]:a_off[i] + la], "b": g["b"][b_off[i]:b_off[i] + lb],
                    "sig5": g["sig5"][t_off[i]:t_off[i] + lb + 2], "sig3": g["sig3"][t_off[i]:t_off[i] + lb + 2],
                    "int53": g["int53"][t_off[i]:t_off[i] + lb + 2],
                    "a_left": 0, "a_right": la, "b_left": 0, "b_right": lb, "lw": lw, "up": up, "index": i})
    return out


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


# residue codes of the reference (measured through oracle/_ref, see tests/tools): amino acids
# (Seq code set of a protein) and "tron" codes = translated codon centred on a nucleotide
# (Seq::nuc2tron, src/seq.cc:774-798); 23 = Ser (AGY), 24 = TGA, 25 = TAA / TAG
AA_CODE = {'A': 3, 'C': 7, 'D': 6, 'E': 9, 'F': 16, 'G': 10, 'H': 11, 'I': 12, 'K': 14, 'L': 13,
           'M': 15, 'N': 5, 'P': 17, 'Q': 8, 'R': 4, 'S': 18, 'T': 19, 'V': 22, 'W': 20, 'Y': 21}
_TRON_AMB = 2
_NT = {"A": 0, "C": 1, "G": 2, "T": 3}
_TRON_TAB = np.zeros(64, np.uint8)
for _i, _aa in enumerate(_AAS):
    _cod = _BASES[_i // 16] + _BASES[(_i // 4) % 4] + _BASES[_i % 4]
    _ix = _NT[_cod[0]] * 16 + _NT[_cod[1]] * 4 + _NT[_cod[2]]
    if _aa == "*":
        _TRON_TAB[_ix] = 24 if _cod == "TGA" else 25
    elif _cod in ("AGC", "AGT"):
        _TRON_TAB[_ix] = 23
    else:
        _TRON_TAB[_ix] = AA_CODE[_aa]


def encode_protein(q: str) -> np.ndarray:
    return np.array([AA_CODE[c] for c in q], np.uint8)


def nuc2tron(g: str) -> np.ndarray:
    """tron code of at(i) = translation of the codon (i-1, i, i+1).  The ends follow
    nuc2tron3 (src/utilseq.cc:204-224) as measured through oracle/_ref: at(0) has no first
    nucleotide (most abundant residue for its middle one), at(len - 1) reads a NUL third
    nucleotide (== A)."""
    v = np.array([_NT[c] for c in g], np.int64)
    out = np.full(len(g), _TRON_AMB, np.uint8)
    if len(g) >= 3:
        out[1:-1] = _TRON_TAB[v[:-2] * 16 + v[1:-1] * 4 + v[2:]]
        out[0] = (14, 3, 10, 13)[v[0]]
        out[-1] = _TRON_TAB[v[-2] * 16 + v[-1] * 4 + 0]
    return out


def stripe31(a_left, a_right, b_left, b_right, sh=100):
    """stripe31() of the reference (src/aln2.cc:178-199), cmode 0"""
    if sh < 0:
        sh = -sh * min(a_right - a_left, b_right - b_left) // 100
    sh *= 3
    up = b_right - 3 * a_right
    lw = b_left - 3 * a_left
    if up < lw:
        up, lw = lw, up
    up += sh
    lw -= sh
    up = min(up, b_right - 3 * a_left)
    lw = max(lw, b_left - 3 * a_right)
    return int(lw), int(up)


def synthetic_sgpt6(g: str, tron: np.ndarray, rng) -> np.ndarray:
    """(len + 2, 8) int16 table in the layout of Exinon::data_p (sig5, sig3, sigS, sigT, sigE,
    sigI, phs5, phs3), dinucleotide / codon driven with the magnitudes of the reference's
    Dictyostelium tables: GT / AG sites score around +60, background around -520; phs5 / phs3
    mark the three columns around a site (phase -1 / 0 / +1, 2 = both -1 and +1)."""
    n = len(g)
    gb = np.frombuffer(g.encode(), np.uint8)
    t = np.zeros((n + 2, 8), np.int16)
    t[:, 0] = rng.normal(-520, 60, n + 2).astype(np.int16)
    t[:, 1] = rng.normal(-520, 60, n + 2).astype(np.int16)
    t[:, 2] = rng.normal(-650, 90, n + 2).astype(np.int16)
    t[:, 3] = -1360
    t[:, 4] = rng.normal(0, 8, n + 2).astype(np.int16)
    t[:, 6] = -2
    t[:, 7] = -2
    is_gt = np.zeros(n + 2, bool)
    is_ag = np.zeros(n + 2, bool)
    # donor boundary at column i: intron starts with nt at(i), at(i + 1) == "GT"
    is_gt[:n - 1] = (gb[:-1] == ord("G")) & (gb[1:] == ord("T"))
    # acceptor boundary at column i: intron ends with nt at(i - 2), at(i - 1) == "AG"
    is_ag[2:n + 1] = (gb[:-1] == ord("A")) & (gb[1:] == ord("G"))
    t[is_gt, 0] = rng.normal(60, 25, int(is_gt.sum())).astype(np.int16)
    t[is_ag, 1] = rng.normal(50, 25, int(is_ag.sum())).astype(np.int16)
    for col, site in ((6, is_gt), (7, is_ag)):
        idx = np.nonzero(site)[0]
        ph = np.full(n + 2, -2, np.int16)
        for d in (-1, 0, 1):            # column c = i + d sees the site at phase d
            c = idx + d
            c = c[(c >= 0) & (c <= n + 1)]
            both = ph[c] != -2
            ph[c[~both]] = d
            # a column already marked (phase -1 then +1 from two sites) becomes "both"
            ph[c[both]] = np.where((ph[c[both]] == -1) & (d == 1), 2, ph[c[both]])
        t[:, col] = ph
    # start / stop signals on the codon that ends at column c (tron code of at(c - 2))
    tr = np.concatenate([[_TRON_AMB, _TRON_AMB], tron, [_TRON_AMB]])[:n + 2]    # tr[c] = tron(at(c - 2))
    t[tr == AA_CODE["M"], 2] = rng.normal(200, 40, int((tr == AA_CODE["M"]).sum())).astype(np.int16)
    stop = tr >= 24
    t[stop, 3] = rng.normal(320, 20, int(stop.sum())).astype(np.int16)
    return t


def protein_problem(rng, plen_range=(100, 400), flank=(100, 600), sh=100, intron_scale=1.0,
                    n_exons=None):
    """one protein x genomic-segment DP problem in the layout of tests/golden (arrays start at
    at(-1)) with a synthetic SGPT6 table"""
    g, q, _ = plant_protein_gene(rng, plen_range=plen_range, flank=flank, intron_scale=intron_scale,
                                 n_exons=n_exons)
    a = encode_protein(q)
    b = nuc2tron(g)
    sg = synthetic_sgpt6(g, b, rng)
    lw, up = stripe31(0, len(a), 0, len(b), sh)
    return {"a": np.concatenate([[0], a, [0]]).astype(np.uint8),
            "b": np.concatenate([[0], b, [0]]).astype(np.uint8),
            "sgpt6": sg, "blen": len(b), "a_left": 0, "a_right": len(a), "b_left": 0,
            "b_right": len(b), "a_exgl": 1, "a_exgr": 1, "b_exgl": 1, "b_exgr": 1,
            "lw": lw, "up": up, "genome": g, "query": q}
