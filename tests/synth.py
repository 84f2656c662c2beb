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
