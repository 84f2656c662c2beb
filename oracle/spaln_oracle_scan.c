/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the splice-signal scan that fills
 * the Exinon tables of a genomic DNA segment:
 *   src/codepot.cc:437-477   Exinon::intron53_c   (INT53: dinucleotide codes, site classes)
 *   src/codepot.cc:479-523   Exinon::intron53_n   (SGPT2: sig5 / sig3 from the two PSSMs)
 *   src/utilseq.cc:905-1002  PatMat::calcPatMat, Markov order <= 2, Seq::many == 1
 * Float arithmetic is the reference's: sequential fp32 adds of table entries, one fp32 multiply,
 * truncation to short.  Pinned against the SGPT2 / INT53 arrays of the unmodified reference
 * (tests/golden/, tests/test_oracle_scan.py).
 */
#include <stdlib.h>
#include <string.h>
#include "spaln_oracle.h"

/* ncredctab, src/seq.cc:31 (A, C, G, T -> 0..3; everything else >= 4) */
static const unsigned char so_ncred[17] = { 15, 15, 0, 1, 4, 2, 5, 6, 10, 3, 7, 8, 10, 9, 12, 13, 14 };

static int red(unsigned c) { return c < 17 ? so_ncred[c] : 15; }

/* one PSSM value: PatMat::calcPatMat for the window that starts at position n
 * (codes[i] == *sd->at(i), i in [0, len)) */
static float patmat_at(const so_patmat* pm, const uint8_t* codes, int len, int n)
{
    const int rows = pm->rows, cols = pm->cols, na = pm->nalpha, order = pm->morder;
    int s = n, e = n + cols;
    if (e > len - order) e = len - order;           /* tt = min(at(n + cols), zz) */
    const float* ptn = pm->mtx;
    if (n < 0) { ptn -= (long) n * rows; s = 0; }
    int q = n + cols >= len;                        /* bad characters */
    float fit = 0;
    if (order <= 1) {
        for (int m = 0; s < e; ptn += rows, ++m, ++s) {
            int k = red(codes[s]);
            if (k < 0 || k >= na) ++q;
            if (order && !q) {
                if (m == 0) fit += ptn[k];
                int j = red(codes[s + 1]);
                if (j < 0 || j >= na) ++q;
                k = na * k + j + na;
            }
            fit += q ? 0.f : ptn[k];
        }
        return fit + pm->tonic;
    }
    for (int m = 0; s < e; ptn += rows, ++m, ++s) {
        int i = red(codes[s]);
        int k = i;
        if (i > 3) ++q;
        if (m == 0 && q == 0) fit += ptn[k];
        i = red(codes[s + 1]);
        if (i > 3) ++q;
        else if (q == 0) { k = na * k + i; if (m == 0) fit += ptn[k + na]; }
        i = red(codes[s + 2]);
        if (i > 3) ++q;
        else if (q == 0) { k = na * k + i; fit += ptn[k + 20]; }
    }
    if (q) fit = (float) cols * pm->min_elem;
    return fit + pm->tonic;
}

void so_exinon_scan_n(const so_scan_params* sp, const uint8_t* codes, int len,
                      int16_t* sig5, int16_t* sig3, uint16_t* int53)
{
    static const unsigned jlevelac[4] = { 0, 2, 3, 1 }, jlevelgt[4] = { 0, 0, 3, 1 };
    const unsigned any = (unsigned) sp->any & 3;
    memset(sig5, 0, sizeof(int16_t) * (size_t) (len + 2));
    memset(sig3, 0, sizeof(int16_t) * (size_t) (len + 2));
    memset(int53, 0, sizeof(uint16_t) * (size_t) (len + 2));
    /* intron53_c: residue i closes the dinucleotide (i - 1, i); it is the 5' code of column
     * i - 1 and the 3' code of column i + 1 */
    unsigned nc = 1;
    for (int i = 0; i < len; ++i) {
        unsigned c = (unsigned) red(codes[i]);
        if (c >= 4) c = 1;
        nc = ((nc << 2) + c) & 15;
        unsigned c5 = any == 3, c3 = any == 3;
        switch (nc) {
          case 0: c3 = jlevelac[any]; break;                    /* AA */
          case 1: c3 = 2; break;                                /* AC */
          case 2: c3 = 3; break;                                /* AG */
          case 3: c5 = 2; c3 = jlevelac[any]; break;            /* AT */
          case 6: c3 = jlevelgt[any]; break;                    /* CG */
          case 7: c5 = jlevelgt[any]; break;                    /* CT */
          case 8: c5 = jlevelgt[any]; break;                    /* GA */
          case 9: c5 = 3; break;                                /* GC */
          case 10: c5 = jlevelgt[any]; c3 = jlevelgt[any]; break;   /* GG */
          case 11: c5 = 3; break;                               /* GT */
          case 14: c3 = jlevelgt[any]; break;                   /* TG */
          case 15: c5 = jlevelgt[any]; break;                   /* TT */
          default: break;
        }
        if (i - 1 >= 0) int53[i - 1] |= (uint16_t) (nc | (c5 << 8));
        int53[i + 1] |= (uint16_t) ((nc << 4) | (c3 << 12));
    }
    /* intron53_n: columns 0 .. len - 1 */
    const float fs = sp->fS * sp->sss;
    for (int n = 0; n < len; ++n) {
        int16_t s5 = sp->pat5.mtx ? (int16_t) (fs * patmat_at(&sp->pat5, codes, len, n - sp->pat5.offset)) : 0;
        int16_t s3 = sp->pat3.mtx ? (int16_t) (fs * patmat_at(&sp->pat3, codes, len, n - sp->pat3.offset)) : 0;
        s5 = (int16_t) (s5 + sp->sig53tab[int53[n] & 15]);
        s3 = (int16_t) (s3 + sp->sig53tab[16 + ((int53[n] >> 4) & 15)]);
        sig5[n] = s5;
        sig3[n] = s3;
    }
}

/* Seq::nuc2tron (src/seq.cc:774-798) with nuc2tron3 (src/utilseq.cc:205-224), Seq::many == 1:
 * the "tron" code of position i is the translation of the codon (i - 1, i, i + 1).  codes points
 * at at(0); codes[-1] and codes[len] are the terminal residues of the Seq (NUL).  gencode: the 64
 * codon table in use (src/utilseq.cc:38). */
void so_nuc2tron(const uint8_t* gencode, const uint8_t* codes, int len, uint8_t* tron)
{
    /* ncelements, src/seq.cc:33; most_abund, src/utilseq.cc:176; residue codes, src/cmn.h:115-117 */
    static const unsigned char ncel[17] = { 0, 0, 0, 1, 2, 2, 0, 2, 0, 3, 3, 3, 1, 1, 2, 3, 0 };
    static const unsigned char abund[4] = { 14, 3, 10, 13 };    /* LYS, ALA, GLY, LEU */
    enum { UNP = 1, AMB = 2, SER = 18, SER2 = 23, TRM2 = 24, TRM = 25, G = 5 };
    for (int i = 0; i < len; ++i) {
        const unsigned m = codes[i];
        int aa;
        if (m <= UNP) aa = UNP;                                 /* IsGap: middle residue deleted */
        else {
            const int c2 = red(m);
            if (c2 >= 4) aa = AMB;
            else {
                const int c1 = red(codes[i - 1]);
                const unsigned t = codes[i + 1];
                aa = c1 >= 4 ? abund[c2] : gencode[16 * c1 + 4 * c2 + ncel[t < 17 ? t : 0]];
                if (aa == SER && m == G) aa = SER2;
                else if (aa == TRM && m == G) aa = TRM2;
            }
        }
        tron[i] = (uint8_t) aa;
    }
}

/* ---- protein-side scan: Exinon::intron53_p (src/codepot.cc:525-619) over a TRON segment ----
 * A tron code keeps the middle nucleotide of its codon (tnredctab, src/seq.cc:41-42), so the
 * PSSMs and the coding potential read the same nucleotides as on the DNA side. */
static const unsigned char so_tnred[26] =
    { 4, 4, 4, 1, 2, 0, 0, 2, 0, 0, 2, 0, 3, 3, 0, 3, 3, 1, 1, 1, 2, 0, 3, 2, 2, 0 };

static int tred(unsigned c) { return c < 26 ? so_tnred[c] : 4; }

/* PatMat::calcPatMat on nucleotides nt[] (0..3, >= 4 bad), same code as patmat_at */
static float patmat_nt(const so_patmat* pm, const unsigned char* nt, int len, int n)
{
    const int rows = pm->rows, cols = pm->cols, na = pm->nalpha, order = pm->morder;
    int s = n, e = n + cols;
    if (e > len - order) e = len - order;
    const float* ptn = pm->mtx;
    if (n < 0) { ptn -= (long) n * rows; s = 0; }
    int q = n + cols >= len;
    float fit = 0;
    if (order <= 1) {
        for (int m = 0; s < e; ptn += rows, ++m, ++s) {
            int k = nt[s];
            if (k < 0 || k >= na) ++q;
            if (order && !q) {
                if (m == 0) fit += ptn[k];
                int j = nt[s + 1];
                if (j < 0 || j >= na) ++q;
                k = na * k + j + na;
            }
            fit += q ? 0.f : ptn[k];
        }
        return fit + pm->tonic;
    }
    for (int m = 0; s < e; ptn += rows, ++m, ++s) {
        int i = nt[s];
        int k = i;
        if (i > 3) ++q;
        if (m == 0 && q == 0) fit += ptn[k];
        i = nt[s + 1];
        if (i > 3) ++q;
        else if (q == 0) { k = na * k + i; if (m == 0) fit += ptn[k + na]; }
        i = nt[s + 2];
        if (i > 3) ++q;
        else if (q == 0) { k = na * k + i; fit += ptn[k + 20]; }
    }
    if (q) fit = (float) cols * pm->min_elem;
    return fit + pm->tonic;
}

/* out: 8 shorts per column n in [0, len + 1]: sig5, sig3, sigS, sigT, sigE, sigI, phs5, phs3
 * (the layout of tests/ref_harness.py::export_p); int53 as in so_exinon_scan_n */
void so_exinon_scan_p(const so_scan_params_p* sp, const uint8_t* tron, int len, int16_t* out, uint16_t* int53)
{
    enum { TRM2 = 24, TRM = 25 };
    static const unsigned jlevelac[4] = { 0, 2, 3, 1 }, jlevelgt[4] = { 0, 0, 3, 1 };
    const unsigned any = (unsigned) sp->base.any & 3;
    unsigned char* nt = (unsigned char*) malloc((size_t) len + 8);
    for (int i = 0; i < len; ++i) nt[i] = (unsigned char) tred(tron[i]);
    for (int i = len; i < len + 8; ++i) nt[i] = 4;
    memset(int53, 0, sizeof(uint16_t) * (size_t) (len + 2));
    for (int n = 0; n <= len + 1; ++n) {
        int16_t* o = out + 8 * n;
        o[0] = o[1] = o[2] = o[3] = o[4] = o[5] = 0;
        o[6] = o[7] = -2;                                   /* ZeroSGPT6 */
    }
    unsigned nc = 1;
    for (int i = 0; i < len; ++i) {                         /* intron53_c */
        unsigned c = nt[i];
        if (c >= 4) c = 1;
        nc = ((nc << 2) + c) & 15;
        unsigned c5 = any == 3, c3 = any == 3;
        switch (nc) {
          case 0: c3 = jlevelac[any]; break;
          case 1: c3 = 2; break;
          case 2: c3 = 3; break;
          case 3: c5 = 2; c3 = jlevelac[any]; break;
          case 6: c3 = jlevelgt[any]; break;
          case 7: c5 = jlevelgt[any]; break;
          case 8: c5 = jlevelgt[any]; break;
          case 9: c5 = 3; break;
          case 10: c5 = jlevelgt[any]; c3 = jlevelgt[any]; break;
          case 11: c5 = 3; break;
          case 14: c3 = jlevelgt[any]; break;
          case 15: c5 = jlevelgt[any]; break;
          default: break;
        }
        if (i - 1 >= 0) int53[i - 1] |= (uint16_t) (nc | (c5 << 8));
        int53[i + 1] |= (uint16_t) ((nc << 4) | (c3 << 12));
    }
    /* ExinPot::calcScr_3 (src/utilseq.cc:1423-1459): value of the character at(t), t = 0 .. len - 1
     * (the scan starts at at(-1), a terminal residue that resets the state) */
    float* pot = (float*) calloc((size_t) len + 8, sizeof(float));
    if (sp->codepot) {
        const int nd = sp->ndata, kk = sp->cp_order + 1;
        int x = kk, w = 0, buf[3] = { 0, 0, 0 };
        static const int nextp[3] = { 1, 2, 0 }, prevp[3] = { 2, 0, 1 };
        int p = nextp[1];                                   /* at(-1) was processed with p = 1 */
        for (int t = 0; t < len; ++t, p = nextp[p]) {
            const int c = nt[t];
            if (c < 4) { buf[p] = 3 * (w = (4 * w + c) % nd); if (x) --x; }
            else { w = 0; x = kk; }
            float val = 0;
            if (!x) {
                val += sp->codepot[buf[nextp[p]] + 2];
                val += sp->codepot[buf[prevp[p]]];
                val += sp->codepot[buf[p] + 1];
            }
            pot[t] = val;
        }
    }
    const float fs = sp->base.fS * sp->base.sss;
    const float fE = sp->z * sp->fact, fT = sp->bti * sp->fact, fO = -sp->o * sp->fact;
    for (int n = 0; n < len; ++n) {
        int16_t* o = out + 8 * n;
        if (sp->patI.mtx) o[2] = (int16_t) (fT * patmat_nt(&sp->patI, nt, len, n - sp->patI.offset));
        if (sp->patT.mtx) o[3] = (int16_t) (fT * patmat_nt(&sp->patT, nt, len, n - sp->patT.offset));
        if (sp->codepot) {
            /* prefE is computed over [left - 1, right + 1) and read from its second entry on, and
             * calcScr_3 delays its output by five characters: column n sees the character n + 5 */
            float sigE = fE * (n + 5 < len ? pot[n + 5] : 0.f);
            if (tron[n] == TRM || tron[n] == TRM2) sigE += fO;
            else if (n + 3 < len && (tron[n + 3] == TRM || tron[n + 3] == TRM2)) sigE = 0;
            o[4] = (int16_t) sigE;
        }
        int16_t s5 = sp->base.pat5.mtx ? (int16_t) (fs * patmat_nt(&sp->base.pat5, nt, len, n - sp->base.pat5.offset)) : 0;
        int16_t s3 = sp->base.pat3.mtx ? (int16_t) (fs * patmat_nt(&sp->base.pat3, nt, len, n - sp->base.pat3.offset)) : 0;
        s5 = (int16_t) (s5 + sp->base.sig53tab[int53[n] & 15]);
        s3 = (int16_t) (s3 + sp->base.sig53tab[16 + ((int53[n] >> 4) & 15)]);
        const int cano5 = (int53[n] >> 8) & 15, cano3 = (int53[n] >> 12) & 15;
        o[0] = s5;
        if (o[6] == -2 && cano5) {                          /* algmode.any == 2 adds a threshold rule: not restated */
            o[6] = 0;
            if (cano5 > 1) {
                o[8 + 6] = 1;
                if (n >= 1) o[-8 + 6] = o[-8 + 6] == 1 ? 2 : -1;
            }
        }
        o[1] = s3;
        if (o[7] == -2 && cano3) {
            o[7] = 0;
            if (cano3 > 1) {
                o[8 + 7] = 1;
                if (n >= 1) o[-8 + 7] = o[-8 + 7] == 1 ? 2 : -1;
            }
        }
    }
    free(nt); free(pot);
}
