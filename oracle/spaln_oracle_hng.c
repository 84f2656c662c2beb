/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the reference's SCALAR protein x genome
 * spliced DP kernel, the one Aln2h1::trcbkalignH_ng falls back to for blocks with fewer than 8
 * query rows (src/fwd2h1.cc:2007) and the `-A0` kernel in general.
 *   src/fwd2h1.cc:143-202   initH_ng          src/fwd2h1.cc:204-292  lastH_ng
 *   src/fwd2h1.cc:294-617   forwardH_ng (cutrng == 0)
 *   src/fwd2h1.cc:1997-2041 trcbkalignH_ng (scalar branch + end-point adjustment)
 *   src/codepot.cc:74-102   SpJunc::spjscr / spjseq, src/vmf.cc:66-140 Vmf
 * Pinned against the unmodified reference (tests/tools/sweep_oracle_scalar_p.py).
 */
#include <limits.h>
#include <stdlib.h>
#include "spaln_oracle.h"

#define NEVSEL32 (INT_MIN / 16 * 7)
enum { H_NCAND = 4, H_NQUE = 3 };
enum { DEAD, RSRV, DIAG, NEWD, VERT, SLA1, SLA2, VERL, HORI, HOR1, HOR2, HORL, NEWV, NEWH, SPIN = 16 };
static const int h_dir2nod[16] = { -1, -1, 0, 0, 2, 2, 2, 4, 1, 1, 1, 3, 2, 1, -1, -1 };
static const int h_nod2dir[5] = { DIAG, HORI, VERT, HORL, VERL };
static const int h_is_diag[16] = { 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
static const int h_is_vert[16] = { 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 1, 0, 0, 0 };
static const int h_is_hori[16] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 0, 1, 0, 0 };
static const unsigned char h_ncred[17] = { 15, 15, 0, 1, 4, 2, 5, 6, 10, 3, 7, 8, 10, 9, 12, 13, 14 };
static const int h_next_p[3] = { 1, 2, 0 };

typedef struct { int val, ptr, dir; } h_rvpd;
typedef struct { int val, ptr, dir, jnc; } h_cand;
typedef struct { int m, n, p; } h_rec;
typedef struct { h_rec* rec; int n, cap, fail; } h_vmf;

static int hv_add(h_vmf* v, int m, int n, int p)
{
    if (v->n == v->cap) {
        int nc = v->cap ? 2 * v->cap : 1024;
        h_rec* r = (h_rec*) realloc(v->rec, (size_t) nc * sizeof(h_rec));
        if (!r) { v->fail = 1; return 0; }
        v->rec = r; v->cap = nc;
    }
    v->rec[v->n].m = m; v->rec[v->n].n = n; v->rec[v->n].p = p;
    return v->n++;
}

/* SGPT6 accessors: 8 shorts per column (sig5, sig3, sigS, sigT, sigE, sigI, phs5, phs3) */
#define SG(t, n, f) ((t)->sgpt6[8 * (n) + (f)])
enum { F_SIG5, F_SIG3, F_SIGS, F_SIGT, F_SIGE, F_SIGI, F_PHS5, F_PHS3 };

static int gap_ext3(const so_params_h* p, int i) { return i > p->codonk1 ? p->lgep : p->gep; }

/* SpJunc::spjseq (src/codepot.cc:79-102): the two residues a split codon translates to */
static const uint8_t* spjseq(const so_ng_h* x, const so_task_h* t, int n5, int n3)
{
    const uint8_t* tab = x->spj_tabs, *amb_tron = tab + 514, *tron_amb = tab + 514 + 128, *aa2nuc = tab + 514 + 256;
    if (n5 < t->b_left || n3 >= t->b_right) return tab + 2 * 256;
    const uint8_t* b5 = t->b + (n5 - 2);
    static const uint8_t pyrim[2] = { 16, 16 };     /* PHE */
    const uint8_t* b3 = n3 ? t->b + n3 : pyrim;
#define NC(c) (h_ncred[aa2nuc[(c) < 26 ? (c) : 0] < 17 ? aa2nuc[(c) < 26 ? (c) : 0] : 0])
    int amb = 0;
    int c = NC(b5[0]);
    if (c >= 4) { amb = 1; c = 0; }
    unsigned w = (unsigned) c;
    if ((c = NC(b5[1])) < 4) {
        w = 4 * w + c;
        if ((c = NC(b3[0])) < 4) {
            w = 4 * w + c;
            if ((c = NC(b3[1])) < 4) w = 4 * w + c;
            else if (amb) w = 256;
            else amb = 2;
        } else w = 256;
    } else w = 256;
#undef NC
    if (amb == 0 || w == 256) return tab + 2 * w;
    if (amb == 1) return amb_tron + 2 * w;
    return tron_amb + 2 * w;
}

static int spjscr_h(const so_ng_h* x, const so_task_h* t, int n5, int n3)
{
    const int len = n3 - n5;
    int pen;
    if (len < 0) pen = x->penalty[0];
    else pen = len < x->n_penalty ? x->penalty[len] : x->penalty[x->n_penalty - 1];
    const int d5 = x->int53[n5] & 15, d3 = (x->int53[n3] >> 4) & 15;
    const int16_t sig = (int16_t) (SG(t, n3, F_SIG3) - x->sig53tab[16 + d3] + x->sig53tab[32 + 16 * d5 + d3]);
    return pen + sig;
}

int so_trcbk_h_ng(const so_params_h* p, const so_ng_h* x, const so_task_h* t, int32_t* score, int32_t* skl, int cap)
{
    const int width = t->up - t->lw + 7;
    *score = NEVSEL32;
    if (width < 0) return 0;
    if (!x || !x->penalty || !x->sig53tab || !x->int53 || !x->spj_tabs) return -3;
    const int noll = x->noll, dagp = noll == 3, nod = 2 * noll - 1;
    const int Local = p->local;
    const int LocalL = Local && t->a_exgl && t->b_exgl, LocalR = Local && t->a_exgr && t->b_exgr;
    const int a_left = t->a_left, a_right = t->a_right, b_left = t->b_left, b_right = t->b_right;
    const int lw = t->lw, up = t->up;
    const int spj = p->spj;
    const int gop_k[3] = { 0, p->gop, p->lgop };
    const int GapE1 = p->gape1, GapE2 = p->gape2, GapW1 = p->gw1, GapW2 = p->gw2, GapW3 = p->gw3, GapW3L = x->gw3l;
    const h_rvpd black = { NEVSEL32, 0, 0 };

    h_rvpd* buf = (h_rvpd*) malloc((size_t) 3 * (width + 8) * sizeof(h_rvpd));
    h_vmf vmf = { 0, 0, 0, 0 };
    if (!buf) return -1;
    for (int i = 0; i < 3 * (width + 8); ++i) buf[i] = black;
    h_rvpd* hh[3];
    hh[0] = buf - lw + 3;
    hh[1] = hh[0] + width;
    hh[2] = hh[1] + width;
    hv_add(&vmf, 0, 0, 0);

    /* ---- initH_ng ---- */
    {
        int n = b_left, r = b_left - 3 * a_left, rr = b_right - 3 * a_left;
        const int dir = t->a_exgl ? DEAD : DIAG;
        int jnc[3] = { n, 0, 0 };
        int bbn = n + 1;
        h_rvpd* h = hh[0] + r;
        h->val = (t->a_exgl && SG(t, bbn, F_SIGS) > 0) ? SG(t, bbn, F_SIGS) : 0;
        h->dir = dir;
        h->ptr = hv_add(&vmf, a_left, n, 0);
        if (t->a_exgl) {
            if (up < rr) rr = up;
            for (int i = 1; ++r <= rr; ++i) {
                ++h; ++bbn; ++n;
                if (i < 3) {
                    h->val = SG(t, bbn, F_SIGS) > 0 ? SG(t, bbn, F_SIGS) : 0;
                    h->dir = dir;
                    h->ptr = hv_add(&vmf, a_left, n, 0);
                    jnc[i] = n;
                } else {
                    *h = h[-3];
                    const int k = n - jnc[i % 3];
                    if (k == 3 && !(t->a_exgl & 1)) h->val += p->gop;
                    if (!(t->a_exgl & 2)) h->val += gap_ext3(p, k);
                    h->val += SG(t, bbn - 3, F_SIGE);
                    h->dir = HORI;
                    int xx = h[-1].val + GapW1;
                    if (xx > h->val) { *h = h[-1]; h->val = xx; h->dir = HOR1; }
                    xx = h[-2].val + GapW2;
                    if (xx > h->val) { *h = h[-2]; h->val = xx; h->dir = HOR2; }
                }
                const int xs = SG(t, bbn, F_SIGS) > 0 ? SG(t, bbn, F_SIGS) : 0;
                if (h->val < xs) {
                    h->val = xs;
                    h->dir = DEAD;
                    h->ptr = hv_add(&vmf, a_left, n, 0);
                    jnc[i % 3] = n;
                }
            }
        }
        r = b_left - 3 * a_left;
        rr = b_left - 3 * a_right;
        h = hh[0] + r - 1;
        if (lw > rr) rr = lw;
        for (int i = 1; --r >= rr; ++i, --h) {
            if (t->b_exgl == 1) { h->val = 0; h->dir = DEAD; h->ptr = 0; }
            else if (i <= 3) {
                *h = h[i];
                if (!(t->b_exgl & 2)) h->val += p->gep;
                if (!(t->b_exgl & 1)) h->val += p->gop;
                if (i < 3) h->val += x->extragop;
                h->dir = VERT;
            } else {
                *h = h[3];
                if (!(t->b_exgl & 2)) h->val += gap_ext3(p, i);
            }
        }
    }

    int best_val = NEVSEL32, best_m = a_left, best_n = b_left, best_p = 0;
    int m = a_left;
    if (!t->a_exgl) --m;
    int n1 = 3 * m + lw - 1, n2 = 3 * m + up;
    for (++m; m <= a_right; ++m) {
        n1 += 3; n2 += 3;
        const int n0 = n1 > b_left ? n1 : b_left;
        const int n9 = n2 < b_right ? n2 : b_right;
        int n = n0;
        int r = n - 3 * m;
        h_rvpd e1[2 * H_NQUE];
        h_rvpd* e2 = e1 + H_NQUE;
        for (int i = 0; i < 2 * H_NQUE; ++i) e1[i] = black;
        if (!t->b_exgl && m == a_left) {
            e1[2] = e2[2] = hh[0][r];
            e1[2].val = GapW3;
            e2[2].val = GapW3L;
        }
        const int32_t* qprof0 = p->simmtx + (size_t) t->a[m > 0 ? m - 1 : 0] * p->simdim;
        const int32_t* qprof1 = p->simmtx + (size_t) t->a[m] * p->simdim;
        h_cand hl[3][H_NCAND + 1];
        int nx[3][H_NCAND + 1];
        for (int ph = 0; ph < 3; ++ph)
            for (int l = 0; l <= H_NCAND; ++l) {
                hl[ph][l].val = NEVSEL32; hl[ph][l].ptr = hl[ph][l].dir = hl[ph][l].jnc = 0;
                nx[ph][l] = l;
            }
        int ncand[3] = { -1, -1, -1 };
        int sigB[3] = { 0, 0, 0 };                  /* src/fwd2h1.cc:352-354 */
        if (t->cip)
            for (int phs = -1; phs < 2; ++phs) sigB[phs + 1] = t->cip[3 * m - phs];
        h_rvpd* h = hh[0] + r;
        h_rvpd* f = hh[1] + r;
        h_rvpd* f2 = dagp ? hh[2] + r : 0;
        h_rvpd* hf[5];
        for (int q = 0; n <= n9; ++n, ++h, ++f) {
            const int bs = n - 2;                           /* b->at(n - 2) */
            const int sigE = n > b_left ? SG(t, n - 2, F_SIGE) : 0;
            h_rvpd* eq1 = e1 + q;
            h_rvpd* eq2 = dagp ? e2 + q : 0;
            hf[0] = h; hf[1] = eq1; hf[2] = f; hf[3] = eq2; hf[4] = f2;
            h_rvpd hq = *h;
            h_rvpd* from = h;
            h_rvpd* mx = h;
            int xv, yv;
            if (m != a_left) {
                if (n < b_left + 3) *h = black;
                else {
                    h->val += qprof0[t->b[bs]] + sigE;
                    h->dir = h_is_diag[from->dir & 15] ? DIAG : NEWD;
                }
                yv = f[3].val + p->gep;
                ++from;
                xv = from->val + (h_is_vert[from->dir & 15] ? GapE1 : GapW1);
                if (xv > yv) { f->val = xv; f->dir = SLA2; f->ptr = from->ptr; }
                else f->val = yv;
                ++from;
                xv = from->val + (h_is_vert[from->dir & 15] ? GapE2 : GapW2);
                if (xv > f->val) { f->val = xv; f->dir = SLA1; f->ptr = from->ptr; }
                ++from;
                xv = from->val + GapW3;
                if (xv >= f->val) { f->val = xv; f->dir = VERT; f->ptr = from->ptr; }
                else if (yv >= f->val) { f->val = yv; f->dir = VERT; f->ptr = f[3].ptr; }
                if (f->val > mx->val) mx = f;
                if (dagp) {
                    xv = from->val + GapW3L;
                    yv = f2[3].val + p->lgep;
                    if (xv >= yv) { f2->val = xv; f2->dir = VERL; f2->ptr = from->ptr; }
                    else { *f2 = f2[3]; f2->val = yv; }
                    if (f2->val > mx->val) mx = f2;
                }
            }
            /* horizontal moves */
            if (n > n0 + 2) {
                from = h - 3;
                xv = from->val + GapW3;
                yv = eq1->val += p->gep;
                if (xv > yv) { *eq1 = *from; eq1->val = xv; }
                eq1->val += sigE;
                eq1->dir = (eq1->dir & SPIN) + HORI;
                if (dagp) {
                    xv = from->val + GapW3L;
                    yv = eq2->val += p->lgep;
                    if (xv > yv) { *eq2 = *from; eq2->val = xv; }
                    eq2->val += sigE;
                    eq2->dir = (eq2->dir & SPIN) + HORL;
                    if (eq2->val > mx->val) mx = e2 + q;
                }
            }
            if (n > n0 + 1) {
                from = h - 2;
                xv = from->val + GapW2;
                if (xv > eq1->val) { *eq1 = *from; eq1->val = xv; eq1->dir = (eq1->dir & SPIN) + HOR2; }
            }
            from = h - 1;
            xv = from->val + GapW1;
            if (xv > eq1->val) { *eq1 = *from; eq1->val = xv; eq1->dir = (eq1->dir & SPIN) + HOR1; }
            if (eq1->val > mx->val) mx = e1 + q;
            if (++q == H_NQUE) q = 0;

            /* intron 3' boundary */
            const int phs3 = SG(t, n, F_PHS3);
            if (spj && phs3 > -2) {
                int phs = phs3 == 2 ? -1 : phs3;
                for (;;) {
                    const int nb = n - phs;
                    const int* pnx = nx[phs + 1];
                    const h_cand* top[5] = { 0, 0, 0, 0, 0 };
                    for (int l = 0; l <= ncand[phs + 1]; ++l) {
                        const h_cand* phl = hl[phs + 1] + pnx[l];
                        if (phs == 1 && phl->dir == 2) continue;
                        if (nb - phl->jnc < x->minl) continue;
                        xv = phl->val + sigB[phs + 1] + spjscr_h(x, t, phl->jnc, nb);
                        if (phl->dir == 0 && phs) {
                            const uint8_t* cs = spjseq(x, t, phl->jnc, nb);
                            if (phs == 1) xv += qprof0[cs[0]];
                            else xv += qprof1[cs[1]] - qprof1[t->b[bs + 3]] - SG(t, n + 1, F_SIGE);
                        }
                        from = hf[phl->dir];
                        if (xv > from->val) { from->val = xv; top[phl->dir] = phl; }
                    }
                    for (int d = 0; d < nod; ++d) {
                        const h_cand* phl = top[d];
                        if (!phl) continue;
                        from = hf[d];
                        from->ptr = hv_add(&vmf, m, n, hv_add(&vmf, m, phl->jnc + phs, phl->ptr));
                        from->dir = h_nod2dir[phl->dir] | SPIN;
                        if (from->val > mx->val) mx = from;
                    }
                    if (phs3 - phs == 3) { phs = 1; continue; }     /* AGAG */
                    break;
                }
            }

            /* best state */
            yv = h->val;
            if (h != mx) *h = *mx;
            else if (Local && yv > hq.val) {
                if (LocalL && hq.dir == 0 && !(h->dir & SPIN)) h->ptr = hv_add(&vmf, m - 1, n - 3, 0);
                else if (LocalR && yv > best_val) { best_val = yv; best_p = h->ptr; best_m = m; best_n = n; }
            }
            if (LocalL && h->val <= 0) h->val = h->dir = 0;
            else if (h->dir == NEWD) h->ptr = hv_add(&vmf, m - 1, n - 3, h->ptr);

            /* intron 5' boundary */
            const int phs5 = SG(t, n, F_PHS5);
            if (spj && phs5 > -2) {
                int phs = phs5 == 2 ? -1 : phs5;
                for (;;) {
                    const int nb = n - phs;
                    const int sigJ = SG(t, nb, F_SIG5);
                    const int hd = h_dir2nod[mx->dir & 15];
                    for (int k = (hd == 0 || phs == 1) ? 0 : 1; k < nod; ++k) {
                        const int crossspj = phs == 1 && k == 0;
                        from = crossspj ? &hq : hf[k];
                        if (!from->dir || (from->dir & SPIN)) continue;
                        if (!crossspj && k != hd && hd >= 0) {
                            yv = mx->val;
                            if (hd == 0 || (k - hd) % 2) yv += gop_k[k / 2];
                            if (from->val <= yv) continue;
                        }
                        xv = from->val + sigJ;
                        h_cand* phl = hl[phs + 1];
                        int* pnx = nx[phs + 1];
                        int* nc = &ncand[phs + 1];
                        int l = *nc < H_NCAND ? ++*nc : H_NCAND;
                        while (--l >= 0) {
                            if (xv >= phl[pnx[l]].val) { int s = pnx[l]; pnx[l] = pnx[l + 1]; pnx[l + 1] = s; }
                            else break;
                        }
                        if (++l < H_NCAND) {
                            phl += pnx[l];
                            phl->val = xv; phl->jnc = nb; phl->dir = k; phl->ptr = from->ptr;
                        } else --*nc;
                    }
                    if (phs5 - phs == 3) { phs = 1; continue; }     /* GTGT */
                    break;
                }
            }
            if (f2) ++f2;
        }
    }

    int ptr = 0, val;
    if (!LocalR || best_m == a_right) {
        /* ---- lastH_ng ---- */
        int glen[3] = { 0, 0, 0 };
        int rw = lw;
        const int m3 = 3 * a_right;
        int rf = b_left - m3;
        if (rf > rw) rw = rf; else rf = rw;
        h_rvpd* h = hh[0] + rw;
        h_rvpd* h9 = hh[0] + b_right - m3;
        h_rvpd* mx = h9;
        int bbn = rw + m3;
        if (t->a_exgr) {
            for (int ph = 0; h <= h9; ++h, ++bbn, ++rf, ph = h_next_p[ph]) {
                glen[ph] += 3;
                int cand[3] = { h->val, NEVSEL32, NEVSEL32 };
                if (rf - rw >= 3 && h[-3].dir != DEAD) {
                    cand[1] = h[-3].val + SG(t, bbn - 2, F_SIGE);
                    if (!(t->a_exgr & 2)) cand[1] += gap_ext3(p, glen[ph]);
                    if (!(t->a_exgr & 1) && glen[ph] == 3) cand[1] += p->gop;
                    if (SG(t, bbn - 2, F_SIGT) > 0 && !(h->dir & SPIN)) cand[2] = h[-3].val + SG(t, bbn - 2, F_SIGT);
                }
                const int sig5 = (Local && SG(t, bbn, F_SIG5) > 0) ? SG(t, bbn, F_SIG5) : 0;
                cand[0] += sig5;
                cand[1] += sig5;
                int k = 0;
                if (cand[1] > cand[k]) k = 1;
                if (cand[2] > cand[k]) k = 2;
                if (k == 0) { if (!h_is_hori[h->dir & 15]) glen[ph] = 0; }
                else if (k == 1) { *h = h[-3]; h->dir = HORI; h->val = cand[k] - sig5; }
                else {
                    *h = h[-3];
                    h->dir = DEAD;
                    h->val = cand[k];
                    if (h->val > mx->val) h->ptr = hv_add(&vmf, a_right, rf + m3 - 3, h->ptr);
                }
                if (h->val > mx->val) mx = h;
            }
        } else {
            bbn += (int) (h9 - h);
            const int yv = h9[-3].val + SG(t, bbn - 2, F_SIGT);
            if (yv > h9->val) { *h9 = h9[-3]; h9->val = yv; h9->dir = HORI; }
        }
        int done = 0;
        if (t->b_exgr == 1) {
            rw = up < b_right - 3 * a_left ? up : b_right - 3 * a_left;
            int g[3] = { NEVSEL32, NEVSEL32, NEVSEL32 };
            h = hh[0] + rw - 3;
            for (int ph = 0; h >= h9; --h) {
                int xv = h[3].val;
                if (!(t->b_exgr & 1)) xv += p->gop;
                if (xv > g[ph]) g[ph] = xv;
                if (!(t->b_exgr & 2)) g[ph] += p->gep;
                if (h->val > g[ph]) g[ph] = NEVSEL32;
                else if (g[ph] > mx->val) { mx = h; mx->val = g[ph]; }
                if (++ph == 3) ph = 0;
            }
        } else if (t->b_exgr == 2) {
            mx = hh[1] + b_right - m3;
            mx->ptr = hv_add(&vmf, a_right, b_right, mx->ptr);
            done = 1;
        }
        if (!done) {
            int pp = (int) (mx - h9);
            int m9 = a_right, n9 = b_right;
            if (pp > 0) { m9 -= (pp + 2) / 3; if (pp %= 3) n9 -= 3 - pp; }
            else if (pp < 0) n9 += pp;
            mx->ptr = hv_add(&vmf, m9, n9, mx->ptr);
        }
        val = mx->val;
        ptr = mx->ptr;
    } else {
        ptr = hv_add(&vmf, best_m, best_n, best_p);
        val = best_val;
    }

    int cnt = 0;
    if (vmf.fail) cnt = -1;
    else if (ptr) {
        int m_last = 0, n_last = 0;
        for (int q = ptr; ; q = vmf.rec[q].p) {
            m_last = vmf.rec[q].m; n_last = vmf.rec[q].n;
            if (cnt < cap) { skl[2 * cnt] = m_last; skl[2 * cnt + 1] = n_last; }
            ++cnt;
            if (!vmf.rec[q].p) break;
        }
        const int rd = Local ? 0 : (n_last - 3 * m_last) - b_left + 3 * a_left;
        if (rd) {
            const int mm = rd > 0 ? a_left : a_left - rd / 3, nn = rd > 0 ? b_left + rd : b_left;
            if (cnt < cap) { skl[2 * cnt] = mm; skl[2 * cnt + 1] = nn; }
            ++cnt;
        }
    }
    *score = val;
    free(buf); free(vmf.rec);
    return cnt;
}

/* =====================================================================================
 * Aln2h1::hirschbergH_ng (src/fwd2h1.cc:1085-1520) with hinitH_ng / hlastH_ng (941-1083): the scalar
 * unidirectional Hirschberg pass of `-A0` for protein queries.  Cell states carry {value,
 * direction, highest / lowest diagonal since the last intermediate row, start row, link}; the
 * intermediates are the bounded form of src/udh_intermediate.h:29-88.  imd_intvl: the member
 * lspH_ng sets (src/fwd2h1.cc:2170-2183).  cpos: (n_im + 1) x 10 ints, all end_of_ulk on entry
 * (src/fwd2h1.cc:2192-2193).  Returns 0, -1 allocation failure, -3 missing tables.
 * Pinned against the unmodified reference (tests/golden/prot_A0_udh*.npz).
 * ===================================================================================== */
#define END_OF_ULK (INT_MAX - 2)
typedef struct { int val, dir, upr, lwr, ml, ulk; } hu_cell;            /* Rvdwml, src/aln.h:138-145 */
typedef struct { int val, dir, upr, lwr, ml, ulk, jnc; } hu_cand;       /* Rvdwmlj */
typedef struct { int mi; int* buf; int *hlnk[3], *vlnk[3], *lwrb[3], *uprb[3]; } hu_imd;

static int hu_imd_init(hu_imd* im, int mi, int lw, int width, int nol)
{
    const size_t u = (size_t) nol * width;
    im->mi = mi;
    im->buf = (int*) malloc(4 * u * sizeof(int));
    if (!im->buf) return -1;
    for (size_t i = 0; i < 2 * u; ++i) im->buf[i] = END_OF_ULK;
    for (size_t i = 0; i < u; ++i) { im->buf[2 * u + i] = INT_MAX; im->buf[3 * u + i] = INT_MIN; }
    im->hlnk[0] = im->buf - lw + 1;         /* UdhIntermediate biases by lw - 1 for every sequence type */
    im->vlnk[0] = im->hlnk[0] + u;
    im->lwrb[0] = im->vlnk[0] + u;
    im->uprb[0] = im->lwrb[0] + u;
    for (int k = 1; k < nol; ++k) {
        im->hlnk[k] = im->hlnk[k - 1] + width; im->vlnk[k] = im->vlnk[k - 1] + width;
        im->lwrb[k] = im->lwrb[k - 1] + width; im->uprb[k] = im->uprb[k - 1] + width;
    }
    return 0;
}

static int hu_min(int a, int b) { return a < b ? a : b; }
static int hu_max(int a, int b) { return a > b ? a : b; }

int so_hirschberg_h_ng(const so_params_h* p, const so_ng_h* x, const so_task_h* t, int n_im, int imd_intvl,
                       int32_t* score, int32_t* cpos, int32_t* ranges)
{
    *score = NEVSEL32;
    if (n_im < 1 || imd_intvl < 1) return -3;
    if (!x || !x->penalty || !x->sig53tab || !x->int53 || !x->spj_tabs) return -3;
    const int width = t->up - t->lw + 7;
    const int noll = x->noll, dagp = noll == 3, nod = 2 * noll - 1;
    const int Local = p->local;
    const int LocalL = Local && t->a_exgl && t->b_exgl, LocalR = Local && t->a_exgr && t->b_exgr;
    int a_left = t->a_left, a_right = t->a_right, b_left = t->b_left, b_right = t->b_right;
    const int lw = t->lw, up = t->up;
    const int spj = p->spj;
    const int gop_k[3] = { 0, p->gop, p->lgop };
    const int GapE1 = p->gape1, GapE2 = p->gape2, GapW1 = p->gw1, GapW2 = p->gw2, GapW3 = p->gw3, GapW3L = x->gw3l;

    const size_t bufsiz = (size_t) noll * width;
    hu_cell* wbuf = (hu_cell*) malloc(bufsiz * sizeof(hu_cell));
    hu_imd* imds = (hu_imd*) calloc(n_im, sizeof(hu_imd));
    if (!wbuf || !imds) { free(wbuf); free(imds); return -1; }
    const int r_black = b_left - 3 * a_right;
    const hu_cell black = { NEVSEL32, 0, r_black, r_black, 0, END_OF_ULK };
    const hu_cand black_cand = { NEVSEL32, 0, INT_MIN, INT_MAX, 0, END_OF_ULK, 0 };
    for (size_t i = 0; i < bufsiz; ++i) wbuf[i] = black;
    hu_cell* hhg[3];
    hhg[0] = wbuf - lw + 3;
    hhg[1] = hhg[0] + width;
    hhg[2] = dagp ? hhg[1] + width : 0;
    hu_cell* blackvdwuj = wbuf + bufsiz - 1;
    for (int i = 0; i <= n_im; ++i) for (int j = 0; j < 10; ++j) cpos[10 * i + j] = END_OF_ULK;

    /* ---- hinitH_ng ---- */
    {
        int n = b_left, r = b_left - 3 * a_left;
        const int r0 = r;
        int rr = b_right - 3 * a_left;
        const int dir = t->a_exgl ? DEAD : DIAG;
        int bbn = n + 1;
        hu_cell* h = hhg[0] + r;
        h->val = (t->a_exgl && SG(t, bbn, F_SIGS) > 0) ? SG(t, bbn, F_SIGS) : 0;
        h->dir = dir;
        h->lwr = h->upr = h->ulk = r0;
        h->ml = a_left;
        if (t->a_exgl) {
            if (up < rr) rr = up;
            int jnc = n;
            for (int i = 1; ++r <= rr; ++i) {
                ++h; ++bbn; ++n;
                if (i < 3) {
                    h->val = SG(t, bbn, F_SIGS) > 0 ? SG(t, bbn, F_SIGS) : 0;
                    h->dir = dir;
                    h->lwr = h->ulk = r;
                    h->ml = a_left;
                } else {
                    *h = h[-3];
                    const int d = n - jnc;
                    if (!(t->a_exgl & 1) && d == 3) h->val += p->gop;
                    if (!(t->a_exgl & 2)) h->val += gap_ext3(p, d);
                    h->val += SG(t, bbn - 3, F_SIGE);
                    h->dir = HORI;
                    int xx = h[-1].val + GapW1;
                    if (xx > h->val) { *h = h[-1]; h->val = xx; h->dir = HOR1; }
                    xx = h[-2].val + GapW2;
                    if (xx > h->val) { *h = h[-2]; h->val = xx; h->dir = HOR2; }
                }
                const int xs = SG(t, bbn, F_SIGS) > 0 ? SG(t, bbn, F_SIGS) : 0;
                if (h->val < xs) {          /* start codon */
                    h->val = xs;
                    h->dir = DEAD;
                    jnc = n;
                    h->lwr = h->ulk = r;
                }
                h->upr = r;
            }
        }
        r = r0;
        rr = b_left - 3 * a_right;
        if (lw > rr) rr = lw;
        h = hhg[0] + r - 1;
        for (int i = 1; --r >= rr; ++i, --h) {
            if (t->b_exgl == 1) {
                h->val = 0; h->dir = DEAD;
                h->upr = h->lwr = h->ulk = r;
                h->ml = a_left + i / 3;
            } else if (i <= 3) {
                *h = h[i];
                if (!(t->b_exgl & 2)) h->val += p->gep;
                if (!(t->b_exgl & 1)) h->val += p->gop;
                if (i < 3) h->val += x->extragop;
                h->dir = VERT;
                h->ml += i / 3;
                h->lwr = h->ulk = r;
            } else {
                *h = h[3];
                if (!(t->b_exgl & 2)) h->val += gap_ext3(p, i);
                h->lwr = h->ulk = r;
                ++h->ml;
            }
        }
    }

    {
        int mi = a_left;
        for (int i = 0; i < n_im; ++i)
            if (hu_imd_init(&imds[i], mi += imd_intvl, lw, width, noll)) return -1;
    }
    hu_imd* imd = &imds[0];
    int mm = imd->mi;
    int rlst[3] = { INT_MAX, INT_MAX, INT_MAX };
    struct { int val, upr, lwr, ml, ulk, mr, nr; } maxh = { NEVSEL32, 0, 0, a_left, 0, a_right, b_right };

    int m = a_left;
    if (!t->a_exgl) --m;
    int n1 = 3 * m + lw - 1, n2 = 3 * m + up;
    int ii = 0;
    for (++m; m <= a_right; ++m) {
        n1 += 3; n2 += 3;
        const int n0 = n1 > b_left ? n1 : b_left;
        const int n9 = n2 < b_right ? n2 : b_right;
        const int is_imd = m == mm;
        int n = n0;
        int r = n - 3 * m;
        hu_cell e1[2 * H_NQUE];
        hu_cell* e2 = e1 + H_NQUE;
        for (int i = 0; i < 2 * H_NQUE; ++i) e1[i] = black;
        if (!t->b_exgl && m == a_left) {
            e1[2] = e2[2] = hhg[0][r];
            e1[2].val += GapW3;
            e2[2].val += GapW3L;
        }
        const int32_t* qprof0 = p->simmtx + (size_t) t->a[m > 0 ? m - 1 : 0] * p->simdim;
        const int32_t* qprof1 = p->simmtx + (size_t) t->a[m] * p->simdim;
        hu_cand hl[3][H_NCAND + 1];
        int nx[3][H_NCAND + 1];
        for (int ph = 0; ph < 3; ++ph)
            for (int l = 0; l <= H_NCAND; ++l) { hl[ph][l] = black_cand; nx[ph][l] = l; }
        int ncand[3] = { -1, -1, -1 };
        int sigB[3] = { 0, 0, 0 };
        if (t->cip)
            for (int phs = -1; phs < 2; ++phs) sigB[phs + 1] = t->cip[3 * m - phs];
        hu_cell* h = hhg[0] + r;
        hu_cell* f = hhg[1] + r;
        hu_cell* f2 = dagp ? hhg[2] + r : blackvdwuj;
        hu_cell* hf[5];
        for (int q = 0; n <= n9; ++n, ++r) {
            const int bs = n - 2;
            const int sigE = n > b_left ? SG(t, n - 2, F_SIGE) : 0;
            hu_cell* eq1 = e1 + q;
            hu_cell* eq2 = dagp ? e2 + q : 0;
            hf[0] = h; hf[1] = eq1; hf[2] = f; hf[3] = eq2; hf[4] = f2;
            const hu_cell hq = *h;
            hu_cell* from = h;
            hu_cell* mx = h;
            int xv, yv;
            if (m != a_left) {
                if (n < b_left + 3) *h = black;
                else {
                    h->val += qprof0[t->b[bs]] + sigE;
                    h->dir = (from->dir & DIAG) ? DIAG : NEWD;      /* a bit test in this function */
                }
                yv = f[3].val + p->gep;
                ++from;
                xv = from->val + (h_is_vert[from->dir & 15] ? GapE1 : GapW1);
                if (xv > yv) { *f = *from; f->val = xv; f->dir = SLA2; }
                else f->val = yv;
                ++from;
                xv = from->val + (h_is_vert[from->dir & 15] ? GapE2 : GapW2);
                if (xv > f->val) { *f = *from; f->val = xv; f->dir = SLA1; }
                ++from;
                xv = from->val + GapW3;
                if (xv >= f->val) { *f = *from; f->val = xv; f->dir = VERT; }
                else if (yv >= f->val) { *f = f[3]; f->val = yv; f->dir = VERT; }
                if (f->val >= mx->val) mx = f;
                if (dagp) {
                    xv = from->val + GapW3L;
                    yv = f2[3].val + p->lgep;
                    if (xv >= yv) { *f2 = *from; f2->val = xv; f2->dir = VERL; }
                    else { *f2 = f2[3]; f2->val = yv; }
                    if (f2->val >= mx->val) mx = f2;
                }
            }
            /* horizontal moves */
            if (n > n0 + 2) {
                from = h - 3;
                xv = from->val + GapW3;
                yv = eq1->val += p->gep;
                if (xv > yv) { *eq1 = *from; eq1->val = xv; }
                eq1->val += sigE;
                eq1->dir = (eq1->dir & SPIN) + HORI;
                if (dagp) {
                    xv = from->val + GapW3L;
                    yv = eq2->val += p->lgep;
                    if (xv > yv) { *eq2 = *from; eq2->val = xv; }
                    eq2->val += sigE;
                    eq2->dir = (eq2->dir & SPIN) + HORL;
                    if (eq2->val > mx->val) mx = eq2;
                }
            }
            if (n > n0 + 1) {
                from = h - 2;
                xv = from->val + GapW2;
                if (xv > eq1->val) { *eq1 = *from; eq1->val = xv; eq1->dir = HOR2; }
            }
            from = h - 1;
            xv = from->val + GapW1;
            if (xv > eq1->val) { *eq1 = *from; eq1->val = xv; eq1->dir = HOR1; }
            if (eq1->val > mx->val) mx = eq1;
            if (++q == H_NQUE) q = 0;

            /* intron 3' boundary */
            int spj3 = 0;
            const int phs3 = SG(t, n, F_PHS3);
            if (spj && phs3 > -2) {
                int phs = phs3 == 2 ? -1 : phs3;
                for (;;) {
                    const int nb = n - phs;
                    const int* pnx = nx[phs + 1];
                    const hu_cand* top[5] = { 0, 0, 0, 0, 0 };
                    for (int l = 0; l <= ncand[phs + 1]; ++l) {
                        const hu_cand* phl = hl[phs + 1] + pnx[l];
                        if (phs == 1 && phl->dir == 2) continue;
                        if (nb - phl->jnc < x->minl) continue;
                        xv = phl->val + sigB[phs + 1] + spjscr_h(x, t, phl->jnc, nb);
                        if (phl->dir == 0 && phs) {
                            const uint8_t* cs = spjseq(x, t, phl->jnc, nb);
                            if (phs == 1) xv += qprof0[cs[0]];
                            else xv += qprof1[cs[1]] - qprof1[t->b[bs + 3]] - SG(t, n + 1, F_SIGE);
                        }
                        from = hf[phl->dir];
                        if (xv > from->val) { from->val = xv; top[phl->dir] = phl; }
                    }
                    int maxk = nod;
                    for (int k = 0; k < nod; ++k) {
                        const hu_cand* phl = top[k];
                        if (!phl) continue;
                        from = hf[k];
                        from->dir = h_nod2dir[phl->dir] | SPIN;
                        from->upr = hu_max(phl->upr, r);
                        from->lwr = hu_min(phl->lwr, r);
                        from->ml = phl->ml;
                        from->ulk = phl->ulk;
                        if (from->val >= mx->val) { maxk = k; mx = from; }
                    }
                    if (is_imd && maxk < nod) {
                        const hu_cand* phl = top[maxk];
                        imd->hlnk[0][r] = phl->ulk;
                        mx->ulk = rlst[q] = r;
                        spj3 = 1;
                        if (maxk == 0) {
                            for (int c = 1, d = 1; c < noll; ++c, d += 2) {
                                if ((phl = top[d]) && hf[d]->val > mx->val + gop_k[c]) {
                                    hf[d]->ulk = r + c * width;
                                    imd->hlnk[c][r] = phl->ulk;
                                }
                                if (top[d + 1] && hf[d + 1]->val > mx->val + gop_k[c]) hf[d + 1]->ulk = r + c * width;
                            }
                        }
                    }
                    if (phs3 - phs == 3) { phs = 1; continue; }     /* AGAG */
                    break;
                }
            }

            /* best state */
            yv = h->val;
            if (h == mx) {
                if (LocalR && yv > maxh.val) {
                    maxh.val = h->val; maxh.upr = h->upr; maxh.lwr = h->lwr; maxh.ml = h->ml;
                    maxh.ulk = h->ulk; maxh.mr = m; maxh.nr = n;
                }
            } else {
                if (mx->upr < r) mx->upr = r;       /* the source state is widened, then copied */
                if (mx->lwr > r) mx->lwr = r;
                *h = *mx;
            }
            if (LocalL && h->val <= 0) {
                h->val = h->dir = 0;
                h->ml = m;
                h->ulk = h->upr = h->lwr = r;
            }

            /* intron 5' boundary */
            const int hd = h_dir2nod[mx->dir & 15];
            const int phs5 = SG(t, n, F_PHS5);
            if (spj && phs5 > -2) {
                int phs = phs5 == 2 ? -1 : phs5;
                for (;;) {
                    const int nb = n - phs;
                    const int sigJ = SG(t, nb, F_SIG5);
                    for (int k = (hd == 0 || phs == 1) ? 0 : 1; k < nod; ++k) {
                        const int crossspj = phs == 1 && k == 0;
                        const hu_cell* src = crossspj ? &hq : hf[k];
                        if (!src->dir || (src->dir & SPIN)) continue;
                        if (k != hd && !crossspj && hd >= 0) {
                            yv = mx->val;
                            if (hd == 0 || (k - hd) % 2) yv += gop_k[k / 2];
                            if (src->val <= yv) continue;
                        }
                        xv = src->val + sigJ;
                        hu_cand* phl = hl[phs + 1];
                        int* pnx = nx[phs + 1];
                        int* nc = &ncand[phs + 1];
                        int l = *nc < H_NCAND ? ++*nc : H_NCAND;
                        while (--l >= 0) {
                            if (xv >= phl[pnx[l]].val) { int s = pnx[l]; pnx[l] = pnx[l + 1]; pnx[l + 1] = s; }
                            else break;
                        }
                        if (++l < H_NCAND) {
                            phl += pnx[l];
                            phl->val = xv; phl->jnc = nb; phl->dir = k;
                            phl->upr = src->upr; phl->lwr = src->lwr; phl->ml = src->ml;
                            if (is_imd) {
                                if (k == 1) imd->hlnk[0][r] = rlst[q];
                                phl->ulk = r;
                            } else
                                phl->ulk = src->ulk;
                        } else --*nc;
                    }
                    if (phs5 - phs == 3) { phs = 1; continue; }     /* GTGT */
                    break;
                }
            }

            /* intermediate row */
            if (is_imd) {
                if (hd == 0) rlst[q] = r;
                else if (!spj3 && hd % 2) imd->hlnk[0][r] = rlst[q];
                for (int k = 0; k < noll; ++k) {
                    hu_cell* s = hf[2 * k];
                    imd->vlnk[k][r] = s->ulk;
                    imd->lwrb[k][r] = hu_min(r, s->lwr);
                    imd->uprb[k][r] = hu_max(r, s->upr);
                    s->lwr = s->upr = r;
                    s->ulk = r + k * width;
                }
            }
            ++h; ++f;
            if (dagp) ++f2;
        }
        if (is_imd && ++ii < n_im) { imd = &imds[ii]; mm = imd->mi; }
    }

    const int rr = b_right - 3 * a_right;
    int r;
    if (LocalR) {
        int i = n_im;
        while (--i >= 0 && imds[i].mi > a_right) ;
        a_right = maxh.mr;
        b_right = maxh.nr;
        if (i < 0) i = 0;
        cpos[10 * i + 8] = maxh.lwr;
        cpos[10 * i + 9] = maxh.upr;
    } else {
        /* ---- hlastH_ng ---- */
        int glen[3] = { 0, 0, 0 };
        const int m3 = 3 * a_right;
        int rw = lw;
        int rf = b_left - m3;
        if (rf > rw) rw = rf; else rf = rw;
        hu_cell* h = hhg[0] + rw;
        hu_cell* h9 = hhg[0] + b_right - m3;
        hu_cell* mx = h9;
        int bbn = rw + m3;
        if (t->a_exgr) {
            for (int ph = 0; h <= h9; ++h, ++bbn, ++rf, ph = h_next_p[ph]) {
                glen[ph] += 3;
                int cand[3] = { h->val, NEVSEL32, NEVSEL32 };
                if (rf - rw >= 3 && h[-3].dir != DEAD) {
                    cand[1] = h[-3].val + SG(t, bbn - 2, F_SIGE);
                    if (!(t->a_exgr & 2)) cand[1] += gap_ext3(p, glen[ph]);
                    if (glen[ph] == 3 && !(t->a_exgr & 1)) cand[1] += p->gop;
                    if ((p->lcl & 2) && !(h->dir & SPIN)) cand[2] = h[-3].val + SG(t, bbn - 2, F_SIGT);
                }
                const int sig5 = (Local && SG(t, bbn, F_SIG5) > 0) ? SG(t, bbn, F_SIG5) : 0;
                cand[0] += sig5;
                cand[1] += sig5;
                int k = 0;
                if (cand[1] > cand[k]) k = 1;
                if (cand[2] > cand[k]) k = 2;
                if (k == 0) { if (!h_is_hori[h->dir & 15]) glen[ph] = 0; }
                else if (k == 1) { *h = h[-3]; h->dir = HORI; h->val = cand[k] - sig5; }
                else {
                    *h = h[-3];
                    h->dir = DEAD;
                    h->val = cand[k];
                    h->upr = hu_max(rf, h->upr);
                }
                if (h->val > mx->val) mx = h;
            }
        } else {
            bbn += (int) (h9 - h);
            const int yv = h9[-3].val + SG(t, bbn - 2, F_SIGT);
            if (yv > h9->val) {
                *h9 = h9[-3];
                h9->val = yv;
                h9->dir = HORI;
                h9->upr = hu_max(b_right - m3, h9->upr);
            }
        }
        if (t->b_exgr == 1) {
            rw = hu_min(up, b_right - 3 * a_left);
            for (h = hhg[0] + rw; h > h9; --h, --rw) {
                const int xv = h->val + (rw % 3 ? x->extragop : 0);
                if (xv > mx->val) { mx = h; mx->val = xv; }
            }
        } else if (t->b_exgr == 2)
            mx = hhg[1] + b_right - m3;
        maxh.val = mx->val; maxh.lwr = mx->lwr; maxh.upr = mx->upr; maxh.ulk = mx->ulk; maxh.ml = mx->ml;
        r = (int) (mx - hhg[0]);
        if (t->b_exgr && rr < r) a_right = (b_right - r) / 3;
        if (t->a_exgr && rr > r) b_right = 3 * a_right + r;
    }

    int i = n_im;
    while (--i >= 0 && imds[i].mi > a_right) ;
    if (i < 0 && imds[0].mi > a_right) cpos[2] = b_right;
    r = b_right - 3 * a_right;
    cpos[10 * (i + 1) + 8] = hu_min(maxh.lwr, r);
    cpos[10 * (i + 1) + 9] = hu_max(maxh.upr, r);

    r = maxh.ulk;
    for ( ; i >= 0 && (imd = &imds[i])->mi > maxh.ml; --i) {
        int c = 0, d = 0;
        for ( ; r > up; r -= width) ++d;
        if (d >= noll || r < lw - 1) { cpos[10 * i] = END_OF_ULK; break; }      /* (a link no pass wrote) */
        if (imd->vlnk[d][r] < END_OF_ULK) {
            cpos[10 * i + c++] = imd->mi;
            cpos[10 * i + c++] = d > 0 ? 1 : 0;
            const int mm3 = 3 * imd->mi;
            for (int rp = imd->hlnk[d][r]; lw <= rp && rp < up && r != rp; rp = imd->hlnk[0][r = rp])
                if (c < 7) cpos[10 * i + c++] = r + mm3;
            if (c < 8) cpos[10 * i + c++] = r + mm3;
            cpos[10 * i + c] = END_OF_ULK;
            cpos[10 * i + 8] = imd->lwrb[d][r];
            cpos[10 * i + 9] = imd->uprb[d][r];
            r = imd->vlnk[d][r];
            if (r == END_OF_ULK) break;
        } else
            cpos[10 * i] = END_OF_ULK;
    }
    for ( ; r > up; r -= width) ;
    if (LocalL) {
        a_left = maxh.ml;
        b_left = r + 3 * maxh.ml;
    } else {
        const int rl = b_left - 3 * a_left;
        if (t->b_exgl && rl > r) {
            a_left = (b_left - r) / 3;
            for (int j = 0; j < n_im && imds[j].mi < a_left; ++j) cpos[10 * j] = END_OF_ULK;
        }
        if (t->a_exgl && rl < r) b_left = 3 * a_left + r;
    }
    ++i;
    if ((i < n_im && imds[i].mi < a_left) || cpos[10 * i + 2] < b_left) maxh.val = NEVSEL32;
    else {
        r = b_left - 3 * a_left;
        cpos[10 * i + 8] = hu_min(r, cpos[10 * i + 8]);
        cpos[10 * i + 9] = hu_max(r, cpos[10 * i + 9]);
    }
    *score = maxh.val;
    ranges[0] = a_left; ranges[1] = a_right; ranges[2] = b_left; ranges[3] = b_right;
    for (int k = 0; k < n_im; ++k) free(imds[k].buf);
    free(imds); free(wbuf);
    return 0;
}
