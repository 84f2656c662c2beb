/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the reference's SCALAR spliced DP
 * kernel with exact intron scoring, the one Aln2s1::trcbkalignS_ng falls back to for blocks
 * with fewer than 8 query rows (src/fwd2s1.cc:1676) and the `-A0` kernel in general.
 *
 * Reference code restated (paths relative to /root/reference):
 *   src/fwd2s1.cc:141-215   initS_ng / lastS_ng
 *   src/fwd2s1.cc:217-444   forwardS_ng (cutrng == 0)
 *   src/fwd2s1.cc:1667-1710 trcbkalignS_ng (scalar branch + end-point adjustment)
 *   src/vmf.cc:66-140       Vmf::add / traceback
 *   src/codepot.cc:74-77, 401-416  SpJunc::spjscr, Exinon::sig53(IE5 | IE53), alprm2.Z == 0
 * Pinned against the unmodified reference (tests/test_oracle_vs_reference.py, tests/golden/).
 */
#include <limits.h>
#include <stdlib.h>
#include "spaln_oracle.h"

#define NEVSEL32 (INT_MIN / 16 * 7)     /* src/cmn.h:79 */
enum { NG_NCAND = 4, NG_NEWD = 8 };     /* src/aln.h:55, src/fwd2s1.cc:48 */
static const int ng_psp_bit[5] = { 4, 1, 8, 2, 16 };   /* src/aln.h:56 */

typedef struct { int val, ptr; } ng_rvp;
typedef struct { int val, ptr, dir, jnc; } ng_cand;
typedef struct { int m, n, p; } ng_rec;
typedef struct { ng_rec* rec; int n, cap, fail; } ng_vmf;

static int vmf_add(ng_vmf* v, int m, int n, int p)
{
    if (v->n == v->cap) {
        int nc = v->cap ? 2 * v->cap : 1024;
        ng_rec* r = (ng_rec*) realloc(v->rec, (size_t) nc * sizeof(ng_rec));
        if (!r) { v->fail = 1; return 0; }
        v->rec = r; v->cap = nc;
    }
    v->rec[v->n].m = m; v->rec[v->n].n = n; v->rec[v->n].p = p;
    return v->n++;
}

static int cano5(const so_task* t, int n) { return (t->int53[n] >> 8) & 15; }
static int cano3(const so_task* t, int n) { return (t->int53[n] >> 12) & 15; }

/* SpJunc::spjscr(n5, n3): length penalty + pair-corrected 3' signal (narrowed to STYPE) */
static int spjscr(const so_params* p, const so_task* t, int n5, int n3)
{
    const int len = n3 - n5;
    const int pen = len < p->n_penalty ? p->penalty[len] : p->penalty[p->n_penalty - 1];
    const int d5 = t->int53[n5] & 15, d3 = (t->int53[n3] >> 4) & 15;
    const int16_t sig = (int16_t) (t->sig3[n3] - p->sig53tab[16 + d3] + p->sig53tab[32 + 16 * d5 + d3]);
    return pen + sig;
}

static int gap_ext(const so_params* p, int i) { return i > p->codonk1 ? p->lgep : p->gep; }

int so_trcbk_ng(const so_params* p, const so_task* t, int32_t* score, int32_t* skl, int cap)
{
    const int width = t->up - t->lw + 3;
    *score = NEVSEL32;
    if (width < 0) return 0;
    if (p->spj && (!p->penalty || !p->sig53tab || !t->int53 || t->b_right - t->b_left >= p->n_penalty))
        return -3;          /* no tables, or an intron could be longer than the penalty table */
    const int dagp = p->noll == 3, spj = p->spj;
    const int nod = 2 * p->noll - 1;
    const int gop_k[3] = { 0, p->gop, p->lgop };            /* PwdB::GOP, src/aln2.cc:111 */
    const int LocalL = p->local && t->a_exgl && t->b_exgl;
    const int LocalR = p->local && t->a_exgr && t->b_exgr;
    const int a_left = t->a_left, a_right = t->a_right, b_left = t->b_left, b_right = t->b_right;
    const int lw = t->lw, up = t->up;

    ng_rvp* buf = (ng_rvp*) malloc((size_t) 3 * width * sizeof(ng_rvp));
    unsigned char* dbuf = (unsigned char*) calloc((size_t) width, 1);
    ng_vmf vmf = { 0, 0, 0, 0 };
    if (!buf || !dbuf) { free(buf); free(dbuf); return -1; }
    for (int i = 0; i < 3 * width; ++i) { buf[i].val = NEVSEL32; buf[i].ptr = 0; }
    /* band rows indexed by diagonal r = n - m in [lw - 1, up + 1] */
    ng_rvp* H = buf - lw + 1;
    ng_rvp* F = H + width;
    ng_rvp* F2 = F + width;
    unsigned char* dirs = dbuf - lw + 1;
    vmf_add(&vmf, 0, 0, 0);                                 /* record 0 is never a path node */

    /* ---- initS_ng ---- */
    {
        int r = b_left - a_left, rr = b_right - a_left;
        H[r].val = 0; dirs[r] = 0;
        H[r].ptr = vmf_add(&vmf, a_left, b_left, 0);
        if (t->a_exgl) {
            if (up < rr) rr = up;
            while (++r <= rr) { H[r].val = 0; H[r].ptr = 0; dirs[r] = 1; }
        }
        r = b_left - a_left;
        rr = b_left - a_right;
        if (lw > rr) rr = lw;
        for (int i = 1; --r >= rr; ++i) {
            dirs[r] = 2;
            if (t->b_exgl) { H[r].val = 0; H[r].ptr = 0; }
            else {
                H[r] = H[r + 1];
                H[r].val += i == 1 ? p->gappen1 : gap_ext(p, i);
            }
        }
    }

    int best_val = NEVSEL32, best_m = a_left, best_n = b_left, best_p = 0;   /* LocalR */
    int m = a_left;
    if (!t->a_exgl) --m;
    for (++m; m <= a_right; ++m) {
        const int internal = spj && (!t->a_exgr || m < a_right);
        const int sigB = t->cip ? t->cip[m] : 0;    /* src/fwd2s1.cc:254 */
        int n = (m - 1) + lw > b_left ? (m - 1) + lw : b_left;
        const int n9 = (m - 1) + up + 1 < b_right ? (m - 1) + up + 1 : b_right;
        const int32_t* qprof = p->simmtx + (size_t) t->a[m > 0 ? m - 1 : 0] * p->simdim;
        ng_rvp e1 = { NEVSEL32, 0 }, e2 = { NEVSEL32, 0 };
        ng_cand rcd[NG_NCAND + 1];
        int idx[NG_NCAND + 1];
        for (int l = 0; l <= NG_NCAND; ++l) {
            rcd[l].val = NEVSEL32; rcd[l].ptr = rcd[l].dir = rcd[l].jnc = 0;
            idx[l] = l;
        }
        int ncand = -1;
        int psp = 0;
        while (++n <= n9) {
            const int r = n - m;
            ng_rvp* hf[5] = { &H[r], &e1, &F[r], &e2, dagp ? &F2[r] : 0 };
            ng_rvp* h = hf[0];
            ng_rvp* mx = h;
            const int diag = h->val;
            if (m != a_left) {
                h->val += qprof[t->b[n - 1]];
                dirs[r] = (dirs[r] % NG_NEWD) ? NG_NEWD : 0;
                /* vertical */
                const ng_rvp* from = &H[r + 1];
                int x = from->val + p->gop;
                if (x >= F[r + 1].val) { F[r].val = x; F[r].ptr = from->ptr; }
                else F[r] = F[r + 1];
                F[r].val += p->gep;
                if (F[r].val > mx->val) mx = &F[r];
                if (dagp) {
                    x = from->val + p->lgop;
                    if (x >= F2[r + 1].val) { F2[r].val = x; F2[r].ptr = from->ptr; }
                    else F2[r] = F2[r + 1];
                    F2[r].val += p->lgep;
                    if (F2[r].val > mx->val) mx = &F2[r];
                }
            }
            /* horizontal */
            {
                int x = H[r - 1].val + p->gop;
                const int prev_psp = psp;
                if (x >= e1.val) { e1.val = x; e1.ptr = H[r - 1].ptr; psp = psp ? 1 : 0; }
                else psp &= 1;
                e1.val += p->gep;
                if (e1.val >= mx->val) mx = &e1;
                if (dagp) {
                    x = H[r - 1].val + p->lgop;
                    if (x >= e2.val) { e2.val = x; e2.ptr = H[r - 1].ptr; if (prev_psp) psp |= 2; }
                    else psp |= prev_psp & 2;
                    e2.val += p->lgep;
                    if (e2.val >= mx->val) mx = &e2;
                }
            }
            /* acceptor: every stored donor of this row, per gap state */
            if (internal && cano3(t, n)) {
                const ng_cand* top[5] = { 0, 0, 0, 0, 0 };
                for (int l = 0; l <= ncand; ++l) {
                    const ng_cand* c = rcd + idx[l];
                    if (n - c->jnc < p->llmt) continue;
                    const int x = c->val + sigB + spjscr(p, t, c->jnc, n);
                    ng_rvp* to = hf[c->dir];
                    if (x >= to->val) { to->val = x; top[c->dir] = c; }
                }
                for (int k = 0; k < nod; ++k) {
                    const ng_cand* c = top[k];
                    if (!c) continue;
                    psp |= ng_psp_bit[k];
                    const int inner = vmf_add(&vmf, m, c->jnc, c->ptr);
                    hf[k]->ptr = vmf_add(&vmf, m, n, inner);
                    if (hf[k]->val >= mx->val) mx = hf[k];
                }
            }
            /* best state */
            int hd = 0;
            if (h != mx) {
                *h = *mx;
                while (mx != hf[++hd]) ;
                dirs[r] = (unsigned char) hd;
            } else if (p->local && h->val > diag) {
                if (LocalL && diag == 0) h->ptr = vmf_add(&vmf, m - 1, n - 1, 0);
                else if (LocalR && h->val > best_val) {
                    best_val = h->val; best_p = h->ptr; best_m = m; best_n = n;
                }
            }
            if (LocalL && h->val <= 0) { h->val = 0; dirs[r] = 1; }
            else if (dirs[r] == NG_NEWD && !(psp & ng_psp_bit[0]))
                h->ptr = vmf_add(&vmf, m - 1, n - 1, h->ptr);
            /* donor: keep the NCAND best (value + 5' signal) of this row */
            if (internal && cano5(t, n)) {
                const int sigJ = t->sig5[n];
                for (int k = hd == 0 ? 0 : 1; k < nod; ++k) {
                    const ng_rvp* from = hf[k];
                    if (psp & ng_psp_bit[k]) continue;
                    if (k != hd) {
                        int z = mx->val;
                        if (hd == 0 || (k - hd) % 2) z += gop_k[k / 2];
                        if (from->val <= z) continue;
                    }
                    const int x = from->val + sigJ;
                    int l = ncand < NG_NCAND ? ++ncand : NG_NCAND;
                    while (--l >= 0) {
                        if (x > rcd[idx[l]].val) { int s = idx[l]; idx[l] = idx[l + 1]; idx[l + 1] = s; }
                        else break;
                    }
                    if (++l < NG_NCAND) {
                        ng_cand* c = rcd + idx[l];
                        c->val = x; c->jnc = n; c->dir = k; c->ptr = from->ptr;
                    } else --ncand;
                }
            }
        }
    }

    int ptr, val;
    if (LocalR) {
        ptr = vmf_add(&vmf, best_m, best_n, best_p);
        val = best_val;
    } else {
        /* lastS_ng */
        int rw = lw > b_left - a_right ? lw : b_left - a_right;
        const int r9 = b_right - a_right;
        int mxr = r9;
        if (t->a_exgr)
            for (int r = rw; r <= r9; ++r) if (H[r].val > H[mxr].val) mxr = r;
        if (t->b_exgr) {
            rw = up < b_right - a_left ? up : b_right - a_left;
            for (int r = rw; r > r9; --r) if (H[r].val > H[mxr].val) mxr = r;
        }
        const int i = mxr - r9;
        int m9 = a_right, n9 = b_right;
        if (i > 0) m9 -= i;
        if (i < 0) n9 += i;
        H[mxr].ptr = vmf_add(&vmf, m9, n9, H[mxr].ptr);
        val = H[mxr].val;
        ptr = H[mxr].ptr;
    }

    /* trcbkalignS_ng: Vmf::traceback, corners end -> start, then the start-point adjustment */
    int cnt = 0;
    if (vmf.fail) { cnt = -1; }
    else if (ptr) {
        int m_last = 0, n_last = 0;
        for (int q = ptr; ; q = vmf.rec[q].p) {
            m_last = vmf.rec[q].m; n_last = vmf.rec[q].n;
            if (cnt < cap) { skl[2 * cnt] = m_last; skl[2 * cnt + 1] = n_last; }
            ++cnt;
            if (!vmf.rec[q].p) break;
        }
        const int rd = p->local ? 0 : (n_last - m_last) - b_left + a_left;
        if (rd) {
            const int mm = rd > 0 ? a_left : a_left - rd, nn = rd > 0 ? b_left + rd : b_left;
            if (cnt < cap) { skl[2 * cnt] = mm; skl[2 * cnt + 1] = nn; }
            ++cnt;
        }
    }
    *score = val;
    free(buf); free(dbuf); free(vmf.rec);
    return cnt;
}

/* Aln2s1::scorealoneS_ng (src/fwd2s1.cc:1163-1336) with sinitS_ng / slastS_ng (1112-1161): the
 * scalar score-only kernel (HomScoreS_ng under -A0 and for queries shorter than 4 residues,
 * src/fwd2s1.cc:2704-2705).  Same recurrences as forwardS_ng without path records, but with its
 * own tie rules (strict comparisons where forwardS_ng accepts ties, ties accepted in the donor
 * list) and its own initial rows.  Returns 0, -1 allocation failure, -3 missing tables. */
int so_scorealone_ng(const so_params* p, const so_task* t, int32_t* score)
{
    const int width = t->up - t->lw + 3;
    *score = NEVSEL32;
    if (width < 3) return 0;
    if (!p->penalty || !p->sig53tab || !t->int53 || t->b_right - t->b_left >= p->n_penalty) return -3;
    const int dagp = p->noll == 3;
    const int nod = 2 * p->noll - 1;
    const int gop_k[3] = { 0, p->gop, p->lgop };
    const int LocalL = p->local && t->a_exgl && t->b_exgl;
    const int LocalR = p->local && t->a_exgr && t->b_exgr;
    const int a_left = t->a_left, a_right = t->a_right, b_left = t->b_left, b_right = t->b_right;
    const int lw = t->lw, up = t->up;
    int* buf = (int*) malloc((size_t) 3 * width * sizeof(int));
    if (!buf) return -1;
    for (int i = 0; i < 3 * width; ++i) buf[i] = NEVSEL32;
    int* H = buf - lw + 1;
    int* F = H + width;
    int* F2 = F + width;
    /* ---- sinitS_ng ---- */
    {
        int r = b_left - a_left, rr = b_right - a_left;
        H[r] = 0;
        if (t->a_exgl) {
            if (up < rr) rr = up;
            for (int q = r + 1; q <= rr; ++q) H[q] = 0;
        }
        rr = b_left - a_right;
        if (lw > rr) rr = lw;
        if (t->b_exgl) {
            for (int q = rr; q < r; ++q) H[q] = 0;
        } else {
            for (int i = 1; --r >= rr; ++i) {
                H[r] = H[r + 1];
                if (i == 1) { H[r] += p->gappen1; F[r] = H[r]; }
                else { F[r] = F[r + 1]; H[r] += gap_ext(p, i); F[r] += p->gep; }
            }
        }
    }
    int maxh = NEVSEL32;
    int m = a_left;
    if (!t->a_exgl) --m;
    for (++m; m <= a_right; ++m) {
        int n = (m - 1) + lw > b_left ? (m - 1) + lw : b_left;
        const int n9 = (m - 1) + up + 1 < b_right ? (m - 1) + up + 1 : b_right;
        const int32_t* qprof = p->simmtx + (size_t) t->a[m > 0 ? m - 1 : 0] * p->simdim;
        int e1 = NEVSEL32, e2 = NEVSEL32;
        struct { int val, dir, jnc; } rcd[NG_NCAND + 1];
        int idx[NG_NCAND + 1];
        for (int l = 0; l <= NG_NCAND; ++l) { rcd[l].val = NEVSEL32; rcd[l].dir = rcd[l].jnc = 0; idx[l] = l; }
        int ncand = -1, psp = 0;
        const int sigB = t->cip ? t->cip[m] : 0;    /* src/fwd2s1.cc:1191 */
        while (++n <= n9) {
            const int r = n - m;
            int black = NEVSEL32;
            int* hf[5] = { &H[r], &e1, &F[r], &e2, dagp ? &F2[r] : &black };
            int* h = hf[0];
            int* mx = h;
            if (m != a_left) {
                *h += qprof[t->b[n - 1]];
                int x = H[r + 1] + p->gop;
                F[r] = (x > F[r + 1] ? x : F[r + 1]) + p->gep;
                if (F[r] > *mx) mx = &F[r];
                if (dagp) {
                    x = H[r + 1] + p->lgop;
                    F2[r] = (x > F2[r + 1] ? x : F2[r + 1]) + p->lgep;
                    if (F2[r] > *mx) mx = &F2[r];
                }
            }
            {
                int x = H[r - 1] + p->gop;
                const int prev_psp = psp;
                if (x > e1) { e1 = x; psp = psp ? 1 : 0; }
                else psp &= 1;
                e1 += p->gep;
                if (e1 > *mx) mx = &e1;
                if (dagp) {
                    x = H[r - 1] + p->lgop;
                    if (x > e2) { e2 = x; if (prev_psp) psp |= 2; }
                    else psp |= prev_psp & 2;
                    e2 += p->lgep;
                    if (e2 > *mx) mx = &e2;
                }
            }
            if (cano3(t, n)) {
                int top[5] = { 0, 0, 0, 0, 0 };
                for (int l = 0; l <= ncand; ++l) {
                    const int j = idx[l];
                    if (n - rcd[j].jnc < p->llmt) continue;
                    const int x = rcd[j].val + sigB + spjscr(p, t, rcd[j].jnc, n);
                    if (x > *hf[rcd[j].dir]) { *hf[rcd[j].dir] = x; top[rcd[j].dir] = 1; }
                }
                for (int k = 0; k < nod; ++k) {
                    if (!top[k]) continue;
                    psp |= ng_psp_bit[k];
                    if (*hf[k] > *mx) mx = hf[k];
                }
            }
            const int y = *h;
            if (h != mx) *h = *mx;
            else if (LocalR && y > maxh) maxh = y;
            if (LocalL && *h < 0) *h = 0;
            int hd = 0;
            while (mx != hf[hd]) ++hd;
            if (cano5(t, n)) {
                const int sigJ = t->sig5[n];
                for (int k = hd == 0 ? 0 : 1; k < nod; ++k) {
                    if (psp & ng_psp_bit[k]) continue;
                    if (k != hd) {
                        int z = *mx;
                        if (hd == 0 || (k - hd) % 2) z += gop_k[k / 2];
                        if (*hf[k] <= z) continue;
                    }
                    const int x = *hf[k] + sigJ;
                    int l = ncand < NG_NCAND ? ++ncand : NG_NCAND;
                    while (--l >= 0) {
                        if (x >= rcd[idx[l]].val) { int s = idx[l]; idx[l] = idx[l + 1]; idx[l + 1] = s; }
                        else break;
                    }
                    if (++l < NG_NCAND) { rcd[idx[l]].val = x; rcd[idx[l]].jnc = n; rcd[idx[l]].dir = k; }
                    else --ncand;
                }
            }
        }
    }
    if (!LocalR) {
        /* slastS_ng */
        const int r9 = b_right - a_right;
        int mxv = H[r9];
        if (t->b_exgr) {
            const int rw = up < b_right - a_left ? up : b_right - a_left;
            for (int r = rw; r > r9; --r) if (H[r] > mxv) mxv = H[r];
        }
        if (t->a_exgr) {
            const int rw = lw > b_left - a_right ? lw : b_left - a_right;
            for (int r = rw; r < r9; ++r) if (H[r] > mxv) mxv = H[r];
        }
        maxh = mxv;
    }
    *score = maxh;
    free(buf);
    return 0;
}
