// gspaln_hng.cuh -- the scalar protein x genome spliced DP kernel on the device.
//
// Reference: Aln2h1::trcbkalignH_ng on its scalar branch (src/fwd2h1.cc:1997-2041): forwardH_ng
// (294-617) with initH_ng / lastH_ng (143-292), the Vmf record store and walk (src/vmf.cc:66-140),
// exact intron scoring SpJunc::spjscr and the split-codon translation SpJunc::spjseq
// (src/codepot.cc:74-102).  The reference takes this branch for every block with fewer than 8
// query rows (src/fwd2h1.cc:2007), which the protein driver meets between the intermediate rows
// of a Hirschberg pass.  Exactness kernel, not a throughput kernel: one thread per problem, band
// rows ({value, record, direction} per diagonal for H, F, F2) and the record store in a per-thread
// HBM workspace, raw inputs (residues, SGPT6 records, INT53) copied per problem with a margin.
#pragma once
#include "gspaln_kernels.cuh"

namespace gspaln {

constexpr int HNG_THREADS = 32;                     // threads per CTA (one problem each)
constexpr int HNG_NEVSEL = INT_MIN / 16 * 7;        // NEVSEL, src/cmn.h:79

constexpr int HNG_NCAND = 4, HNG_NQUE = 3;      // NCAND (src/aln.h:55), NQUE (src/fwd2h1.cc:43)
enum { DEAD, RSRV, DIAG, NEWD, VERT, SLA1, SLA2, VERL, HORI, HOR1, HOR2, HORL, NEWV, NEWH, SPIN = 16 };   // src/aln.h:30-35
__constant__ int c_dir2nod[16] = {-1, -1, 0, 0, 2, 2, 2, 4, 1, 1, 1, 3, 2, 1, -1, -1};
__constant__ int c_nod2dir[5] = {DIAG, HORI, VERT, HORL, VERL};
__constant__ int c_is_diag[16] = {0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
__constant__ int c_is_vert[16] = {0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 1, 0, 0, 0};
__constant__ int c_is_hori[16] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 0, 1, 0, 0};
__constant__ int c_next_p[3] = {1, 2, 0};
__constant__ unsigned char c_hncred[17] = {15, 15, 0, 1, 4, 2, 5, 6, 10, 3, 7, 8, 10, 9, 12, 13, 14};

struct HRvpd { int val, ptr, dir; };
struct HCand { int val, ptr, dir, jnc; };

struct DevNgHParams {           // frozen scalars + device pointers of the tables
    int gop, gep, lgop, lgep, codonk1, gw1, gw2, gw3, gw3l, gape1, gape2, extragop;
    int local, spj, noll, minl, simdim, n_penalty;
    const int* mtx;             // simmtx[aa][tron], simdim x simdim
    const short* penalty;       // IntronPenalty::Penalty(len)
    const short* sig53tab;      // 544 shorts
    const unsigned char* spj_tabs;  // spj_tron_tab | spj_amb_tron_tab | spj_tron_amb_tab | aa2nuc
};

struct DevNgHTask {
    int a_left, a_right, b_left, b_right, lw, up;
    int a_exgl, a_exgr, b_exgl, b_exgr;
    int skl_cap, rec_cap;
    int a_lo, b_lo;             // first copied position of the query / genome arrays
    long long a_off, b_off;     // into the byte pools (element 0 == position a_lo / b_lo)
    long long sg_off;           // SGPT6 shorts (8 per column), int53: column b_lo first
    long long skl_off, work_off;    // corners (int2), workspace bytes
};

struct HngIn {                  // accessors with the reference's indexing
    const unsigned char* a; const unsigned char* b; const short* sg; const unsigned short* i53;
    int a_lo, b_lo, b_left, b_right;
};

struct HngVmf {
    int* rec; int cap, n; bool overflow;
    __device__ __forceinline__ int add(int m, int n_, int p)
    {
        if (n >= cap) { overflow = true; return 0; }
        int* r = rec + 3 * (long long) n;
        r[0] = m; r[1] = n_; r[2] = p;
        return n++;
    }
};

#define A (T.a - T.a_lo)
#define B (T.b - T.b_lo)
#define I53 (T.i53 - T.b_lo)
#define SGD(n, f) (T.sg[8 * ((long long) (n) - T.b_lo) + (f)])
enum { F_SIG5, F_SIG3, F_SIGS, F_SIGT, F_SIGE, F_SIGI, F_PHS5, F_PHS3 };

__device__ __forceinline__ int gap_ext3(const DevNgHParams& P, int i) { return i > P.codonk1 ? P.lgep : P.gep; }

// SpJunc::spjseq (src/codepot.cc:79-102): the two residues a split codon translates to
__device__ const unsigned char* spjseq(const DevNgHParams& P, const HngIn& T, int n5, int n3)
{
    const unsigned char* tab = P.spj_tabs, *amb_tron = tab + 514, *tron_amb = tab + 514 + 128, *aa2nuc = tab + 514 + 256;
    if (n5 < T.b_left || n3 >= T.b_right) return tab + 2 * 256;
    const unsigned char* b5 = &B[n5 - 2];
    const unsigned char* b3 = &B[n3];               // n3 > 0 here (n3 >= minl)
    auto NC = [&](unsigned c) -> int { const unsigned v = aa2nuc[c < 26 ? c : 0]; return c_hncred[v < 17 ? v : 0]; };
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
    if (amb == 0 || w == 256) return tab + 2 * w;
    if (amb == 1) return amb_tron + 2 * w;
    return tron_amb + 2 * w;
}

__device__ __forceinline__ int spjscr_h(const DevNgHParams& P, const HngIn& T, int n5, int n3)
{
    const int len = n3 - n5;
    const int pen = P.penalty[len < 0 ? 0 : (len < P.n_penalty ? len : P.n_penalty - 1)];
    const int d5 = I53[n5] & 15, d3 = (I53[n3] >> 4) & 15;
    const short sig = (short) (SGD(n3, F_SIG3) - P.sig53tab[16 + d3] + P.sig53tab[32 + 16 * d5 + d3]);
    return pen + sig;
}

__global__ void __launch_bounds__(HNG_THREADS)
dp_hng_kernel(const DevNgHParams* __restrict__ gP, const DevNgHTask* __restrict__ tasks, int ntasks, int* ticket,
              const unsigned char* __restrict__ apool, const unsigned char* __restrict__ bpool,
              const short* __restrict__ sgpool, const unsigned short* __restrict__ i53pool,
              unsigned char* workpool, int2* sklpool, DevResult* results)
{
    __shared__ DevNgHParams P;
    if (threadIdx.x < sizeof(DevNgHParams) / 4)
        reinterpret_cast<int*>(&P)[threadIdx.x] = reinterpret_cast<const int*>(gP)[threadIdx.x];
    __syncthreads();
    for (;;) {
        const int ti = atomicAdd(ticket, 1);
        if (ti >= ntasks) break;
        const DevNgHTask t = tasks[ti];
        HngIn T;
        T.a = apool + t.a_off; T.b = bpool + t.b_off; T.sg = sgpool + 8 * t.sg_off; T.i53 = i53pool + t.sg_off;
        T.a_lo = t.a_lo; T.b_lo = t.b_lo; T.b_left = t.b_left; T.b_right = t.b_right;
        struct { int a_exgl, a_exgr, b_exgl, b_exgr; } TF = {t.a_exgl, t.a_exgr, t.b_exgl, t.b_exgr};
        const int width = t.up - t.lw + 7;
        const int noll = P.noll, nod = 2 * noll - 1;
        const bool dagp = noll == 3;
        const int Local = P.local;
        const int LocalL = Local && t.a_exgl && t.b_exgl, LocalR = Local && t.a_exgr && t.b_exgr;
        const int a_left = t.a_left, a_right = t.a_right, b_left = t.b_left, b_right = t.b_right;
        const int lw = t.lw, up = t.up;
        const int spj = P.spj;
        const int gop_k[3] = {0, P.gop, P.lgop};
        const int GapE1 = P.gape1, GapE2 = P.gape2, GapW1 = P.gw1, GapW2 = P.gw2, GapW3 = P.gw3, GapW3L = P.gw3l;
        const HRvpd black = {HNG_NEVSEL, 0, 0};
        HRvpd* buf = reinterpret_cast<HRvpd*>(workpool + t.work_off);
        HngVmf W;
        W.rec = reinterpret_cast<int*>(buf + 3 * (width + 8));
        W.cap = t.rec_cap; W.n = 0; W.overflow = false;
        for (int i = 0; i < 3 * (width + 8); ++i) buf[i] = black;
        HRvpd* hh[3];
        hh[0] = buf - lw + 3;
        hh[1] = hh[0] + width;
        hh[2] = hh[1] + width;
        W.add(0, 0, 0);
        /* ---- initH_ng ---- */
        {
            int n = b_left, r = b_left - 3 * a_left, rr = b_right - 3 * a_left;
            const int dir = TF.a_exgl ? DEAD : DIAG;
            int jnc[3] = { n, 0, 0 };
            int bbn = n + 1;
            HRvpd* h = hh[0] + r;
            h->val = (TF.a_exgl && SGD(bbn, F_SIGS) > 0) ? SGD(bbn, F_SIGS) : 0;
            h->dir = dir;
            h->ptr = W.add(a_left, n, 0);
            if (TF.a_exgl) {
                if (up < rr) rr = up;
                for (int i = 1; ++r <= rr; ++i) {
                    ++h; ++bbn; ++n;
                    if (i < 3) {
                        h->val = SGD(bbn, F_SIGS) > 0 ? SGD(bbn, F_SIGS) : 0;
                        h->dir = dir;
                        h->ptr = W.add(a_left, n, 0);
                        jnc[i] = n;
                    } else {
                        *h = h[-3];
                        const int k = n - jnc[i % 3];
                        if (k == 3 && !(TF.a_exgl & 1)) h->val += P.gop;
                        if (!(TF.a_exgl & 2)) h->val += gap_ext3(P, k);
                        h->val += SGD(bbn - 3, F_SIGE);
                        h->dir = HORI;
                        int xx = h[-1].val + GapW1;
                        if (xx > h->val) { *h = h[-1]; h->val = xx; h->dir = HOR1; }
                        xx = h[-2].val + GapW2;
                        if (xx > h->val) { *h = h[-2]; h->val = xx; h->dir = HOR2; }
                    }
                    const int xs = SGD(bbn, F_SIGS) > 0 ? SGD(bbn, F_SIGS) : 0;
                    if (h->val < xs) {
                        h->val = xs;
                        h->dir = DEAD;
                        h->ptr = W.add(a_left, n, 0);
                        jnc[i % 3] = n;
                    }
                }
            }
            r = b_left - 3 * a_left;
            rr = b_left - 3 * a_right;
            h = hh[0] + r - 1;
            if (lw > rr) rr = lw;
            for (int i = 1; --r >= rr; ++i, --h) {
                if (TF.b_exgl == 1) { h->val = 0; h->dir = DEAD; h->ptr = 0; }
                else if (i <= 3) {
                    *h = h[i];
                    if (!(TF.b_exgl & 2)) h->val += P.gep;
                    if (!(TF.b_exgl & 1)) h->val += P.gop;
                    if (i < 3) h->val += P.extragop;
                    h->dir = VERT;
                } else {
                    *h = h[3];
                    if (!(TF.b_exgl & 2)) h->val += gap_ext3(P, i);
                }
            }
        }

        int best_val = HNG_NEVSEL, best_m = a_left, best_n = b_left, best_p = 0;
        int m = a_left;
        if (!TF.a_exgl) --m;
        int n1 = 3 * m + lw - 1, n2 = 3 * m + up;
        for (++m; m <= a_right; ++m) {
            n1 += 3; n2 += 3;
            const int n0 = n1 > b_left ? n1 : b_left;
            const int n9 = n2 < b_right ? n2 : b_right;
            int n = n0;
            int r = n - 3 * m;
            HRvpd e1[2 * HNG_NQUE];
            HRvpd* e2 = e1 + HNG_NQUE;
            for (int i = 0; i < 2 * HNG_NQUE; ++i) e1[i] = black;
            if (!TF.b_exgl && m == a_left) {
                e1[2] = e2[2] = hh[0][r];
                e1[2].val = GapW3;
                e2[2].val = GapW3L;
            }
            const int* qprof0 = P.mtx + A[m > 0 ? m - 1 : 0] * P.simdim;
            const int* qprof1 = P.mtx + A[m] * P.simdim;
            HCand hl[3][HNG_NCAND + 1];
            int nx[3][HNG_NCAND + 1];
            for (int ph = 0; ph < 3; ++ph)
                for (int l = 0; l <= HNG_NCAND; ++l) {
                    hl[ph][l].val = HNG_NEVSEL; hl[ph][l].ptr = hl[ph][l].dir = hl[ph][l].jnc = 0;
                    nx[ph][l] = l;
                }
            int ncand[3] = { -1, -1, -1 };
            HRvpd* h = hh[0] + r;
            HRvpd* f = hh[1] + r;
            HRvpd* f2 = dagp ? hh[2] + r : 0;
            HRvpd* hf[5];
            for (int q = 0; n <= n9; ++n, ++h, ++f) {
                const int bs = n - 2;                           /* b->at(n - 2) */
                const int sigE = n > b_left ? SGD(n - 2, F_SIGE) : 0;
                HRvpd* eq1 = e1 + q;
                HRvpd* eq2 = dagp ? e2 + q : 0;
                hf[0] = h; hf[1] = eq1; hf[2] = f; hf[3] = eq2; hf[4] = f2;
                HRvpd hq = *h;
                HRvpd* from = h;
                HRvpd* mx = h;
                int xv, yv;
                if (m != a_left) {
                    if (n < b_left + 3) *h = black;
                    else {
                        h->val += qprof0[B[bs]] + sigE;
                        h->dir = c_is_diag[from->dir & 15] ? DIAG : NEWD;
                    }
                    yv = f[3].val + P.gep;
                    ++from;
                    xv = from->val + (c_is_vert[from->dir & 15] ? GapE1 : GapW1);
                    if (xv > yv) { f->val = xv; f->dir = SLA2; f->ptr = from->ptr; }
                    else f->val = yv;
                    ++from;
                    xv = from->val + (c_is_vert[from->dir & 15] ? GapE2 : GapW2);
                    if (xv > f->val) { f->val = xv; f->dir = SLA1; f->ptr = from->ptr; }
                    ++from;
                    xv = from->val + GapW3;
                    if (xv >= f->val) { f->val = xv; f->dir = VERT; f->ptr = from->ptr; }
                    else if (yv >= f->val) { f->val = yv; f->dir = VERT; f->ptr = f[3].ptr; }
                    if (f->val > mx->val) mx = f;
                    if (dagp) {
                        xv = from->val + GapW3L;
                        yv = f2[3].val + P.lgep;
                        if (xv >= yv) { f2->val = xv; f2->dir = VERL; f2->ptr = from->ptr; }
                        else { *f2 = f2[3]; f2->val = yv; }
                        if (f2->val > mx->val) mx = f2;
                    }
                }
                /* horizontal moves */
                if (n > n0 + 2) {
                    from = h - 3;
                    xv = from->val + GapW3;
                    yv = eq1->val += P.gep;
                    if (xv > yv) { *eq1 = *from; eq1->val = xv; }
                    eq1->val += sigE;
                    eq1->dir = (eq1->dir & SPIN) + HORI;
                    if (dagp) {
                        xv = from->val + GapW3L;
                        yv = eq2->val += P.lgep;
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
                if (++q == HNG_NQUE) q = 0;

                /* intron 3' boundary */
                const int phs3 = SGD(n, F_PHS3);
                if (spj && phs3 > -2) {
                    int phs = phs3 == 2 ? -1 : phs3;
                    for (;;) {
                        const int nb = n - phs;
                        const int* pnx = nx[phs + 1];
                        const HCand* top[5] = { 0, 0, 0, 0, 0 };
                        for (int l = 0; l <= ncand[phs + 1]; ++l) {
                            const HCand* phl = hl[phs + 1] + pnx[l];
                            if (phs == 1 && phl->dir == 2) continue;
                            if (nb - phl->jnc < P.minl) continue;
                            xv = phl->val + spjscr_h(P, T, phl->jnc, nb);
                            if (phl->dir == 0 && phs) {
                                const unsigned char* cs = spjseq(P, T, phl->jnc, nb);
                                if (phs == 1) xv += qprof0[cs[0]];
                                else xv += qprof1[cs[1]] - qprof1[B[bs + 3]] - SGD(n + 1, F_SIGE);
                            }
                            from = hf[phl->dir];
                            if (xv > from->val) { from->val = xv; top[phl->dir] = phl; }
                        }
                        for (int d = 0; d < nod; ++d) {
                            const HCand* phl = top[d];
                            if (!phl) continue;
                            from = hf[d];
                            from->ptr = W.add(m, n, W.add(m, phl->jnc + phs, phl->ptr));
                            from->dir = c_nod2dir[phl->dir] | SPIN;
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
                    if (LocalL && hq.dir == 0 && !(h->dir & SPIN)) h->ptr = W.add(m - 1, n - 3, 0);
                    else if (LocalR && yv > best_val) { best_val = yv; best_p = h->ptr; best_m = m; best_n = n; }
                }
                if (LocalL && h->val <= 0) h->val = h->dir = 0;
                else if (h->dir == NEWD) h->ptr = W.add(m - 1, n - 3, h->ptr);

                /* intron 5' boundary */
                const int phs5 = SGD(n, F_PHS5);
                if (spj && phs5 > -2) {
                    int phs = phs5 == 2 ? -1 : phs5;
                    for (;;) {
                        const int nb = n - phs;
                        const int sigJ = SGD(nb, F_SIG5);
                        const int hd = c_dir2nod[mx->dir & 15];
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
                            HCand* phl = hl[phs + 1];
                            int* pnx = nx[phs + 1];
                            int* nc = &ncand[phs + 1];
                            int l = *nc < HNG_NCAND ? ++*nc : HNG_NCAND;
                            while (--l >= 0) {
                                if (xv >= phl[pnx[l]].val) { int s = pnx[l]; pnx[l] = pnx[l + 1]; pnx[l + 1] = s; }
                                else break;
                            }
                            if (++l < HNG_NCAND) {
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
            HRvpd* h = hh[0] + rw;
            HRvpd* h9 = hh[0] + b_right - m3;
            HRvpd* mx = h9;
            int bbn = rw + m3;
            if (TF.a_exgr) {
                for (int ph = 0; h <= h9; ++h, ++bbn, ++rf, ph = c_next_p[ph]) {
                    glen[ph] += 3;
                    int cand[3] = { h->val, HNG_NEVSEL, HNG_NEVSEL };
                    if (rf - rw >= 3 && h[-3].dir != DEAD) {
                        cand[1] = h[-3].val + SGD(bbn - 2, F_SIGE);
                        if (!(TF.a_exgr & 2)) cand[1] += gap_ext3(P, glen[ph]);
                        if (!(TF.a_exgr & 1) && glen[ph] == 3) cand[1] += P.gop;
                        if (SGD(bbn - 2, F_SIGT) > 0 && !(h->dir & SPIN)) cand[2] = h[-3].val + SGD(bbn - 2, F_SIGT);
                    }
                    const int sig5 = (Local && SGD(bbn, F_SIG5) > 0) ? SGD(bbn, F_SIG5) : 0;
                    cand[0] += sig5;
                    cand[1] += sig5;
                    int k = 0;
                    if (cand[1] > cand[k]) k = 1;
                    if (cand[2] > cand[k]) k = 2;
                    if (k == 0) { if (!c_is_hori[h->dir & 15]) glen[ph] = 0; }
                    else if (k == 1) { *h = h[-3]; h->dir = HORI; h->val = cand[k] - sig5; }
                    else {
                        *h = h[-3];
                        h->dir = DEAD;
                        h->val = cand[k];
                        if (h->val > mx->val) h->ptr = W.add(a_right, rf + m3 - 3, h->ptr);
                    }
                    if (h->val > mx->val) mx = h;
                }
            } else {
                bbn += (int) (h9 - h);
                const int yv = h9[-3].val + SGD(bbn - 2, F_SIGT);
                if (yv > h9->val) { *h9 = h9[-3]; h9->val = yv; h9->dir = HORI; }
            }
            int done = 0;
            if (TF.b_exgr == 1) {
                rw = up < b_right - 3 * a_left ? up : b_right - 3 * a_left;
                int g[3] = { HNG_NEVSEL, HNG_NEVSEL, HNG_NEVSEL };
                h = hh[0] + rw - 3;
                for (int ph = 0; h >= h9; --h) {
                    int xv = h[3].val;
                    if (!(TF.b_exgr & 1)) xv += P.gop;
                    if (xv > g[ph]) g[ph] = xv;
                    if (!(TF.b_exgr & 2)) g[ph] += P.gep;
                    if (h->val > g[ph]) g[ph] = HNG_NEVSEL;
                    else if (g[ph] > mx->val) { mx = h; mx->val = g[ph]; }
                    if (++ph == 3) ph = 0;
                }
            } else if (TF.b_exgr == 2) {
                mx = hh[1] + b_right - m3;
                mx->ptr = W.add(a_right, b_right, mx->ptr);
                done = 1;
            }
            if (!done) {
                int pp = (int) (mx - h9);
                int m9 = a_right, n9 = b_right;
                if (pp > 0) { m9 -= (pp + 2) / 3; if (pp %= 3) n9 -= 3 - pp; }
                else if (pp < 0) n9 += pp;
                mx->ptr = W.add(m9, n9, mx->ptr);
            }
            val = mx->val;
            ptr = mx->ptr;
        } else {
            ptr = W.add(best_m, best_n, best_p);
            val = best_val;
        }


        int2* skl = sklpool + t.skl_off;
        int cnt = 0;
        if (!W.overflow && ptr) {
            int m_last = 0, n_last = 0;
            for (int q = ptr; ; ) {
                const int* rr = W.rec + 3 * (long long) q;
                m_last = rr[0]; n_last = rr[1];
                if (cnt < t.skl_cap) skl[cnt] = make_int2(m_last, n_last);
                ++cnt;
                q = rr[2];
                if (!q) break;
            }
            const int rd = Local ? 0 : (n_last - 3 * m_last) - b_left + 3 * a_left;
            if (rd) {
                if (cnt < t.skl_cap)
                    skl[cnt] = rd > 0 ? make_int2(a_left, b_left + rd) : make_int2(a_left - rd / 3, b_left);
                ++cnt;
            }
        }
        DevResult res;
        res.score = val;
        res.status = W.overflow ? 5 : (cnt > t.skl_cap ? 1 : 0);
        res.n_skl = cnt; res.pad = 0;
        results[ti] = res;
    }
}

#undef A
#undef B
#undef I53
#undef SGD

}   // namespace gspaln
