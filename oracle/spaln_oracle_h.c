/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's protein x genome
 * spliced DP with quantised intron penalty at the AVX2 lane count (16 int16 lanes):
 *   SimdAln2h1::forwardH1_wip   src/fwd2h1_wip_simd.h:50-336
 *   SimdAln2h1::fhinitH1        src/fwd2h1_simd.h:546-689   (mode 1: no Vmf, no links)
 *   SimdAln2h1::fhlastH1        src/fwd2h1_simd.h:691-789
 *   Anti_rhomb_coord<SHORT>, step 3   src/rhomb_coord.h:65-235
 * Lane k of a strip sits on query row ml+1+k and genome column n - 3k at step n; six
 * generations of H and F are kept (ring index q = (n + 3(ml+1)) % 6) because moves come from
 * 1, 2, 3 (frame shifts / codon) columns back in the own row and 3, 4, 5, 6 back in the row
 * above.  Pinned against oracle/_ref (tests/test_oracle_protein.py).
 */
#include "spaln_oracle.h"

#include <limits.h>
#include <stdlib.h>
#include <string.h>

#define NELEM 16
#define NP1 (NELEM + 1)
typedef int16_t var_t;
#define CHECK_SCR ((int) (0.9 * SHRT_MAX))
#define NEVSEL16 ((var_t) (SHRT_MIN + 1024))
#define MIN_SSV (-1000)                     /* src/fwd2h1_wip_simd.h:48 */

enum { TB_DIAG = 1, TB_HORI = 2, TB_HOR1 = 4, TB_HOR2 = 5, TB_VERT = 8, TB_VER1 = 10, TB_VER2 = 11,
       TB_ACCM = 13, TB_ACCZ = 14, TB_ACCP = 15, TB_NHOR = 16, TB_NVER = 32,
       TB_DONM = 64, TB_DONZ = 128, TB_DONP = 256 };
static const int donor_code[4] = { TB_DONM, TB_DONZ, TB_DONP, 0 };
static const int accpr_code[4] = { TB_ACCM, TB_ACCZ, TB_ACCP, 0 };
static const int next_p[3] = { 1, 2, 0 };

static inline var_t adds16(int a, int b)
{
    int x = a + b;
    return (var_t) (x > SHRT_MAX ? SHRT_MAX : (x < SHRT_MIN ? SHRT_MIN : x));
}
static inline var_t subs16(int a, int b)
{
    int x = a - b;
    return (var_t) (x > SHRT_MAX ? SHRT_MAX : (x < SHRT_MIN ? SHRT_MIN : x));
}
static inline int mod6(int q) { return q < 0 ? q + 6 : q % 6; }

/* SGPT6 accessors: 8 shorts per column */
#define SG(t, n, f) ((t)->sgpt6[8 * (n) + (f)])
enum { F_SIG5 = 0, F_SIG3, F_SIGS, F_SIGT, F_SIGE, F_SIGI, F_PHS5, F_PHS3 };

typedef struct {
    int m_base, n_base, m_width, n_width;
    uint16_t* bbuf;
    int cur_m, cur_n;
    uint16_t* cur_p;
} trb3;

static uint16_t* trb3_set_point(trb3* tb, int m, int n)
{
    tb->cur_m = m - tb->m_base;
    tb->cur_n = n - tb->n_base;
    tb->cur_p = tb->bbuf + (size_t) (3 * tb->cur_m + tb->cur_n) * tb->m_width + tb->cur_m;
    return tb->cur_p;
}
static unsigned trb3_to_left(trb3* tb, int* m, int* n, int s)
{
    *m = tb->cur_m;
    *n = tb->cur_n -= s;
    if (*n < 0) { tb->cur_n = *n = 0; return 0; }
    tb->cur_p -= (size_t) s * tb->m_width;
    return *tb->cur_p;
}
static unsigned trb3_to_upper(trb3* tb, int* m, int* n, int s)
{
    *m = --tb->cur_m;
    *n = tb->cur_n -= s;
    if (*m < 0) { tb->cur_m = *m = 0; tb->cur_n = *n += s; return 0; }
    else if (*n < 0) {
        if (s > 0) tb->cur_m = *m -= *n / s;
        tb->cur_n = *n = 0;
        return 0;
    }
    tb->cur_p -= ((size_t) (3 + s) * tb->m_width + 1);
    return *tb->cur_p;
}
/* returns 0 ok, -1 unexpected code */
static int trb3_go_back(trb3* tb, unsigned code, int* m, int* n, unsigned* out)
{
    unsigned dir = code & 15;
#define RET0 do { *out = 0; return 0; } while (0)
    switch (dir) {
      case 0: break;
      case TB_DIAG:
        do { if (!(code = trb3_to_upper(tb, m, n, 3))) RET0; } while ((code & 15) == TB_DIAG);
        break;
      case TB_HORI:
        while (!(code & TB_NHOR)) if (!(code = trb3_to_left(tb, m, n, 3))) RET0;
        dir = code & 15;
        if (dir != TB_HOR1 && dir != TB_HOR2) code = trb3_to_left(tb, m, n, 3);
        break;
      case TB_VERT:
        while (!(code & TB_NVER)) if (!(code = trb3_to_upper(tb, m, n, 0))) RET0;
        dir = code & 15;
        if (dir != TB_VER1 && dir != TB_VER2) code = trb3_to_upper(tb, m, n, 0);
        break;
      case TB_ACCZ:
        do { if (!(code = trb3_to_left(tb, m, n, 1))) RET0; } while (!(code & TB_DONZ));
        break;
      case TB_ACCM:
        do { if (!(code = trb3_to_left(tb, m, n, 1))) RET0; } while (!(code & TB_DONM));
        break;
      case TB_ACCP:
        do { if (!(code = trb3_to_left(tb, m, n, 1))) RET0; } while (!(code & TB_DONP));
        code = trb3_to_upper(tb, m, n, 3);
        ++*m; *n += 3;
        break;
      case TB_HOR1: code = trb3_to_left(tb, m, n, 1); break;
      case TB_HOR2: code = trb3_to_left(tb, m, n, 2); break;
      case TB_VER1: code = trb3_to_upper(tb, m, n, 1); break;
      case TB_VER2: code = trb3_to_upper(tb, m, n, 2); break;
      default: return -1;
    }
    *out = code;
    return 0;
#undef RET0
}

static int gap_ext_pen3(const so_params_h* p, int i) { return i > p->codonk1 ? p->lgep : p->gep; }

int so_forward_h1_wip(const so_params_h* p, const so_task_h* t, int want_trace, int32_t* score,
                      int32_t* skl, int cap)
{
    const int lw = t->lw, up = t->up;
    const int width = up - lw + 7;
    const int buf_size = width + 6 * NELEM;                 /* src/fwd2h1_simd.h:212 */
    const int a_left = t->a_left, a_right = t->a_right, b_left = t->b_left, b_right = t->b_right;
    const int Local = p->local;
    const int LocalL = Local && t->a_exgl && t->b_exgl;
    const int LocalR = Local && t->a_exgr && t->b_exgr;
    var_t* vbuf = (var_t*) malloc(sizeof(var_t) * 2 * (size_t) buf_size);
    trb3 trb;
    trb.m_base = a_left; trb.n_base = b_left;
    trb.m_width = a_right - a_left + 1;
    trb.n_width = b_right - b_left + 1 + 3 * trb.m_width;
    trb.bbuf = (uint16_t*) calloc((size_t) trb.m_width * trb.n_width + 32 + 64, sizeof(uint16_t));
    if (!vbuf || !trb.bbuf) return -1;
    var_t* hv = vbuf - lw + 3;
    var_t* fv = hv + buf_size;
    const var_t ge = (var_t) p->gep, g1 = (var_t) p->gw1, g2 = (var_t) p->gw2, g3 = (var_t) p->gw3;
    const var_t mil = (var_t) p->llmt;
    const int ipen = p->spj ? p->ipen : NEVSEL16;
    var_t quant[SO_MAXQUANT], mean[SO_MAXQUANT];
    for (int j = 0; j < p->nquant; ++j) { quant[j] = (var_t) p->quant_len[j]; mean[j] = (var_t) p->quant_pen[j]; }

    if (!t->a_exgl) {                               /* trb.initialize_m0(4) */
        uint16_t* q = trb.bbuf;
        for (int n = 1; n < trb.n_width; ++n) *(q += trb.m_width) = 4;
    }
    /* ---- fhinitH1 (mode 1) */
    {
        for (int i = 0; i < 2 * buf_size; ++i) vbuf[i] = NEVSEL16;
        const int rl = b_left - 3 * a_left;
        uint16_t* row0 = trb3_set_point(&trb, a_left, b_left);
        if (t->b_exgl == 1) { for (int r = lw; r < rl; ++r) hv[r] = 0; }
        else if (t->b_exgl == 2) fv[rl] = 0;
        int rr = b_right - 3 * a_left;
        if (up < rr) rr = up;
        int r = rl;
        if (!t->a_exgl) {
            if (t->b_exgl) fv[r] = 0;
            hv[r++] = 0;
            hv[r++] = (var_t) p->gw1;
            hv[r++] = (var_t) p->gw2;
            hv[r++] = (var_t) p->gw3;
            if (p->gep) {
                int x = (NEVSEL16 - p->gw3) / p->gep + r;
                if (x < rr) rr = x;
                for ( ; r < rr; ++r) hv[r] = (var_t) (hv[r - 3] + p->gep);
            } else if (rr > r) {
                for (int i = r; i < rr; ++i) hv[i] = hv[r - 1];
            }
        } else {
            var_t* h = hv + r;
            int n = b_left;
            int lend[3] = { r, r + 1, r + 2 };
            int bn = n + 1;                         /* bb = score_p(n + 1) */
            for (int ph = 0; ph < 3; ++r, ++n, ++bn, ++ph) {
                *h++ = SG(t, bn, F_SIGS) > 0 ? SG(t, bn, F_SIGS) : 0;
                row0 += trb.m_width;
            }
            for (int ph = 0; r < rr; ++r, ++h, ++n, ++bn, ph = next_p[ph]) {
                *h = h[-3];
                const int gl = r - lend[ph];
                if (!(t->a_exgl & 1) && gl == 3) *h = (var_t) (*h + p->gop);
                if (!(t->a_exgl & 2)) *h = (var_t) (*h + gap_ext_pen3(p, gl));
                *h = (var_t) (*h + SG(t, bn - 3, F_SIGE));
                if (*h < NEVSEL16) break;
                var_t x = (var_t) (h[-1] + p->gw1);
                if (x > *h) { *h = x; *row0 = TB_HOR1; }
                x = (var_t) (h[-2] + p->gw2);
                if (x > *h) { *h = x; *row0 = TB_HOR2; }
                x = SG(t, bn, F_SIGS) > 0 ? SG(t, bn, F_SIGS) : 0;
                if (x > *h) { *h = x; lend[ph] = r; }
                else *row0 = TB_HORI;
                row0 += trb.m_width;
            }
        }
    }

    int accscr = 0;
    const int md = (CHECK_SCR - 0) / p->avmch / NELEM * NELEM;
    int mc = md + a_left;
    const int mw = a_right - a_left;
    const int mb = a_right - NELEM;
    const int mt = a_left + mw / NELEM * NELEM;
    struct { int val, mr, nr; } maxh = { NEVSEL16, a_right, b_right };

    var_t SM[NP1], CP[3][NP1], S5[6][NP1], S3[6][NP1], P5[6][NP1], P3[6][NP1];
    var_t HA[6][NP1], FA[6][NP1], EV[3][NELEM];
    for (int ml = a_left; ml < a_right; ml += NELEM) {
        const int j9 = NELEM < a_right - ml ? NELEM : a_right - ml;
        const int j8 = j9 - 1;
        int n = b_left > lw + 3 * ml ? b_left : lw + 3 * ml;
        const int lim = b_right < up + 3 * (ml + j9) + 1 ? b_right : up + 3 * (ml + j9) + 1;
        const int n9 = lim + 3 * j9;
        const int mp1 = ml + 1;
        int q = (n + 3 * mp1) % 6;
        int r = n - 3 * mp1;
        for (int s = 0; s < 6; ++s) for (int k = 0; k < NP1; ++k) {
            HA[s][k] = FA[s][k] = NEVSEL16;
            S5[s][k] = S3[s][k] = P5[s][k] = P3[s][k] = 0;
        }
        for (int s = 0; s < 3; ++s) {
            for (int k = 0; k < NELEM; ++k) EV[s][k] = NEVSEL16;
            for (int k = 0; k < NP1; ++k) CP[s][k] = 0;
        }
        for (int k = 0; k < NP1; ++k) SM[k] = 0;
        var_t hiv[3][NELEM], hil[3][NELEM];
        for (int f = 0; f < 3; ++f) for (int k = 0; k < NELEM; ++k) { hiv[f][k] = NEVSEL16; hil[f][k] = 0; }

        for ( ; n <= n9; ++n, ++r, q = mod6(q + 1)) {
            const int ph = q % 3;
            const int nb = n - b_right + 1 > 0 ? n - b_right + 1 : 0;
            const int kb = (nb - 1) / 3;
            const int ke = j9 < (n - b_left) / 3 ? j9 : (n - b_left) / 3;
            uint16_t* dst = trb3_set_point(&trb, mp1, n);
            /* coding potential of the codon that ends at column n (good(bb - 2)) */
            var_t cv[NELEM];
            CP[ph][0] = (n - 2 >= 0 && n - 2 < t->b_len) ? SG(t, n - 2, F_SIGE) : 0;   /* data_p[-1] is zero */
            for (int k = 0; k < NELEM; ++k) cv[k] = CP[ph][k];
            for (int k = 0; k < NELEM; ++k) CP[ph][k + 1] = cv[k];

            const int q1 = mod6(q - 1), q2 = mod6(q - 2), q3 = mod6(q - 3), q4 = mod6(q - 4), q5 = mod6(q - 5);
            var_t H1[NELEM], H2[NELEM], H3[NELEM], U3[NELEM], U4[NELEM], U5[NELEM], UF[NELEM], DV[NELEM];
            for (int k = 0; k < NELEM; ++k) { H1[k] = HA[q1][k + 1]; H2[k] = HA[q2][k + 1]; H3[k] = HA[q3][k + 1]; }
            FA[q3][0] = fv[r + 3];
            for (int k = 0; k < NELEM; ++k) UF[k] = FA[q3][k];
            HA[q3][0] = hv[r + 3];
            for (int k = 0; k < NELEM; ++k) U3[k] = HA[q3][k];
            HA[q4][0] = hv[r + 2];
            for (int k = 0; k < NELEM; ++k) U4[k] = HA[q4][k];
            HA[q5][0] = hv[r + 1];
            for (int k = 0; k < NELEM; ++k) U5[k] = HA[q5][k];
            if (nb) for (int k = 0; k < NELEM; ++k) SM[k] = 0;
            for (int k = kb; k < ke; ++k)
                SM[k] = (var_t) p->simmtx[t->a[ml + k] * p->simdim + t->b[n - 3 * k - 2]];
            HA[q][0] = hv[r];
            for (int k = 0; k < NELEM; ++k) DV[k] = HA[q][k];

            /* splice signals entering lane 0 at this step */
            var_t s3v[2][NELEM], p3v[2][NELEM], s5v[2][NELEM], p5v[2][NELEM];
            if (p->spj) {
                for (int kk = 0; kk < 2; ++kk) {
                    const int pk = 2 * ph + kk;
                    int phs = nb ? -2 : SG(t, n, F_PHS3);
                    int leg = !nb && phs > -2 && (!kk || phs == 2);
                    int phase = leg ? (phs == 2 ? (kk ? 1 : -1) : (kk ? 2 : phs)) : 2;
                    S3[pk][0] = phase < 2 ? SG(t, n - phase, F_SIG3) : MIN_SSV;
                    P3[pk][0] = (var_t) accpr_code[phase + 1];
                    for (int k = 0; k < NELEM; ++k) { s3v[kk][k] = S3[pk][k]; p3v[kk][k] = P3[pk][k]; }
                    for (int k = 0; k < NELEM; ++k) { S3[pk][k + 1] = s3v[kk][k]; P3[pk][k + 1] = p3v[kk][k]; }
                    phs = nb ? -2 : SG(t, n, F_PHS5);
                    leg = !nb && phs > -2 && (!kk || phs == 2);
                    phase = leg ? (phs == 2 ? (kk ? 1 : -1) : (kk ? 2 : phs)) : 2;
                    S5[pk][0] = phase < 2 ? (var_t) (SG(t, n - phase, F_SIG5) + ipen) : MIN_SSV;
                    P5[pk][0] = (var_t) donor_code[phase + 1];
                    for (int k = 0; k < NELEM; ++k) { s5v[kk][k] = S5[pk][k]; p5v[kk][k] = P5[pk][k]; }
                    for (int k = 0; k < NELEM; ++k) { S5[pk][k + 1] = s5v[kk][k]; P5[pk][k + 1] = p5v[kk][k]; }
                }
            }

            /* `if (AllZero(ph_v)) continue;` (src/fwd2h1_wip_simd.h:223): a phase slot none of
             * whose 16 lanes carries an acceptor is skipped for the whole vector */
            int any3[2] = { 0, 0 };
            if (p->spj)
                for (int kk = 0; kk < 2; ++kk)
                    for (int k = 0; k < NELEM; ++k) if (p3v[kk][k]) any3[kk] = 1;

            uint16_t tb[NELEM];
            for (int k = 0; k < NELEM; ++k) {
                unsigned hb, pb, eb;
                var_t h, x, e, f;
                /* horizontal: 1- / 2-nt frame shifts, codon insertion, extension */
                h = adds16(H1[k], g1);
                x = adds16(H2[k], g2);
                if (h > x) eb = TB_HOR1; else { h = x; eb = TB_HOR2; }
                x = adds16(adds16(H3[k], g3), cv[k]);
                if (!(h > x)) { h = x; eb = TB_HORI; }
                e = adds16(adds16(EV[ph][k], ge), cv[k]);
                if (e > h) { hb = 0; eb = TB_HORI; } else { e = h; hb = TB_NHOR; }
                EV[ph][k] = e;
                /* vertical: codon deletion, frame shifts, extension */
                f = adds16(UF[k], ge);
                h = adds16(U3[k], g3);
                x = adds16(U4[k], g2);
                if (h > x) pb = TB_VERT; else { h = x; pb = TB_VER1; }
                x = adds16(U5[k], g1);
                if (!(h > x)) { h = x; pb = TB_VER2; }
                if (f > h) pb = TB_VERT; else { f = h; hb |= TB_NVER; }
                FA[q][k + 1] = f;
                /* diagonal */
                h = adds16(adds16(SM[k], DV[k]), cv[k]);
                if (f > h) h = f; else pb = TB_DIAG;
                if (e > h) { h = e; pb = eb; }
                /* acceptor */
                unsigned ab = 0;
                if (p->spj) {
                    for (int kk = 0; kk < 2; ++kk) {
                        if (!any3[kk]) continue;
                        for (int fz = kk ? 2 : 0; fz < 3; ++fz) {
                            var_t qv = adds16(hiv[fz][k], s3v[kk][k]);
                            var_t pen = mean[0];
                            for (int j = 1; j < p->nquant; ++j)
                                if (hil[fz][k] > quant[j - 1]) pen = mean[j];
                            qv = adds16(qv, pen);
                            if (!(p3v[kk][k] == accpr_code[fz])) qv = NEVSEL16;
                            if (!(hil[fz][k] > mil)) qv = NEVSEL16;
                            if (qv > h) { h = qv; pb = accpr_code[fz]; ab |= (unsigned) (uint16_t) p3v[kk][k]; }
                        }
                    }
                }
                if (LocalL && !accscr && 0 > h) { h = 0; hb = 0; }
                HA[q][k + 1] = h;
                /* donor */
                if (p->spj) {
                    for (int kk = 0; kk < 2; ++kk) {
                        const var_t qv = adds16(h, s5v[kk][k]);
                        for (int fz = kk ? 2 : 0; fz < 3; ++fz) {
                            var_t pv = fz == 2 ? adds16(DV[k], s5v[kk][k]) : qv;
                            if (ab) pv = NEVSEL16;                  /* non-empty exon */
                            if (!(p5v[kk][k] == donor_code[fz])) pv = NEVSEL16;
                            if (pv > hiv[fz][k]) {
                                hiv[fz][k] = pv;
                                hb |= (unsigned) donor_code[fz];
                                hil[fz][k] = 0;
                            }
                        }
                    }
                    for (int fz = 0; fz < 3; ++fz) hil[fz][k] = adds16(hil[fz][k], 1);
                }
                tb[k] = (uint16_t) (hb | pb);
            }
            if (LocalR) {
                int best = 1;
                for (int k = 2; k <= j9; ++k) if (HA[q][k] > HA[q][best]) best = k;
                if (HA[q][best] + accscr > maxh.val) {
                    maxh.val = HA[q][best] + accscr;
                    maxh.mr = ml + best;
                    maxh.nr = n - 3 * best + 3;
                }
            }
            const int r0 = r - 6 * j8;
            if (j9 == ke && lw <= r0 && r0 <= up) {
                hv[r0] = HA[q][j9];
                fv[r0] = FA[q][j9];
            }
            if (ml == mt) for (int k = a_right - mt; k < NELEM; ++k) tb[k] = 0;
            if (ml > mb) for (int k = 0; k < NELEM; ++k) dst[k] |= tb[k];
            else memcpy(dst, tb, sizeof(tb));
        }
        if (ml == mc) {
            /* vec_max / vec_sub_c over `width` entries from hv + lw - 3 (wip.h:318-327) */
            var_t* base = hv + lw - 3;
            var_t c = base[0];
            for (int i = 1; i < width; ++i) if (base[i] > c) c = base[i];
            const int d = (CHECK_SCR - abs(c)) / p->avmch / NELEM * NELEM;
            if (d < md / 2) {
                const int nn = width / NELEM * NELEM;
                for (int i = 0; i < width; ++i) {
                    base[i] = i < nn ? subs16(base[i], c) : (var_t) (base[i] - c);
                    (fv + lw - 3)[i] = i < nn ? subs16((fv + lw - 3)[i], c) : (var_t) ((fv + lw - 3)[i] - c);
                }
                accscr += c;
                mc += md;
            } else
                mc += d;
        }
    }

    if (!LocalR || maxh.mr == a_right) {
        /* ---- fhlastH1 (mode 1, no Vmf) */
        int glen[3] = { 0, 0, 0 };
        int tcdn[3] = { 0, 0, 0 };
        const int m3 = 3 * a_right;
        int rw = lw;
        int rf = b_left - m3;
        if (rf > rw) rw = rf; else rf = rw;
        const int rr = b_right - m3;
        int maxr = rr;
        var_t* h = hv + rw;
        var_t* h9 = hv + rr;
        var_t* mx = h9;
        int bn = rw + m3;                           /* bb = score_p(rw + m3) */
        uint16_t* rowM = trb3_set_point(&trb, a_right, rw + m3);
        int done = 0;
        if (t->a_exgr) {
            for (int ph = 0; h <= h9; ++h, ++rf, ++bn, ph = next_p[ph]) {
                glen[ph] += 3;
                int cand[3] = { *h, NEVSEL16, NEVSEL16 };
                if (rf - rw >= 3 && !tcdn[ph]) {
                    cand[1] = h[-3] + SG(t, bn - 2, F_SIGE);
                    if (!(t->a_exgr & 2)) cand[1] += gap_ext_pen3(p, glen[ph]);
                    if (!(t->a_exgr & 1) && glen[ph] == 3) cand[1] += p->gop;
                    if (p->lcl & 2) cand[2] = h[-3] + SG(t, bn - 2, F_SIGT);
                }
                if (rf - rw >= 3) tcdn[ph] = (tcdn[ph] || SG(t, bn - 2, F_SIGT) > 0);
                const var_t sig5 = (Local && SG(t, bn, F_SIG5) > 0) ? SG(t, bn, F_SIG5) : 0;
                cand[0] += sig5;
                cand[1] += sig5;
                int k = 0;
                if (cand[1] > cand[k]) k = 1;
                if (cand[2] > cand[k]) k = 2;
                if (k == 0) { glen[ph] = 0; tcdn[ph] = 0; }
                else if (k == 1) { *h = (var_t) (cand[1] - sig5); *rowM = TB_HORI; }
                else { *h = (var_t) cand[2]; *rowM = TB_HORI; }
                if (*h > *mx) { mx = h; maxr = rf - ((k == 2) ? 3 : 0); }
                if (glen[ph] == 3) *rowM |= TB_NHOR;
                rowM += trb.m_width;
            }
        } else {
            bn += (int) (h9 - h);
            const var_t y = (var_t) (h9[-3] + SG(t, bn, F_SIGT));
            if (y > *h9) { *h9 = y; maxr = rr - 3; }
        }
        if (t->b_exgr) {
            rw = up - 1 < b_right - 3 * a_left ? up - 1 : b_right - 3 * a_left;
            var_t g[3] = { NEVSEL16, NEVSEL16, NEVSEL16 };
            h = hv + rw - 3;
            for (int ph = 0; h > h9; --h, --rw, ph = next_p[ph]) {
                var_t x = h[3];
                if (!(t->b_exgr & 1)) x = (var_t) (x + p->gop);
                if (x > g[ph]) g[ph] = x;
                if (!(t->b_exgr & 2)) g[ph] = (var_t) (g[ph] + p->gep);
                if (*h > g[ph]) g[ph] = NEVSEL16;
                else if (g[ph] > *mx) *(mx = h) = g[ph];
            }
        } else if (t->b_exgr == 2) {
            done = 1;                               /* return (rr) before touching maxh */
        }
        if (!done) {
            const int maxt = (int) (mx - hv);
            const int pp = maxr - rr;
            if (pp > 0) maxh.mr = (b_right - maxr) / 3;
            else maxh.nr = maxt + m3;
        }
        maxh.val += accscr;
    }
    int cnt = 0;
    if (want_trace) {
        int m = maxh.mr, n = maxh.nr;
        unsigned code = 0;
        /* fhlastH1 can return a start point right of b_right (its last-column scan moves mx but
         * not maxr, src/fwd2h1_simd.h:764-788): the reference then reads outside its trace
         * buffer (undefined).  Here such a walk stops at once. */
        if (m - trb.m_base >= 0 && m - trb.m_base < trb.m_width && n - trb.n_base >= 0 &&
            3 * (m - trb.m_base) + (n - trb.n_base) < trb.n_width)
            code = *trb3_set_point(&trb, m, n);
        else
            trb3_set_point(&trb, trb.m_base, trb.n_base);
        m -= trb.m_base; n -= trb.n_base;
        while (code) {
            if (cnt < cap) { skl[2 * cnt] = m + trb.m_base; skl[2 * cnt + 1] = n + trb.n_base; }
            ++cnt;
            if (trb3_go_back(&trb, code, &m, &n, &code) < 0) { cnt = -2; break; }
            /* a start point outside the matrix (reference: undefined for some mixed end-gap
             * flag combinations) can make the walk cycle: give up instead of spinning */
            if (cnt > 4 * (trb.m_width + trb.n_width)) { cnt = -2; break; }
        }
        if (cnt >= 0) {
            if (cnt < cap) { skl[2 * cnt] = m + trb.m_base; skl[2 * cnt + 1] = n + trb.n_base; }
            ++cnt;
        }
    }
    *score = maxh.val;
    free(trb.bbuf);
    free(vbuf);
    return cnt;
}
