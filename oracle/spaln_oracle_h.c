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

            /* `if (AllZero(ph_v)) continue;` (src/fwd2h1_wip_simd.h:223, 276) never skips in the
             * canonical AVX2 build: Simd_functions<short>::all_zero is _mm256_testnzc_si256(v, v)
             * (src/simd_functions.h:1057), which is 0 for every v.  (The SSE4.1 build does skip.) */
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

/* =========================================================================
 * SimdAln2h1::hirschbergH1_wip (src/fwd2h1_wip_simd.h:338-773): the forward pass of the
 * multi-intermediate unidirectional Hirschberg method for protein x genome, with fhinitH1 /
 * fhlastH1 in mode 2 (links, no Vmf: src/fwd2h1_simd.h:546-789).  Links are kept as ints
 * (the reference splits them over two int16 lanes only when mode == 4).  Lane memory the
 * reference never initialises (hc_a / fc_a / ec_a before the first strip, cbuf entries
 * fhinitH1 does not set) is zero here.
 * ========================================================================= */
#define END_OF_ULK_H (INT_MAX - 2)
#define NEVSEL32_H (INT_MIN / 16 * 7)

typedef struct {
    int mi;
    int* buf;
    int* hlnk[2];
    int* vlnk[2];
} so_imd_h;

static int imd_h_init(so_imd_h* im, int mi, int lw, int width)
{
    /* UdhIntermediate(m, wdw, nol = 2), src/udh_intermediate.h:37-58 */
    const size_t u_size = (size_t) 2 * width;
    im->mi = mi;
    im->buf = (int*) malloc(2 * u_size * sizeof(int));
    if (!im->buf) return -1;
    for (size_t i = 0; i < 2 * u_size; ++i) im->buf[i] = END_OF_ULK_H;
    im->hlnk[0] = im->buf - lw + 1;
    im->vlnk[0] = im->hlnk[0] + u_size;
    im->hlnk[1] = im->hlnk[0] + width;
    im->vlnk[1] = im->vlnk[0] + width;
    return 0;
}

int so_hirschberg_h1_wip(const so_params_h* p, const so_task_h* t, int n_im, int32_t* score,
                         int32_t* cpos /* (n_im + 1) x 10 */, int32_t* ranges)
{
    if (n_im < 1) return -3;
    const int lw = t->lw, up = t->up;
    const int width = up - lw + 7;
    const int buf_size = width + 6 * NELEM;
    int a_left = t->a_left, a_right = t->a_right, b_left = t->b_left, b_right = t->b_right;
    const int Local = p->local;
    const int LocalL = Local && t->a_exgl && t->b_exgl;
    const int LocalR = Local && t->a_exgr && t->b_exgr;
    var_t* vbuf = (var_t*) malloc(sizeof(var_t) * 2 * (size_t) buf_size);
    var_t* bbuf = (var_t*) malloc(sizeof(var_t) * 2 * (size_t) buf_size);
    int* cbuf = (int*) calloc(2 * (size_t) buf_size, sizeof(int));
    so_imd_h* imds = (so_imd_h*) calloc((size_t) n_im, sizeof(so_imd_h));
    if (!vbuf || !bbuf || !cbuf || !imds) return -1;
    var_t* hv = vbuf - lw + 3; var_t* fv = hv + buf_size;
    var_t* hb = bbuf - lw + 3; var_t* fb = hb + buf_size;
    int* hc = cbuf - lw + 3; int* fc = hc + buf_size;
    const var_t ge = (var_t) p->gep, g1 = (var_t) p->gw1, g2 = (var_t) p->gw2, g3 = (var_t) p->gw3;
    const var_t mil = (var_t) p->llmt;
    const int ipen = p->spj ? p->ipen : NEVSEL16;
    var_t quant[SO_MAXQUANT], mean[SO_MAXQUANT];
    for (int j = 0; j < p->nquant; ++j) { quant[j] = (var_t) p->quant_len[j]; mean[j] = (var_t) p->quant_pen[j]; }
    for (int i = 0; i <= n_im; ++i) for (int j = 0; j < 10; ++j) cpos[10 * i + j] = END_OF_ULK_H;

    /* ---- fhinitH1, mode 2 */
    {
        for (int i = 0; i < 2 * buf_size; ++i) { vbuf[i] = NEVSEL16; bbuf[i] = (var_t) a_left; }
        const int rl = b_left - 3 * a_left;
        {
            int r = lw;
            const int rre = t->a_exgl ? rl : up;
            for ( ; r < rre; ++r) hc[r] = r;
            for (int i = 0, rq = rl; rq >= lw; --rq) hb[rq] = (var_t) (a_left + (i++ / 3));
        }
        if (t->b_exgl == 1) { for (int r = lw; r < rl; ++r) hv[r] = 0; }
        else if (t->b_exgl == 2) { fv[rl] = 0; fc[rl] = rl; }
        int rr = b_right - 3 * a_left;
        if (up < rr) rr = up;
        int r = rl;
        if (!t->a_exgl) {
            if (t->b_exgl) { fv[r] = 0; fc[r] = hc[r]; }
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
            int* c = hc + r;
            int lend[3] = { r, r + 1, r + 2 };
            int bn = b_left + 1;
            for (int ph = 0; ph < 3; ++r, ++bn, ++ph) {
                *h++ = SG(t, bn, F_SIGS) > 0 ? SG(t, bn, F_SIGS) : 0;
                *c++ = r;
            }
            for (int ph = 0; r < rr; ++r, ++h, ++bn, ph = next_p[ph]) {
                *h = h[-3];
                *c = c[-3];
                const int gl = r - lend[ph];
                if (!(t->a_exgl & 1) && gl == 3) *h = (var_t) (*h + p->gop);
                if (!(t->a_exgl & 2)) *h = (var_t) (*h + gap_ext_pen3(p, gl));
                *h = (var_t) (*h + SG(t, bn - 3, F_SIGE));
                if (*h < NEVSEL16) break;
                var_t x = (var_t) (h[-1] + p->gw1);
                if (x > *h) { *h = x; *c = c[-1]; }
                x = (var_t) (h[-2] + p->gw2);
                if (x > *h) { *h = x; *c = c[-2]; }
                x = SG(t, bn, F_SIGS) > 0 ? SG(t, bn, F_SIGS) : 0;
                if (x > *h) { *h = x; lend[ph] = r; *c = r; }
                ++c;
            }
        }
    }

    /* ---- intermediates (wip.h:366-371) */
    int mm = (a_right - a_left + n_im) / (n_im + 1);
    {
        int mi = a_left;
        for (int i = 0; i < n_im; ++i)
            if (imd_h_init(&imds[i], mi += mm, lw, width)) return -1;
    }
    so_imd_h* imd = &imds[0];
    mm = a_left + (imd->mi - a_left - 1) / NELEM * NELEM;
    int k9 = imd->mi - mm, k8 = k9 - 1;
    int rlst[3] = { INT_MAX, INT_MAX, INT_MAX };

    int accscr = 0;
    const int md = (CHECK_SCR - 0) / p->avmch / NELEM * NELEM;
    int mc = md + a_left;
    struct { int val, ulk, ml, mr, nr; } maxh = { NEVSEL16, END_OF_ULK_H, (int16_t) a_left, (int16_t) a_right, b_right };

    var_t SM[NP1], CP[3][NP1], S5[6][NP1], S3[6][NP1], P5[6][NP1], P3[6][NP1];
    var_t HA[6][NP1], FA[6][NP1], EV[3][NELEM], PV[3][NELEM];
    var_t HB[6][NP1], FB[6][NP1], EB[3][NELEM];
    int HC[6][NP1], FC[6][NP1], EC[3][NELEM];
    memset(HC, 0, sizeof(HC)); memset(FC, 0, sizeof(FC)); memset(EC, 0, sizeof(EC));
    memset(PV, 0, sizeof(PV));
    for (int k = 0; k < NP1; ++k) SM[k] = 0;

    for (int ml = a_left, ii = 0; ml < a_right; ml += NELEM) {
        const int j9 = NELEM < a_right - ml ? NELEM : a_right - ml;
        const int j8 = j9 - 1;
        int n = b_left > lw + 3 * ml ? b_left : lw + 3 * ml;
        const int lim = b_right < up + 3 * (ml + j9) + 1 ? b_right : up + 3 * (ml + j9) + 1;
        const int n9 = lim + 3 * j9;
        const int mp1 = ml + 1;
        int q = mod6(n + 3 * mp1);
        int r = n - 3 * mp1;
        int donor_r[3] = { r, r, r };
        for (int s = 0; s < 6; ++s) for (int k = 0; k < NP1; ++k) {
            HA[s][k] = FA[s][k] = NEVSEL16;
            HB[s][k] = FB[s][k] = 0;
            S5[s][k] = S3[s][k] = P5[s][k] = P3[s][k] = 0;
        }
        for (int s = 0; s < 3; ++s) {
            for (int k = 0; k < NELEM; ++k) { EV[s][k] = NEVSEL16; EB[s][k] = 0; }
            for (int k = 0; k < NP1; ++k) CP[s][k] = 0;
        }
        for (int k = 0; k < NP1; ++k) SM[k] = 0;
        const int is_imd_ = ml == mm;
        var_t hiv[3][NELEM], hib[3][NELEM], hil[3][NELEM];
        int hic[3][NELEM];
        for (int f = 0; f < 3; ++f) for (int k = 0; k < NELEM; ++k) {
            hiv[f][k] = NEVSEL16; hib[f][k] = 0; hic[f][k] = 0; hil[f][k] = 0;
        }

        for ( ; n < n9; ++n, ++r, q = mod6(q + 1)) {
            const int ph = q % 3;
            const int rj = r - 6 * k8;
            const int nb = n - b_right + 1 > 0 ? n - b_right + 1 : 0;
            const int kb = (nb - 1) / 3;
            const int ke = j9 < (n - b_left) / 3 ? j9 : (n - b_left) / 3;
            const int is_imd = is_imd_ && rj >= lw && rj <= up;
            var_t cv[NELEM];
            CP[ph][0] = (n - 2 >= 0 && n - 2 < t->b_len) ? SG(t, n - 2, F_SIGE) : 0;
            for (int k = 0; k < NELEM; ++k) cv[k] = CP[ph][k];
            for (int k = 0; k < NELEM; ++k) CP[ph][k + 1] = cv[k];

            const int q1 = mod6(q - 1), q2 = mod6(q - 2), q3 = mod6(q - 3), q4 = mod6(q - 4), q5 = mod6(q - 5);
            var_t ev[NELEM], ebv[NELEM], fvv[NELEM], fbv[NELEM], hvv[NELEM], hbv[NELEM], dv[NELEM], pbv[NELEM];
            int ecv[NELEM], fcv[NELEM], hcv[NELEM];
            /* ---- horizontal (wip.h:412-461) */
            for (int k = 0; k < NELEM; ++k) {
                var_t h = adds16(HA[q1][k + 1], g1); var_t b_ = HB[q1][k + 1]; int c_ = HC[q1][k + 1];
                var_t x = adds16(HA[q2][k + 1], g2);
                if (!(h > x)) { h = x; b_ = HB[q2][k + 1]; c_ = HC[q2][k + 1]; }
                x = adds16(adds16(HA[q3][k + 1], g3), cv[k]);
                if (!(h > x)) { h = x; b_ = HB[q3][k + 1]; c_ = HC[q3][k + 1]; }
                var_t e = adds16(adds16(EV[ph][k], ge), cv[k]);
                if (e > h) { ev[k] = e; ebv[k] = EB[ph][k]; ecv[k] = EC[ph][k]; }
                else { ev[k] = h; ebv[k] = b_; ecv[k] = c_; }
                EV[ph][k] = ev[k]; if (LocalL) EB[ph][k] = ebv[k]; EC[ph][k] = ecv[k];
            }
            /* ---- vertical (wip.h:463-530) */
            FA[q3][0] = fv[r + 3]; if (LocalL) FB[q3][0] = fb[r + 3]; FC[q3][0] = fc[r + 3];
            HA[q3][0] = hv[r + 3]; if (LocalL) HB[q3][0] = hb[r + 3]; HC[q3][0] = hc[r + 3];
            HA[q4][0] = hv[r + 2]; if (LocalL) HB[q4][0] = hb[r + 2]; HC[q4][0] = hc[r + 2];
            HA[q5][0] = hv[r + 1]; if (LocalL) HB[q5][0] = hb[r + 1]; HC[q5][0] = hc[r + 1];
            for (int k = 0; k < NELEM; ++k) {
                var_t h = adds16(FA[q3][k], ge); var_t b_ = FB[q3][k]; int c_ = FC[q3][k];
                var_t x = adds16(HA[q3][k], g3);
                if (!(h > x)) { h = x; b_ = HB[q3][k]; c_ = HC[q3][k]; }
                x = adds16(HA[q4][k], g2);
                if (!(h > x)) { h = x; b_ = HB[q4][k]; c_ = HC[q4][k]; }
                x = adds16(HA[q5][k], g1);
                if (!(h > x)) { h = x; b_ = HB[q5][k]; c_ = HC[q5][k]; }
                fvv[k] = h; fbv[k] = b_; fcv[k] = c_;
            }
            for (int k = 0; k < NELEM; ++k) {
                FA[q][k + 1] = fvv[k]; if (LocalL) FB[q][k + 1] = fbv[k]; FC[q][k + 1] = fcv[k];
            }
            /* ---- diagonal (wip.h:532-551) */
            if (nb) for (int k = 0; k < NELEM; ++k) SM[k] = 0;
            for (int k = kb; k < ke; ++k)
                SM[k] = (var_t) p->simmtx[t->a[ml + k] * p->simdim + t->b[n - 3 * k - 2]];
            HA[q][0] = hv[r]; if (LocalL) HB[q][0] = hb[r]; HC[q][0] = hc[r];
            for (int k = 0; k < NELEM; ++k) {
                dv[k] = HA[q][k];
                var_t h = adds16(adds16(SM[k], dv[k]), cv[k]);
                var_t b_ = HB[q][k]; int c_ = HC[q][k];
                var_t pb = 0;
                if (fvv[k] > h) { h = fvv[k]; b_ = fbv[k]; c_ = fcv[k]; pb = 2; }
                if (ev[k] > h) { h = ev[k]; b_ = ebv[k]; c_ = ecv[k]; pb = 1; }
                hvv[k] = h; hbv[k] = b_; hcv[k] = c_; pbv[k] = pb;
            }
            if (is_imd) for (int k = 0; k < NELEM; ++k) PV[ph][k] = pbv[k];
            /* ---- acceptors (wip.h:560-606) */
            var_t ab[NELEM];
            for (int k = 0; k < NELEM; ++k) ab[k] = 0;
            if (p->spj) {
                for (int kk = 0; kk < 2; ++kk) {
                    const int pk = 2 * ph + kk;
                    int phs = nb ? -2 : SG(t, n, F_PHS3);
                    int leg = !nb && phs > -2 && (!kk || phs == 2);
                    int phase = leg ? (phs == 2 ? (kk ? 1 : -1) : (kk ? 2 : phs)) : 2;
                    S3[pk][0] = phase < 2 ? SG(t, n - phase, F_SIG3) : MIN_SSV;
                    P3[pk][0] = (var_t) accpr_code[phase + 1];
                    var_t ss[NELEM], pv_[NELEM];
                    for (int k = 0; k < NELEM; ++k) { ss[k] = S3[pk][k]; pv_[k] = P3[pk][k]; }
                    for (int k = 0; k < NELEM; ++k) { S3[pk][k + 1] = ss[k]; P3[pk][k + 1] = pv_[k]; }
                    /* AllZero(ph_v) is always false in the AVX2 build: never skipped */
                    for (int fz = 0; fz < 3; ++fz) {
                        var_t flag[NELEM];
                        for (int k = 0; k < NELEM; ++k) {
                            var_t qv = adds16(hiv[fz][k], ss[k]);
                            var_t pen = mean[0];
                            for (int j = 1; j < p->nquant; ++j)
                                if (hil[fz][k] > quant[j - 1]) pen = mean[j];
                            qv = adds16(qv, pen);
                            if (!(pv_[k] == accpr_code[fz])) qv = NEVSEL16;
                            if (!(hil[fz][k] > mil)) qv = NEVSEL16;
                            flag[k] = 0;
                            if (qv > hvv[k]) {
                                hvv[k] = qv;
                                if (LocalL) hbv[k] = hib[fz][k];
                                hcv[k] = hic[fz][k];
                                flag[k] = 1;
                            }
                            ab[k] |= flag[k];
                        }
                        if (is_imd) {
                            for (int k = 0; k < NELEM; ++k) SM[k] = flag[k];    /* Store(sm_a, qv_v) */
                            if (SM[k8]) {
                                imd->hlnk[0][rj] = donor_r[fz];
                                imd->hlnk[1][rj] = donor_r[fz] + width;
                                rlst[ph] = rj;
                            }
                        }
                    }
                }
            }
            /* ---- store H, left / right ends (wip.h:608-639) */
            if (LocalL && !accscr) for (int k = 0; k < NELEM; ++k) if (0 > hvv[k]) hvv[k] = 0;
            for (int k = 0; k < NELEM; ++k) {
                HA[q][k + 1] = hvv[k]; if (LocalL) HB[q][k + 1] = hbv[k]; HC[q][k + 1] = hcv[k];
            }
            if (LocalL && !accscr) {
                for (int k = kb; k < ke; ++k) {
                    const int kp1 = k + 1;
                    if (HA[q][kp1] == 0) {
                        HB[q][kp1] = (var_t) (ml + k);
                        HC[q][kp1] = r - 6 * k;
                    }
                }
            }
            if (LocalR) {
                int best = 1;
                for (int k = 2; k <= j9; ++k) if (HA[q][k] > HA[q][best]) best = k;
                if (HA[q][best] + accscr > maxh.val) {
                    maxh.val = HA[q][best] + accscr;
                    maxh.ml = HB[q][best];
                    maxh.ulk = HC[q][best];
                    maxh.mr = ml + best + 1;            /* literal: k = mx - hv_a[q] (wip.h:634-635) */
                    maxh.nr = n - 3 * best;
                }
            }
            /* ---- donors (wip.h:643-681): registers hv_v / hb_v / hc_v, not the patched lanes */
            if (p->spj) {
                for (int kk = 0; kk < 2; ++kk) {
                    const int pk = 2 * ph + kk;
                    int phs = nb ? -2 : SG(t, n, F_PHS5);
                    int leg = !nb && phs > -2 && (!kk || phs == 2);
                    int phase = leg ? (phs == 2 ? (kk ? 1 : -1) : (kk ? 2 : phs)) : 2;
                    S5[pk][0] = phase < 2 ? (var_t) (SG(t, n - phase, F_SIG5) + ipen) : MIN_SSV;
                    P5[pk][0] = (var_t) donor_code[phase + 1];
                    var_t ss[NELEM], pv_[NELEM];
                    for (int k = 0; k < NELEM; ++k) { ss[k] = S5[pk][k]; pv_[k] = P5[pk][k]; }
                    for (int k = 0; k < NELEM; ++k) { S5[pk][k + 1] = ss[k]; P5[pk][k + 1] = pv_[k]; }
                    for (int fz = kk ? 2 : 0; fz < 3; ++fz) {
                        var_t flag[NELEM];
                        for (int k = 0; k < NELEM; ++k) {
                            var_t pvv = fz == 2 ? adds16(dv[k], ss[k]) : adds16(hvv[k], ss[k]);
                            if (ab[k]) pvv = NEVSEL16;
                            if (!(pv_[k] == donor_code[fz])) pvv = NEVSEL16;
                            flag[k] = 0;
                            if (pvv > hiv[fz][k]) {
                                hiv[fz][k] = pvv;
                                hil[fz][k] = 0;
                                if (LocalL) hib[fz][k] = hbv[k];
                                hic[fz][k] = hcv[k];
                                flag[k] = 1;
                            }
                            hil[fz][k] = adds16(hil[fz][k], 1);     /* inside the phase loop (wip.h:669): the phase +1 counter advances twice per step */
                        }
                        if (is_imd) {
                            for (int k = 0; k < NELEM; ++k) SM[k] = flag[k];
                            if (SM[k8]) donor_r[fz] = rj;
                        }
                    }
                }
            }
            /* ---- intermediate row (wip.h:684-694) */
            if (is_imd) {
                for (int k = 0; k < NELEM; ++k) SM[k] = ab[k];
                if (PV[ph][k8] == 0) rlst[ph] = rj;
                if (!SM[k8] && PV[ph][k8] == 1) imd->hlnk[0][rj] = rlst[ph];
                imd->vlnk[0][rj] = HC[q][k9];
                HC[q][k9] = rj;
                imd->vlnk[1][rj] = FC[q][k9];
                FC[q][k9] = rj + width;
            }
            const int r0 = r - 6 * j8;
            if (j9 == ke && lw <= r0 && r0 <= up) {
                hv[r0] = HA[q][j9]; if (LocalL) hb[r0] = HB[q][j9]; hc[r0] = HC[q][j9];
                fv[r0] = FA[q][j9]; if (LocalL) fb[r0] = FB[q][j9]; fc[r0] = FC[q][j9];
            }
        }
        if (ml == mc) {
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
        if (is_imd_ && ++ii < n_im) {
            imd = &imds[ii];
            mm = a_left + (imd->mi - a_left - 1) / NELEM * NELEM;
            k9 = imd->mi - mm;
            k8 = k9 - 1;
        }
    }

    if (LocalR && maxh.mr < a_right) {
        a_right = maxh.mr;
        b_right = maxh.nr;
    } else {
        /* ---- fhlastH1, mode 2 (no trace store, no Vmf) */
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
        int bn = rw + m3;
        int ret = rr, early = 0;
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
                else if (k == 1) *h = (var_t) (cand[1] - sig5);
                else *h = (var_t) cand[2];
                if (*h > *mx) { mx = h; maxr = rf - ((k == 2) ? 3 : 0); }
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
            early = 1;
        }
        if (!early) {
            const int maxt = (int) (mx - hv);
            hb[maxt] = hb[maxr];
            maxh.ulk = hc[maxr];
            if (maxr - rr > 0) maxh.mr = (b_right - maxr) / 3;
            else maxh.nr = maxt + m3;
            ret = maxt;
        }
        maxh.val += accscr;
        maxh.ml = LocalL ? hb[ret] : a_left;
        a_right = maxh.mr;
        b_right = maxh.nr;
    }

    /* ---- back-walk over the intermediates (wip.h:736-771) */
    int i = n_im;
    while (--i >= 0 && imds[i].mi > a_right) ;
    if (i < 0 && imds[0].mi > a_right) cpos[2] = b_right;
    int r = maxh.ulk;
    for ( ; i >= 0 && (imd = &imds[i])->mi > maxh.ml; --i) {
        int cc = 0, d = 0;
        for ( ; r > up; r -= width) ++d;
        if (d > 1 || r < lw - 1 || r >= lw - 1 + width) break;     /* foreign memory in the reference */
        if (imd->vlnk[d][r] < END_OF_ULK_H) {
            cpos[10 * i + cc++] = imd->mi;
            cpos[10 * i + cc++] = (d > 0) ? 1 : 0;
            const int mm3 = 3 * imd->mi;
            for (int rp = imd->hlnk[d][r]; lw <= rp && rp < up && r != rp && cc < 8; rp = imd->hlnk[d][r = rp])
                cpos[10 * i + cc++] = r + mm3;
            cpos[10 * i + cc++] = r + mm3;
            cpos[10 * i + cc] = END_OF_ULK_H;
            r = imd->vlnk[d][r];
            if (r == END_OF_ULK_H) break;
        } else
            cpos[10 * i + 0] = END_OF_ULK_H;
    }
    for ( ; r > up; r -= width) ;
    if (LocalL) {
        a_left = maxh.ml;
        b_left = r + 3 * a_left;
    } else {
        const int rl = b_left - 3 * a_left;
        if (t->b_exgl && rl > r) {
            a_left = (b_left - r) / 3;
            for (int j = 0; j < n_im && imds[j].mi < a_left; ++j) cpos[10 * j + 0] = END_OF_ULK_H;
        }
        if (t->a_exgl && rl < r) b_left = 3 * a_left + r;
    }
    ++i;
    int bad = 0;
    if (i >= 0 && i < n_im && imds[i].mi < a_left) bad = 1;
    if (!bad && cpos[10 * i + 2] < b_left) bad = 1;
    *score = bad ? NEVSEL32_H : maxh.val;
    ranges[0] = a_left; ranges[1] = a_right; ranges[2] = b_left; ranges[3] = b_right;
    for (int k = 0; k < n_im; ++k) free(imds[k].buf);
    free(imds); free(bbuf); free(cbuf); free(vbuf);
    return 0;
}


/* =========================================================================
 * The protein DP driver: Aln2h1::lspH_ng (src/fwd2h1.cc:2134-2230) with trcbkalignH_ng
 * (1997-2041, SIMD branch), diagonalH_ng (1963-1995), mimd_postwork (2045-2090),
 * rcsv_postwork (2092-2132) and stripe31 (src/aln2.cc:178-199), for simd = 2 | 3.  Blocks with
 * fewer than 8 query rows go to the scalar kernel forwardH_ng in the reference; that kernel is
 * not restated, so such calls set `unsupported`.
 * ========================================================================= */
#include <math.h>

typedef struct {
    const so_params_h* p;
    so_lsp_opts o;
    int32_t* skl;
    int cap, n;
    int unsupported;
} so_drvh;

static void drvh_write(so_drvh* d, int m, int n)
{
    if (d->n < d->cap) { d->skl[2 * d->n] = m; d->skl[2 * d->n + 1] = n; }
    ++d->n;
}

static void so_stripe31(so_task_h* t, int sh)
{
    if (sh < 0) {
        int am = t->a_right - t->a_left, bn = t->b_right - t->b_left;
        int shorter = am < bn ? am : bn;
        sh = -sh * shorter / 100;
    }
    sh *= 3;
    int up = t->b_right - 3 * t->a_right;
    int lw = t->b_left - 3 * t->a_left;
    if (up < lw) { int x = up; up = lw; lw = x; }
    up += sh; lw -= sh;
    int q;
    if ((q = t->b_right - 3 * t->a_left) < up) up = q;
    if ((q = t->b_left - 3 * t->a_right) > lw) lw = q;
    t->up = up; t->lw = lw;
}

static int gap_penalty_h(const so_params_h* p, int i)
{
    if (i == 0) return 0;
    return i > p->codonk1 ? p->lgop + i * p->lgep : p->gop + i * p->gep;
}
static int gap_ext_pen_h(const so_params_h* p, int i) { return i > p->codonk1 ? p->lgep : p->gep; }
static int unp_penalty3_h(const so_params_h* p, int i)
{
    /* PwdB::UnpPenalty3 (src/aln.h:289-301); beyond codonk1 (-yl3 only) is not restated */
    int unp = (i / 3) * p->gep;
    int egop = i % 3 == 1 ? p->gape1 : (i % 3 == 2 ? p->gape2 : 0);
    return unp + egop;
}

static int drvh_trcbk(so_drvh* d, const so_task_h* t)
{
    const int width = t->up - t->lw + 7;
    if (width < 0) return NEVSEL32_H;
    const int m = t->a_right - t->a_left;
    const int scalar = m < 8 || (d->o.alg & 3) == 0;
    if ((scalar && !d->p->ng) || t->b_right < t->b_left || t->a_left < 0 || t->b_left < 0 || t->b_right > t->b_len ||
        t->a_right > t->a_len) { d->unsupported = 1; return NEVSEL32_H; }
    int32_t score = 0;
    int room = d->cap > d->n ? d->cap - d->n : 0;
    if (scalar) {
        /* scalar kernel forwardH_ng (src/fwd2h1.cc:2007; every trace-back of -A0), restated in spaln_oracle_hng.c */
        int c = so_trcbk_h_ng(d->p, d->p->ng, t, &score, d->skl + 2 * (d->n < d->cap ? d->n : d->cap), room);
        if (c < 0) { d->unsupported = 1; return NEVSEL32_H; }
        d->n += c;
        return score;
    }
    int cnt = so_forward_h1_wip(d->p, t, 1, &score, d->skl + 2 * (d->n < d->cap ? d->n : d->cap), room);
    if (cnt < 0) { d->unsupported = 1; return NEVSEL32_H; }
    d->n += cnt;
    return score;
}

static int drvh_diagonal(so_drvh* d, const so_task_h* t)
{
    const so_params_h* p = d->p;
    const int LocalL = p->local && t->a_exgl && t->b_exgl;
    const int LocalR = p->local && t->a_exgr && t->b_exgr;
    int scr = 0, maxh = NEVSEL32_H;
    int mL = t->a_left, mR = t->a_right;
    for (int m = t->a_left, k = 0; m < t->a_right; ++k) {
        scr += p->simmtx[t->a[t->a_left + k] * p->simdim + t->b[t->b_left + 1 + 3 * k]] +
               SG(t, t->b_left + 1 + 3 * k, F_SIGE);
        ++m;
        if (LocalL && scr < 0) { scr = 0; mL = m; }
        if (LocalR && scr > maxh) { maxh = scr; mR = m; }
    }
    drvh_write(d, mL, 3 * (mL - t->a_left) + t->b_left);
    drvh_write(d, mR, 3 * (mR - t->a_left) + t->b_left);
    return LocalR ? maxh : scr;
}

static int drvh_lsp(so_drvh* d, so_task_h* t);

/* band of a block after a Hirschberg pass: re-derived from the ranges by the SIMD modes, the diagonal
 * bounds the pass recorded (cpos[.][8], [9]) under -A0 (src/fwd2h1.cc:2066-2072) */
static void drvh_window(so_drvh* d, so_task_h* t, const int32_t* row)
{
    if (d->o.alg & 3) so_stripe31(t, d->o.sh);
    else { t->lw = row[8]; t->up = row[9]; }
}

static void drvh_mimd_postwork(so_drvh* d, so_task_h* t, const int32_t* cpos, int n_imd)
{
    const int aleft = t->a_left, bleft = t->b_left;
    t->a_exgl = t->b_exgl = t->a_exgr = t->b_exgr = 0;
    int i = n_imd;
    while (--i >= 0 && cpos[10 * i] == END_OF_ULK_H) ;
    for ( ; i >= 0 && cpos[10 * i] != END_OF_ULK_H; --i) {
        int c = 0;
        t->a_left = cpos[10 * i + c];
        t->b_exgl = cpos[10 * i + (++c)];
        t->b_left = cpos[10 * i + (++c)];
        if (t->a_right > t->a_len || t->b_right > t->b_len || t->a_left < 0 || t->b_left < 0) return;
        if (t->b_left < 0 || t->b_left > t->b_right) break;
        while (cpos[10 * i + (++c)] < END_OF_ULK_H) drvh_write(d, t->a_left, cpos[10 * i + c]);
        drvh_window(d, t, cpos + 10 * (i + 1));
        drvh_trcbk(d, t);
        t->a_right = t->a_left;
        t->b_right = cpos[10 * i + c - 1];
    }
    if ((i < 0 && cpos[0] != END_OF_ULK_H) || cpos[2] != END_OF_ULK_H) {
        t->a_left = aleft;
        t->b_left = bleft;
        drvh_window(d, t, cpos);
        drvh_trcbk(d, t);
    }
}

static void drvh_rcsv_postwork(so_drvh* d, so_task_h* t, const int32_t* cpos)
{
    t->a_exgl = t->b_exgl = t->a_exgr = t->b_exgr = 0;
    int c = 0;
    if (cpos[c++] < END_OF_ULK_H) {
        while (cpos[++c] < END_OF_ULK_H) drvh_write(d, cpos[0], cpos[c]);
        const int aright = t->a_right, bright = t->b_right;
        t->a_right = cpos[0];
        t->b_right = cpos[c - 1];
        drvh_window(d, t, cpos);
        drvh_lsp(d, t);
        t->a_left = cpos[0];
        t->b_exgl = cpos[1];
        t->b_left = cpos[2];
        t->a_right = aright;
        t->b_right = bright;
        drvh_window(d, t, cpos + 10);
        drvh_lsp(d, t);
    } else if (d->p->local) {
        so_stripe31(t, d->o.sh);
        drvh_trcbk(d, t);
    }
}

static int drvh_lsp(so_drvh* d, so_task_h* t)
{
    const so_params_h* p = d->p;
    const int m = t->a_right - t->a_left;
    const int n = t->b_right - t->b_left;
    if (m < 0 || n < 0 || t->a_left < 0 || t->b_left < 0 || t->b_right > t->b_len || t->a_right > t->a_len) {
        /* a range a Hirschberg pass narrowed to outside the sequences (fhlastH1's start point
         * right of b_right): the reference reads foreign memory from here on */
        d->unsupported = 1;
        return NEVSEL32_H;
    }
    if (!m && !n) return 0;
    const int aexgl = t->a_exgl, aexgr = t->a_exgr, bexgl = t->b_exgl, bexgr = t->b_exgr;
    if (!m || !n) {
        drvh_write(d, t->a_left, t->b_left);
        drvh_write(d, t->a_right, t->b_right);
        if (m) return (aexgl || aexgr) ? gap_ext_pen_h(p, m) : gap_penalty_h(p, m);
        return (bexgl || bexgr) ? gap_ext_pen3(p, n) : unp_penalty3_h(p, n);
    }
    if (t->up == t->lw) return drvh_diagonal(d, t);
    if (abs(n - m) < NELEM || m == 1 || n <= 3) return drvh_trcbk(d, t);
    int n_imd = 1;
    int recursive = d->o.alg & 4;
    const int simd = d->o.alg & 3;
    if (simd == 1) { d->unsupported = 1; return NEVSEL32_H; }       /* -A1: kernels not restated */
    const float coef_B = 2.f;                           /* sizeof(short) */
    const float coef_C = (float) (((p->ng ? p->ng->noll : 2) + 1) * 4);    /* (Noll + 1) * sizeof(int) */
    float cvol = (float) m * (n + 3 * m);               /* rhombic, simd >= 2 */
    if (simd < 2) {                                     /* hexagonal (src/fwd2h1.cc:2161-2164) */
        const float k = (float) (t->lw - t->b_left + 3 * t->a_right);
        const float q = (float) (t->b_right - 3 * t->a_left - t->up);
        cvol = (float) m * n - (k * k + q * q) / 6;
    }
    if (coef_B * cvol < d->o.max_vmf_space) return drvh_trcbk(d, t);
    int imd_intvl = (m + 1) / 2;
    if (!recursive) {
        const double z = 2. * m * coef_B / coef_C;
        const int imd1 = (int) (pow(z, 1. / 3) + 0.5) - 1;
        const float spc = coef_C * n * imd1 + coef_B * cvol / (imd1 + 1) / (imd1 + 1);
        if (spc > d->o.max_vmf_space) recursive = 1;
        else {
            const int imd3 = m / NELEM;
            if (d->o.ubh) n_imd = d->o.ubh;
            else n_imd = imd1 < imd3 ? imd1 : imd3;
            imd_intvl = (m + n_imd) / (n_imd + 1);
            if (imd_intvl * n_imd == m) --n_imd;
            if (n_imd == 0) return drvh_trcbk(d, t);
        }
    }
    so_task_h saved = *t;
    int32_t* cpos = (int32_t*) malloc(sizeof(int32_t) * 10 * (n_imd + 1));
    int32_t ranges[4];
    int32_t scr = 0;
    const int rc = simd ? so_hirschberg_h1_wip(p, t, n_imd, &scr, cpos, ranges)
                        : so_hirschberg_h_ng(p, p->ng, t, n_imd, imd_intvl, &scr, cpos, ranges);
    if (rc < 0) { d->unsupported = 1; free(cpos); return NEVSEL32_H; }
    t->a_left = ranges[0]; t->a_right = ranges[1]; t->b_left = ranges[2]; t->b_right = ranges[3];
    if (scr > NEVSEL32_H) {
        if (cpos[0] == END_OF_ULK_H) {
            drvh_write(d, t->a_left, t->b_left);
            drvh_write(d, t->a_right, t->b_right);
        } else if (recursive)
            drvh_rcsv_postwork(d, t, cpos);
        else
            drvh_mimd_postwork(d, t, cpos, n_imd);
    }
    *t = saved;
    free(cpos);
    return scr;
}

int so_lsp_h(const so_params_h* p, const so_task_h* t0, const so_lsp_opts* o, int32_t* score,
             int32_t* skl, int cap, int* unsupported)
{
    so_drvh d;
    d.p = p; d.o = *o; d.skl = skl; d.cap = cap; d.n = 0; d.unsupported = 0;
    so_task_h t = *t0;
    *score = drvh_lsp(&d, &t);
    if (unsupported) *unsupported = d.unsupported;
    return d.n;
}
