/* TEST INFRASTRUCTURE ONLY -- see spaln_oracle.h.
 *
 * Scalar restatement of the reference's 16-lane int16 "strip" dynamic
 * programme.  The reference evaluates a strip of NELEM query rows per pass;
 * at step n lane k holds cell (row ml+1+k, column n-k).  Everything the
 * vector code does implicitly is spelled out per lane here, including what
 * happens in lanes that lie outside the matrix (they are evaluated with a
 * zero substitution score and zero splice signals and DO feed real cells),
 * the persistence of the diagonal-indexed band rows hv[]/fv[] between
 * strips, int16 saturation, and the re-basing schedule.
 */
#include "spaln_oracle.h"

#include <limits.h>
#include <stdlib.h>
#include <string.h>

#define NELEM 16
#define NP1 (NELEM + 1)

typedef int16_t var_t;

/* src/fwd2s1_simd.h:44,47,202 */
#define CHECK_SCR ((int) (0.9 * SHRT_MAX))
#define NEVSEL16 ((var_t) (SHRT_MIN + 1024))

/* TraceBackCode, src/rhomb_coord.h:36-61 */
enum { TB_STOP = 0, TB_DIAG = 1, TB_HORI = 2, TB_HORL = 3, TB_VERT = 8,
       TB_VERL = 9, TB_ACCR = 14, TB_NHOR = 16, TB_NVER = 32, TB_NHOL = 64,
       TB_NVEL = 128, TB_DONR = 128 };

/* _mm256_adds_epi16 / _mm256_subs_epi16, src/simd_functions.h:1025-1030 */
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

/* VecSub_c, src/simd_functions.h:116-125: whole vectors saturate, the scalar
 * tail (n % 16 elements) wraps */
static void vec_sub_c16(var_t* d, var_t c, int n)
{
    int nn = n / NELEM * NELEM, i = 0;
    for (; i < nn; ++i) d[i] = subs16(d[i], c);
    for (; i < n; ++i) d[i] = (var_t) (d[i] - c);
}

/* VecMax, src/simd_functions.h:128-151 (plain maximum; d[0] for n <= 0) */
static var_t vec_max16(const var_t* d, int n)
{
    var_t m = d[0];
    for (int i = 1; i < n; ++i) if (d[i] > m) m = d[i];
    return m;
}

/* SimdAln2s1::checkpoint, src/fwd2s1_simd.h:179-182 */
static int checkpoint(const so_params* p, int pv)
{
    return (CHECK_SCR - abs(pv)) / p->avmch / NELEM * NELEM;
}

typedef struct {
    const so_params* p;
    const so_task* t;
    int width, buf_size;
    var_t* vbuf;        /* Noll * buf_size */
    var_t *hv, *fv, *fv2;   /* diagonal-indexed band rows (biased pointers) */
    int dagp;
    var_t gn, ge, gn2, ge2, mil, ipen;
    var_t quant[SO_MAXQUANT], mean[SO_MAXQUANT];
} so_ctx;

static int ctx_init(so_ctx* c, const so_params* p, const so_task* t)
{
    memset(c, 0, sizeof(*c));
    c->p = p; c->t = t;
    c->width = t->up - t->lw + 3;
    c->buf_size = c->width + 2 * NELEM;         /* src/fwd2s1_simd.h:205 */
    c->dagp = p->noll == 3;
    size_t n = (size_t) p->noll * c->buf_size;
    c->vbuf = (var_t*) malloc(n * sizeof(var_t));
    if (!c->vbuf) return -1;
    for (size_t i = 0; i < n; ++i) c->vbuf[i] = NEVSEL16;   /* :297 */
    c->hv = c->vbuf - t->lw + 1;                /* :299-301 */
    c->fv = c->hv + c->buf_size;
    c->fv2 = c->dagp ? c->fv + c->buf_size : 0;
    /* Splat() narrows to short, src/fwd2s1_wip_simd.h:52-65 */
    c->ge = (var_t) p->gep;
    c->gn = (var_t) (p->gep + p->gop);
    c->ge2 = (var_t) p->lgep;
    c->gn2 = (var_t) (p->lgep + p->lgop);
    c->mil = (var_t) p->llmt;
    c->ipen = (var_t) (p->spj ? p->ipen : NEVSEL16);
    for (int j = 0; j < p->nquant && j < SO_MAXQUANT; ++j) {
        c->quant[j] = (var_t) p->quant_len[j];
        c->mean[j] = (var_t) p->quant_pen[j];
    }
    return 0;
}

/* SimdAln2s1::fhinitS1, mode <= 1 and no Vmf (src/fwd2s1_simd.cc:163-184) */
static void fhinit(so_ctx* c)
{
    const so_task* t = c->t;
    const so_params* p = c->p;
    var_t* hv = c->hv;
    const int rl = t->b_left - t->a_left;
    if (t->b_exgl)
        for (int r = t->lw; r < rl; ++r) hv[r] = 0;
    int rr = t->b_right - t->a_left;
    if (t->up < rr) rr = t->up;
    if (t->a_exgl) {
        for (int r = rl; r <= rr; ++r) hv[r] = 0;
    } else {
        int r = rl;
        hv[r++] = 0;
        hv[r] = (var_t) p->gappen1;
        if (p->gep) {
            int x = (NEVSEL16 - p->gop) / p->gep + rl;
            if (x < rr) rr = x;
            while (++r < rr) hv[r] = (var_t) (hv[r - 1] + p->gep);
        } else if (rr > r) {
            for (int i = r; i < rr; ++i) hv[i] = hv[r];
        }
    }
}

typedef struct { int val, mr, nr; } so_maxh;

/* SimdAln2s1::fhlastS1 (src/fwd2s1_simd.cc:241-262); vmax keeps the first
 * maximum and returns its argument for n <= 0 (src/clib.h:108-113) */
static int argvmax(const var_t* hv, int from, int n)
{
    int best = from;
    for (int i = 1; i < n; ++i) if (hv[from + i] > hv[best]) best = from + i;
    return best;
}

static void fhlast(const so_ctx* c, so_maxh* maxh)
{
    const so_task* t = c->t;
    const var_t* hv = c->hv;
    const int rr = t->b_right - t->a_right;
    int maxr = rr;
    if (t->a_exgr) {
        int r = t->lw > t->b_left - t->a_right ? t->lw : t->b_left - t->a_right;
        maxr = argvmax(hv, r, rr - r);
    }
    if (t->b_exgr) {
        int r = t->up - 1 < t->b_right - t->a_left ? t->up - 1 : t->b_right - t->a_left;
        int max_vert = argvmax(hv, rr, r - rr);
        if (hv[max_vert] > hv[maxr]) maxr = max_vert;
    }
    maxh->val = hv[maxr];
    if (maxr > rr) maxh->mr = t->b_right - maxr;
    else maxh->nr = t->a_right + maxr;
}

/* ---- Anti_rhomb_coord<CHAR>, step = 1 (src/rhomb_coord.h:65-235) -------- */
typedef struct {
    int m_base, n_base, m_width, n_width;
    uint8_t* bbuf;
    size_t size;
    int cur_m, cur_n;
    uint8_t* cur_p;
} so_trb;

static int trb_init(so_trb* tb, int mmax, int nmax, int mbase, int nbase)
{
    tb->m_base = mbase; tb->n_base = nbase;
    tb->m_width = mmax - mbase + 1;
    tb->n_width = nmax - nbase + 1 + tb->m_width;
    tb->size = (size_t) tb->m_width * tb->n_width + 32;
    tb->bbuf = (uint8_t*) calloc(tb->size + 64, 1);
    return tb->bbuf ? 0 : -1;
}
static void trb_initialize_m0(so_trb* tb, uint8_t dir)
{
    uint8_t* q = tb->bbuf;
    for (int n = 1; n < tb->n_width; ++n) *(q += tb->m_width) = dir;
}
static uint8_t* trb_set_point(so_trb* tb, int m, int n)
{
    tb->cur_m = m - tb->m_base;
    tb->cur_n = n - tb->n_base;
    tb->cur_p = tb->bbuf + (size_t) (tb->cur_m + tb->cur_n) * tb->m_width + tb->cur_m;
    return tb->cur_p;
}
static unsigned trb_to_left(so_trb* tb, int* m, int* n, int s)
{
    *m = tb->cur_m;
    *n = tb->cur_n -= s;
    if (*n < 0) { tb->cur_n = *n = 0; return 0; }
    tb->cur_p -= (size_t) s * tb->m_width;
    return *tb->cur_p;
}
static unsigned trb_to_upper(so_trb* tb, int* m, int* n, int s)
{
    *m = --tb->cur_m;
    *n = tb->cur_n -= s;
    if (*m < 0) {
        tb->cur_m = *m = 0;
        tb->cur_n = *n += s;
        return 0;
    } else if (*n < 0) {
        if (s > 0) tb->cur_m = *m -= *n / s;
        tb->cur_n = *n = 0;
        return 0;
    }
    tb->cur_p -= ((size_t) (1 + s) * tb->m_width + 1);
    return *tb->cur_p;
}
static int trb_go_back(so_trb* tb, unsigned code, int* m, int* n, unsigned* out)
{
    unsigned dir = code & 15;
    switch (dir) {
      case TB_STOP: break;
      case TB_DIAG:
        do {
            if (!(code = trb_to_upper(tb, m, n, 1))) { *out = 0; return 0; }
        } while ((code & 15) == TB_DIAG);
        break;
      case TB_HORI:
        while (!(code & TB_NHOR))
            if (!(code = trb_to_left(tb, m, n, 1))) { *out = 0; return 0; }
        code = trb_to_left(tb, m, n, 1);
        break;
      case TB_HORL:
        while (!(code & TB_NHOL))
            if (!(code = trb_to_left(tb, m, n, 1))) { *out = 0; return 0; }
        code = trb_to_left(tb, m, n, 1);
        break;
      case TB_VERT:
        while (!(code & TB_NVER))
            if (!(code = trb_to_upper(tb, m, n, 0))) { *out = 0; return 0; }
        code = trb_to_upper(tb, m, n, 0);
        break;
      case TB_VERL:
        while (!(code & TB_NVEL))
            if (!(code = trb_to_upper(tb, m, n, 0))) { *out = 0; return 0; }
        code = trb_to_upper(tb, m, n, 0);
        break;
      case TB_ACCR:
        do {
            if (!(code = trb_to_left(tb, m, n, 1))) { *out = 0; return 0; }
        } while (!(code & TB_DONR));
        break;
      default:
        return -1;      /* reference: fatal("Unexpected dir") */
    }
    *out = code;
    return 0;
}
static int trb_traceback(so_trb* tb, int m, int n, int32_t* skl, int cap)
{
    unsigned code = *trb_set_point(tb, m, n);
    int cnt = 0;
    m -= tb->m_base;
    n -= tb->n_base;
    while (code) {
        if (cnt < cap) { skl[2 * cnt] = m + tb->m_base; skl[2 * cnt + 1] = n + tb->n_base; }
        ++cnt;
        if (trb_go_back(tb, code, &m, &n, &code) < 0) return -2;
    }
    if (cnt < cap) { skl[2 * cnt] = m + tb->m_base; skl[2 * cnt + 1] = n + tb->n_base; }
    ++cnt;
    return cnt;
}

/* ---- one strip-step, shared by score-only and trace-back kernels -------- */
typedef struct {
    var_t HA[2][NP1];       /* hv_a[2] */
    var_t FA[NP1], F2A[NP1];/* fv_a, fv2_a */
    var_t S5[NP1], S3[NP1]; /* s5_a, s3_a */
    var_t PV[NELEM];        /* pv_a */
    var_t ev[NELEM], ev2[NELEM], hv2[NELEM], hil[NELEM], fv2r[NELEM];
} so_strip;

static void strip_reset(so_strip* s)
{
    /* vec_set(hv_a[0], nevsel, 4*Np1 + 2*nelem); vec_clear(s5_a, 2*Np1);
     * vec_clear(ps_a, 2*nelem)   (src/fwd2s1_wip_simd.h:81-83, 290-292) */
    for (int k = 0; k < NP1; ++k) {
        s->HA[0][k] = s->HA[1][k] = s->FA[k] = s->F2A[k] = NEVSEL16;
        s->S5[k] = s->S3[k] = 0;
    }
    for (int k = 0; k < NELEM; ++k) {
        s->PV[k] = 0;
        s->ev[k] = s->ev2[k] = s->hv2[k] = s->fv2r[k] = NEVSEL16;
        s->hil[k] = 0;
    }
}

/* trace == 0: scoreonlyS1_wip body; trace != 0: forwardS1_wip body, tb[k]
 * receives the 8-bit trace code of lane k.  Returns nothing; updates strip,
 * band rows and (LocalR) maxh. */
static void strip_step(so_ctx* c, so_strip* s, int ml, int j9, int n, int r, int p,
                       int LocalL_now, int LocalR, int accscr, so_maxh* maxh,
                       int trace, uint8_t* tb)
{
    const so_task* t = c->t;
    const so_params* P = c->p;
    const int q = 1 - p;
    const int j8 = j9 - 1;
    const int r0 = r - 2 * j8;
    const int kb = n - t->b_right > 0 ? n - t->b_right : 0;
    const int ke = j9 < n - t->b_left ? j9 : n - t->b_left;
    var_t Hleft[NELEM], Hup[NELEM], Fup[NELEM], F2up[NELEM], Hdg[NELEM];
    var_t s3[NELEM], s5[NELEM];

    for (int k = 0; k < NELEM; ++k) Hleft[k] = s->HA[q][k + 1];
    s->HA[q][0] = c->hv[r + 1];
    for (int k = 0; k < NELEM; ++k) Hup[k] = s->HA[q][k];
    s->FA[0] = c->fv[r + 1];
    for (int k = 0; k < NELEM; ++k) Fup[k] = s->FA[k];
    if (c->dagp) {
        s->F2A[0] = c->fv2[r + 1];
        for (int k = 0; k < NELEM; ++k) F2up[k] = s->F2A[k];
    }
    /* substitution scores of the lanes inside the matrix */
    if (kb) for (int k = 0; k < NELEM; ++k) s->PV[k] = 0;
    for (int k = kb; k < ke; ++k) {
        int ac = t->a[ml + k], bc = t->b[n - 1 - k];
        s->PV[k] = (var_t) P->simmtx[ac * P->simdim + bc];
    }
    s->HA[p][0] = c->hv[r];
    for (int k = 0; k < NELEM; ++k) Hdg[k] = s->HA[p][k];
    if (P->spj) {
        s->S3[0] = kb ? 0 : t->sig3[n];
        s->S5[0] = kb ? 0 : (var_t) (t->sig5[n] + c->ipen);
        for (int k = 0; k < NELEM; ++k) { s3[k] = s->S3[k]; s5[k] = s->S5[k]; }
        for (int k = 0; k < NELEM; ++k) { s->S3[k + 1] = s3[k]; s->S5[k + 1] = s5[k]; }
    }

    for (int k = 0; k < NELEM; ++k) {
        unsigned hb = 0, pb;
        var_t x, h, f, f2 = 0, e;
        /* horizontal (genome residue against a gap in the query) */
        x = adds16(Hleft[k], c->gn);
        e = adds16(s->ev[k], c->ge);
        if (!(e > x)) { e = x; hb |= TB_NHOR; }
        s->ev[k] = e;
        if (c->dagp) {
            var_t x2 = adds16(Hleft[k], c->gn2);
            var_t e2 = adds16(s->ev2[k], c->ge2);
            if (!(e2 > x2)) { e2 = x2; hb |= TB_NHOL; }
            s->ev2[k] = e2;
            if (!trace && e2 > e) { e = e2; s->ev[k] = e; } /* wip.h:108-109 */
        }
        /* vertical */
        f = adds16(Fup[k], c->ge);
        x = adds16(Hup[k], c->gn);
        if (!(f > x)) { f = x; hb |= TB_NVER; }
        if (c->dagp) {
            f2 = adds16(F2up[k], c->ge2);
            /* forwardS1_wip re-uses qv_v after it was overwritten with the
             * NVER flag (0 | 32), src/fwd2s1_wip_simd.h:333,339: the long
             * gap therefore opens from the flag value, not from H */
            x = trace ? adds16((hb & TB_NVER) ? TB_NVER : 0, c->gn2)
                      : adds16(Hup[k], c->gn2);
            if (!(f2 > x)) { f2 = x; hb |= TB_NVEL; }
            s->F2A[k + 1] = f2;
            if (!trace && f2 > f) f = f2;               /* wip.h:129-130 */
        }
        s->FA[k + 1] = f;
        /* diagonal */
        h = adds16(s->PV[k], Hdg[k]);
        pb = TB_DIAG;
        if (f > h) { h = f; pb = TB_VERT; }
        if (trace && c->dagp && f2 > h) { h = f2; pb = TB_VERL; }
        if (e > h) { h = e; pb = TB_HORI; }
        if (trace && c->dagp && s->ev2[k] > h) { h = s->ev2[k]; pb = TB_HORL; }
        /* acceptor */
        int acc = 0;
        if (P->spj) {
            var_t qv = adds16(s->hv2[k], s3[k]);
            var_t pen = c->mean[0];
            for (int j = 1; j < P->nquant; ++j)
                if (s->hil[k] > c->quant[j - 1]) pen = c->mean[j];
            qv = adds16(qv, pen);
            if (!(s->hil[k] > c->mil)) qv = NEVSEL16;
            if (qv > h) { h = qv; pb = TB_ACCR; acc = 1; }
        }
        if (LocalL_now && 0 > h) { h = 0; hb = 0; }
        s->HA[p][k + 1] = h;
        /* donor */
        if (P->spj) {
            var_t qv = adds16(h, s5[k]);
            if (trace && acc) qv = NEVSEL16;    /* no empty intron, wip.h:427-428 */
            if (qv > s->hv2[k]) {
                s->hv2[k] = qv;
                s->hil[k] = 0;
                if (trace) hb |= TB_DONR;
            }
            s->hil[k] = adds16(s->hil[k], 1);
        }
        if (tb) tb[k] = (uint8_t) ((hb | pb) & 0xff);
    }
    if (LocalR) {
        int best = 1;
        for (int k = 2; k <= j9; ++k) if (s->HA[p][k] > s->HA[p][best]) best = k;
        if (s->HA[p][best] + accscr > maxh->val) {
            maxh->val = s->HA[p][best] + accscr;
            maxh->mr = ml + best;
            maxh->nr = n - best + 1;
        }
    }
    if (j9 == ke && t->lw <= r0 && r0 <= t->up) {
        c->hv[r0] = s->HA[p][j9];
        c->fv[r0] = s->FA[j9];
        if (c->dagp) c->fv2[r0] = s->F2A[j9];
    }
}

static void rebase(so_ctx* c, int ml, int md, int* mc, int* accscr)
{
    const so_task* t = c->t;
    if (ml != *mc) return;
    var_t cmax = vec_max16(c->hv + t->lw, t->up - t->lw);
    int d = checkpoint(c->p, cmax);
    if (d < md / 2) {
        vec_sub_c16(c->hv + t->lw - 1, cmax, c->width);
        vec_sub_c16(c->fv + t->lw - 1, cmax, c->width);
        if (c->dagp) vec_sub_c16(c->fv2 + t->lw - 1, cmax, c->width);
        *accscr += cmax;
        *mc += md;
    } else
        *mc += d;
}

int so_scoreonly_wip(const so_params* p, const so_task* t, int32_t* score)
{
    so_ctx c;
    if (ctx_init(&c, p, t)) return -1;
    so_maxh maxh = { NEVSEL16, t->a_right, t->b_right };
    const int LocalL = p->local && t->a_exgl && t->b_exgl;
    const int LocalR = p->local && t->a_exgr && t->b_exgr;
    fhinit(&c);
    int accscr = 0;
    const int md = checkpoint(p, 0);
    int mc = md + t->a_left;
    so_strip s;
    for (int ml = t->a_left; ml < t->a_right; ml += NELEM) {
        const int j9 = NELEM < t->a_right - ml ? NELEM : t->a_right - ml;
        int n = t->b_left > t->lw + ml ? t->b_left : t->lw + ml;
        const int lim = t->b_right < t->up + (ml + j9) + 1 ? t->b_right : t->up + (ml + j9) + 1;
        const int n9 = lim + j9;
        int r = n - (ml + 1);
        strip_reset(&s);
        for (int ph = 0; n < n9; ++n, ++r, ph = 1 - ph)     /* n < n9, wip.h:91 */
            strip_step(&c, &s, ml, j9, n, r, ph, LocalL && !accscr, LocalR,
                       accscr, &maxh, 0, 0);
        rebase(&c, ml, md, &mc, &accscr);
    }
    if (!LocalR) {
        fhlast(&c, &maxh);
        maxh.val += accscr;
    }
    *score = maxh.val;
    free(c.vbuf);
    return 0;
}

int so_forward_wip(const so_params* p, const so_task* t, int32_t* score,
                   int32_t* skl, int cap, int64_t* n_cells)
{
    so_ctx c;
    so_trb trb;
    if (ctx_init(&c, p, t)) return -1;
    so_maxh maxh = { NEVSEL16, t->a_right, t->b_right };
    const int LocalL = p->local && t->a_exgl && t->b_exgl;
    const int LocalR = p->local && t->a_exgr && t->b_exgr;
    fhinit(&c);
    if (trb_init(&trb, t->a_right, t->b_right, t->a_left, t->b_left)) {
        free(c.vbuf);
        return -1;
    }
    if (!t->a_exgl) trb_initialize_m0(&trb, TB_HORI);
    int accscr = 0;
    int64_t cells = 0;
    const int md = checkpoint(p, 0);
    int mc = md + t->a_left;
    const int mw = t->a_right - t->a_left;
    const int mb = t->a_right - 2 * NELEM;
    const int mt = t->a_left + mw / NELEM * NELEM;
    so_strip s;
    uint8_t tb[NELEM];
    for (int ml = t->a_left; ml < t->a_right; ml += NELEM) {
        const int j9 = NELEM < t->a_right - ml ? NELEM : t->a_right - ml;
        int n = t->b_left > t->lw + ml ? t->b_left : t->lw + ml;
        const int lim = t->b_right < t->up + (ml + j9) + 1 ? t->b_right : t->up + (ml + j9) + 1;
        const int n9 = lim + j9;
        int r = n - (ml + 1);
        strip_reset(&s);
        for (int ph = 0; n <= n9; ++n, ++r, ph = 1 - ph) {  /* n <= n9, wip.h:300 */
            uint8_t* dst = trb_set_point(&trb, ml + 1, n);
            strip_step(&c, &s, ml, j9, n, r, ph, LocalL && !accscr, LocalR,
                       accscr, &maxh, 1, tb);
            cells += j9;
            /* wip.h:443-450: lanes beyond the last row are blanked in the
             * last (partial) strip; the 256-bit store writes the 16 codes
             * followed by 16 zero bytes; the last two strips OR instead */
            if (ml == mt)
                for (int k = t->a_right - mt; k < NELEM; ++k) tb[k] = 0;
            if (ml > mb) {
                for (int k = 0; k < NELEM; ++k) dst[k] |= tb[k];
            } else {
                memcpy(dst, tb, NELEM);
                memset(dst + NELEM, 0, NELEM);
            }
        }
        rebase(&c, ml, md, &mc, &accscr);
    }
    if (!LocalR) {
        fhlast(&c, &maxh);
        maxh.val += accscr;
    }
    int cnt = trb_traceback(&trb, maxh.mr, maxh.nr, skl, cap);
    *score = maxh.val;
    if (n_cells) *n_cells = cells;
    free(trb.bbuf);
    free(c.vbuf);
    return cnt;
}

/* =========================================================================
 * Unidirectional Hirschberg forward pass with quantised intron penalty:
 * SimdAln2s1::hirschbergS1_wip (src/fwd2s1_wip_simd.h:476-864), single affine.
 * Besides H/F/E every lane carries a link (`hc`: the diagonal where the path
 * crossed the previous intermediate row, or its start diagonal) and, in local
 * mode, the left-end row (`hb`).  At the n_im intermediate rows the links are
 * recorded (UdhIntermediate, src/udh_intermediate.h:29-67) and reset; the
 * back-walk at the end turns them into the crossing records cpos[][10].
 * ========================================================================= */
#define END_OF_ULK (INT_MAX - 2)            /* src/aln.h:49 */
#define NEVSEL32 (INT_MIN / 16 * 7)         /* src/cmn.h:79 */

typedef struct {
    int mi;
    int* buf;
    int* hlnk[2];
    int* vlnk[2];
} so_imd;

static int imd_init(so_imd* im, int mi, int lw, int width)
{
    const int nol = 2;
    size_t u_size = (size_t) nol * width;
    im->mi = mi;
    im->buf = (int*) malloc(2 * u_size * sizeof(int));
    if (!im->buf) return -1;
    for (size_t i = 0; i < 2 * u_size; ++i) im->buf[i] = END_OF_ULK;
    im->hlnk[0] = im->buf - lw + 1;
    im->vlnk[0] = im->hlnk[0] + u_size;
    im->hlnk[1] = im->hlnk[0] + width;
    im->vlnk[1] = im->vlnk[0] + width;
    return 0;
}

int so_hirschberg_wip(const so_params* p, const so_task* t, int n_im,
                      int32_t* score, int32_t* cpos /* (n_im+1) x 10 */,
                      int32_t* ranges /* a_left, a_right, b_left, b_right after the call */)
{
    if (p->noll != 2 || n_im < 1) return -3;
    so_ctx c;
    if (ctx_init(&c, p, t)) return -1;
    const int lw = t->lw, up = t->up, width = c.width;
    int a_left = t->a_left, a_right = t->a_right, b_left = t->b_left, b_right = t->b_right;
    const int LocalL = p->local && t->a_exgl && t->b_exgl;
    const int LocalR = p->local && t->a_exgr && t->b_exgr;
    /* band rows of the links: bbuf (hb, fb) and cbuf (hc, fc) */
    size_t nb = (size_t) 2 * c.buf_size;
    var_t* bbuf = (var_t*) malloc(nb * sizeof(var_t));
    int* cbuf = (int*) malloc(nb * sizeof(int));
    so_imd* imds = (so_imd*) calloc(n_im, sizeof(so_imd));
    if (!bbuf || !cbuf || !imds) return -1;
    var_t* hb = bbuf - lw + 1; var_t* fb = hb + c.buf_size;
    int* hc = cbuf - lw + 1; int* fc = hc + c.buf_size;
    memset(cbuf, 0, nb * sizeof(int));
    for (int i = 0; i <= n_im; ++i) cpos[10 * i + 0] = cpos[10 * i + 2] = END_OF_ULK;

    /* ---- fhinitS1, mode > 1 without Vmf (src/fwd2s1_simd.cc:163-238) */
    fhinit(&c);
    {
        const int ru = up + 2 * NELEM;
        const int rl = b_left - a_left;
        for (size_t i = 0; i < nb; ++i) bbuf[i] = (var_t) a_left;
        hc[rl] = rl;
        if (t->a_exgl) { for (int r = rl; r < ru; ++r) hc[r] = r; }
        else { for (int r = rl + 1; r <= ru; ++r) hc[r] = rl; }
        if (t->b_exgl) {
            int r = lw - 1;
            var_t m = (var_t) a_left;
            var_t* e = hb + rl;
            while (r < rl) { hc[r] = r; ++r; *e-- = m++; }
        } else {
            for (int r = lw - 1; r < rl; ++r) hc[r] = rl;
        }
        for (int i = 0; i < c.buf_size; ++i) fc[lw - 1 + i] = hc[lw - 1 + i];
    }

    /* ---- intermediates (src/fwd2s1_wip_simd.h:505-510; Udh_Imds ctor) */
    int mm = (a_right - a_left + n_im) / (n_im + 1);
    {
        int mi = a_left;
        for (int i = 0; i < n_im; ++i)
            if (imd_init(&imds[i], mi += mm, lw, width)) return -1;
    }
    so_imd* imd = &imds[0];
    mm = a_left + (imd->mi - a_left - 1) / NELEM * NELEM;
    int k9 = imd->mi - mm, k8 = k9 - 1;
    int rlst = INT_MAX;

    const int md = checkpoint(p, 0);
    int mc = a_left + md;
    int accscr = 0;
    struct { int val, ulk, ml, mr, nr; } maxh = { NEVSEL16, END_OF_ULK, a_left, a_right, b_right };

    /* lane state; the link lanes are NOT reset between strips in the reference */
    var_t HA[2][NP1], FA[NP1], S5[NP1], S3[NP1], PV[NELEM], PS[NELEM], PB[NELEM];
    var_t HBA[2][NP1], FBA[NP1];
    int HCA[2][NP1], FCA[NP1];
    memset(HCA, 0, sizeof(HCA)); memset(FCA, 0, sizeof(FCA));

    for (int ml = a_left, ii = 0; ml < a_right; ml += NELEM) {
        const int j9 = NELEM < a_right - ml ? NELEM : a_right - ml;
        const int j8 = j9 - 1;
        int n = b_left > lw + ml ? b_left : lw + ml;
        const int lim = b_right < up + (ml + j9) + 1 ? b_right : up + (ml + j9) + 1;
        const int n9 = lim + j9;
        int r = n - (ml + 1);
        int donor_r = r;
        for (int k = 0; k < NP1; ++k) {
            HA[0][k] = HA[1][k] = FA[k] = NEVSEL16;
            HBA[0][k] = HBA[1][k] = FBA[k] = 0;
            S5[k] = S3[k] = 0;
        }
        var_t ev[NELEM], eb[NELEM], hv2[NELEM], hb2[NELEM], hil[NELEM];
        int ec[NELEM], hc2[NELEM];
        for (int k = 0; k < NELEM; ++k) {
            PV[k] = 0; PS[k] = 0;
            ev[k] = NEVSEL16; eb[k] = 0; ec[k] = 0;
            hv2[k] = NEVSEL16; hb2[k] = 0; hc2[k] = 0; hil[k] = 0;
        }
        const int is_imd_ = ml == mm;

        for (int ph = 0; n < n9; ++n, ++r, ph = 1 - ph) {
            const int q = 1 - ph, pp = ph;
            const int r0 = r - 2 * j8;
            const int rj = r - 2 * k8;
            const int kb = n - b_right > 0 ? n - b_right : 0;
            const int ke = j9 < n - b_left ? j9 : n - b_left;
            const int is_imd = is_imd_ && rj >= lw && rj <= up;
            var_t Hleft[NELEM], Hup[NELEM], Fup[NELEM], Hdg[NELEM], s3[NELEM], s5[NELEM];
            var_t HBleft[NELEM], HBup[NELEM], FBup[NELEM], HBdg[NELEM];
            int HCleft[NELEM], HCup[NELEM], FCup[NELEM], HCdg[NELEM];

            for (int k = 0; k < NELEM; ++k) {
                Hleft[k] = HA[q][k + 1]; HBleft[k] = HBA[q][k + 1]; HCleft[k] = HCA[q][k + 1];
            }
            HA[q][0] = c.hv[r + 1]; if (LocalL) HBA[q][0] = hb[r + 1]; HCA[q][0] = hc[r + 1];
            FA[0] = c.fv[r + 1]; if (LocalL) FBA[0] = fb[r + 1]; FCA[0] = fc[r + 1];
            for (int k = 0; k < NELEM; ++k) {
                Hup[k] = HA[q][k]; HBup[k] = HBA[q][k]; HCup[k] = HCA[q][k];
                Fup[k] = FA[k]; FBup[k] = FBA[k]; FCup[k] = FCA[k];
            }
            if (kb) for (int k = 0; k < NELEM; ++k) PV[k] = 0;
            for (int k = kb; k < ke; ++k)
                PV[k] = (var_t) p->simmtx[t->a[ml + k] * p->simdim + t->b[n - 1 - k]];
            var_t pvv[NELEM];
            for (int k = 0; k < NELEM; ++k) pvv[k] = PV[k];
            HA[pp][0] = c.hv[r]; if (LocalL) HBA[pp][0] = hb[r]; HCA[pp][0] = hc[r];
            for (int k = 0; k < NELEM; ++k) { Hdg[k] = HA[pp][k]; HBdg[k] = HBA[pp][k]; HCdg[k] = HCA[pp][k]; }
            if (p->spj) {
                S3[0] = kb ? 0 : t->sig3[n];
                S5[0] = kb ? 0 : (var_t) (t->sig5[n] + c.ipen);
                for (int k = 0; k < NELEM; ++k) { s3[k] = S3[k]; s5[k] = S5[k]; }
                for (int k = 0; k < NELEM; ++k) { S3[k + 1] = s3[k]; S5[k + 1] = s5[k]; }
            }

            var_t hreg[NELEM], hbreg[NELEM];
            int hcreg[NELEM];
            var_t pbv[NELEM], accv[NELEM];
            for (int k = 0; k < NELEM; ++k) {
                var_t x, e, f, h, hbv;
                int hcv;
                /* horizontal */
                x = adds16(Hleft[k], c.gn);
                e = adds16(ev[k], c.ge);
                if (!(e > x)) { e = x; if (LocalL) eb[k] = HBleft[k]; ec[k] = HCleft[k]; }
                ev[k] = e;
                /* vertical */
                f = adds16(Fup[k], c.ge);
                x = adds16(Hup[k], c.gn);
                var_t fbv = FBup[k];
                int fcv = FCup[k];
                if (!(f > x)) { f = x; fbv = HBup[k]; fcv = HCup[k]; }
                FA[k + 1] = f; if (LocalL) FBA[k + 1] = fbv; FCA[k + 1] = fcv;
                /* diagonal, best of three */
                h = adds16(pvv[k], Hdg[k]); hbv = HBdg[k]; hcv = HCdg[k];
                var_t pb = 0;
                if (f > h) { h = f; hbv = fbv; hcv = fcv; pb = 2; }
                if (e > h) { h = e; hbv = eb[k]; hcv = ec[k]; pb = 1; }
                pbv[k] = pb;
                /* acceptor */
                accv[k] = 0;
                if (p->spj) {
                    var_t qv = adds16(hv2[k], s3[k]);
                    var_t pen = c.mean[0];
                    for (int j = 1; j < p->nquant; ++j)
                        if (hil[k] > c.quant[j - 1]) pen = c.mean[j];
                    qv = adds16(qv, pen);
                    if (!(hil[k] > c.mil)) qv = NEVSEL16;
                    if (qv > h) { h = qv; hbv = hb2[k]; hcv = hc2[k]; accv[k] = 1; }
                }
                if (LocalL && !accscr && 0 > h) h = 0;
                hreg[k] = h; hbreg[k] = hbv; hcreg[k] = hcv;
            }
            if (is_imd) for (int k = 0; k < NELEM; ++k) PV[k] = pbv[k];      /* Store(pv_a, pb_v) */
            if (p->spj && is_imd) {
                for (int k = 0; k < NELEM; ++k) PS[k] = accv[k];
                if (PS[k8]) {
                    imd->hlnk[0][rj] = donor_r;
                    imd->hlnk[1][rj] = donor_r + width;
                    rlst = rj;
                }
            }
            for (int k = 0; k < NELEM; ++k) {
                HA[pp][k + 1] = hreg[k];
                if (LocalL) HBA[pp][k + 1] = hbreg[k];
                HCA[pp][k + 1] = hcreg[k];
            }
            if (LocalL && !accscr) {
                for (int k = kb; k < ke; ++k) {
                    const int kp1 = k + 1;
                    if (HA[pp][kp1] == 0) {
                        HBA[pp][kp1] = (var_t) (ml + kp1);
                        HCA[pp][kp1] = r - 2 * k;
                    }
                }
            }
            if (LocalR) {
                int best = 1;
                for (int k = 2; k <= j9; ++k) if (HA[pp][k] > HA[pp][best]) best = k;
                if (HA[pp][best] + accscr > maxh.val) {
                    maxh.val = HA[pp][best] + accscr;
                    maxh.ml = HBA[pp][best];
                    maxh.ulk = HCA[pp][best];
                    maxh.mr = ml + best;
                    maxh.nr = n - best + 1;
                }
            }
            /* donor (registers hreg/hbreg/hcreg, not the patched lane arrays) */
            if (p->spj) {
                for (int k = 0; k < NELEM; ++k) {
                    var_t qv = adds16(hreg[k], s5[k]);
                    int don = qv > hv2[k];
                    if (don) {
                        hv2[k] = qv;
                        if (LocalL) hb2[k] = hbreg[k];
                        hc2[k] = hcreg[k];
                        hil[k] = 0;
                    }
                    hil[k] = adds16(hil[k], 1);
                    if (is_imd) PB[k] = (var_t) don;
                }
                if (is_imd && PB[k8]) donor_r = rj;
            }
            /* intermediate row */
            if (is_imd) {
                if (PV[k8] == 0) rlst = rj;
                if (PV[k8] == 1) imd->hlnk[0][rj] = rlst;
                imd->vlnk[0][rj] = HCA[pp][k9];
                HCA[pp][k9] = rj;
                imd->vlnk[1][rj] = FCA[k9];
                FCA[k9] = rj + width;
            }
            if (j9 == ke && lw <= r0 && r0 <= up) {
                c.hv[r0] = HA[pp][j9];
                if (LocalL) hb[r0] = HBA[pp][j9];
                hc[r0] = HCA[pp][j9];
                c.fv[r0] = FA[j9];
                if (LocalL) fb[r0] = FBA[j9];
                fc[r0] = FCA[j9];
            }
        }
        rebase(&c, ml, md, &mc, &accscr);
        if (is_imd_ && ++ii < n_im) {
            imd = &imds[ii];
            mm = a_left + (imd->mi - a_left - 1) / NELEM * NELEM;
            k9 = imd->mi - mm;
            k8 = k9 - 1;
        }
    }

    if (LocalR) {
        a_right = maxh.mr;
        b_right = maxh.nr;
    } else {
        /* fhlastS1 with mode > 1 (src/fwd2s1_simd.cc:241-262) */
        so_maxh mh = { maxh.val, maxh.mr, maxh.nr };
        const int rr = t->b_right - t->a_right;
        int maxr = rr;
        if (t->a_exgr) {
            int r = lw > t->b_left - t->a_right ? lw : t->b_left - t->a_right;
            maxr = argvmax(c.hv, r, rr - r);
        }
        if (t->b_exgr) {
            int r = up - 1 < t->b_right - t->a_left ? up - 1 : t->b_right - t->a_left;
            int max_vert = argvmax(c.hv, rr, r - rr);
            if (c.hv[max_vert] > c.hv[maxr]) maxr = max_vert;
        }
        mh.val = c.hv[maxr];
        if (maxr > rr) mh.mr = t->b_right - maxr;
        else mh.nr = t->a_right + maxr;
        maxh.val = mh.val + accscr; maxh.mr = mh.mr; maxh.nr = mh.nr;
        maxh.ulk = hc[maxr];
        maxh.ml = LocalL ? hb[maxr] : a_left;
        a_right = maxh.mr;
        b_right = maxh.nr;
    }

    /* ---- back-walk over the intermediates (src/fwd2s1_wip_simd.h:825-863) */
    int i = n_im;
    while (--i >= 0 && imds[i].mi > a_right) ;
    if (i < 0 && imds[0].mi > a_right) cpos[2] = b_right;
    int r = maxh.ulk;
    for ( ; i >= 0 && (imd = &imds[i])->mi > maxh.ml; --i) {
        int cc = 0, d = 0;
        for ( ; r >= up; r -= width) ++d;
        if (lw < imd->vlnk[d][r] && imd->vlnk[d][r] < up) {
            cpos[10 * i + cc++] = imd->mi;
            cpos[10 * i + cc++] = (d > 0) ? 1 : 0;
            for (int rp = imd->hlnk[d][r]; lw <= rp && rp < up && r != rp; rp = imd->hlnk[d][r = rp])
                cpos[10 * i + cc++] = r + imd->mi;
            cpos[10 * i + cc++] = r + imd->mi;
            cpos[10 * i + cc] = END_OF_ULK;
            r = imd->vlnk[d][r];
            if (r == END_OF_ULK) break;
        } else
            cpos[10 * i + 0] = END_OF_ULK;
    }
    for ( ; r > up; r -= width) ;
    if (LocalL) {
        a_left = maxh.ml;
        b_left = r + a_left;
    } else {
        const int rl = b_left - a_left;
        if (t->b_exgl && rl > r) {
            a_left = b_left - r;
            for (int j = 0; j < n_im && imds[j].mi < a_left; ++j) cpos[10 * j + 0] = END_OF_ULK;
        }
        if (t->a_exgl && rl < r) b_left = a_left + r;
    }
    ++i;
    int bad = 0;
    if (i >= 0 && i < n_im && imds[i].mi < a_left) bad = 1;
    if (!bad && cpos[10 * i + 2] < b_left) bad = 1;
    *score = bad ? NEVSEL32 : maxh.val;
    ranges[0] = a_left; ranges[1] = a_right; ranges[2] = b_left; ranges[3] = b_right;
    for (int k = 0; k < n_im; ++k) free(imds[k].buf);
    free(imds); free(bbuf); free(cbuf); free(c.vbuf);
    return 0;
}

/* =========================================================================
 * The DP driver: Aln2s1::lspS_ng (src/fwd2s1.cc:1801-1897) with its helpers
 * trcbkalignS_ng (1667-1710, SIMD branch), mimd_postwork (1714-1756),
 * rcsv_postwork (1758-1799), diagonalS_ng (1629-1665) and stripe
 * (src/aln2.cc:156-176), for simd = 2 | 3 (-A2 / -A3: `_wip` kernels, single affine gaps on the
 * Hirschberg route) and simd = 0 (-A0: scalar kernels, hirschbergS_ng, blocks banded by the
 * diagonal bounds the pass records).
 * Problems with fewer than 8 query rows go to the scalar exact-ILD kernel in
 * the reference (src/fwd2s1.cc:1676), restated in spaln_oracle_ng.c; without
 * its tables (so_params.penalty / sig53tab, so_task.int53) such calls set
 * `unsupported`.
 * ========================================================================= */
#include <math.h>

typedef struct {
    const so_params* p;
    so_lsp_opts o;
    int32_t* skl;
    int cap, n;
    int unsupported;
} so_drv;

static void drv_write(so_drv* d, int m, int n)
{
    if (d->n < d->cap) { d->skl[2 * d->n] = m; d->skl[2 * d->n + 1] = n; }
    ++d->n;
}

static void so_stripe(so_task* t, int sh)
{
    if (sh < 0) {
        int am = t->a_right - t->a_left, bn = t->b_right - t->b_left;
        int shorter = am < bn ? am : bn;
        sh = -sh * shorter / 100;
    }
    int up = t->b_right - t->a_right;
    int lw = t->b_left - t->a_left;
    if (up < lw) { int x = up; up = lw; lw = x; }
    up += sh; lw -= sh;
    int q;
    if ((q = t->b_right - t->a_left) < up) up = q;
    if ((q = t->b_left - t->a_right) > lw) lw = q;
    t->up = up; t->lw = lw;
}

/* PwdB::GapExtPen / GapPenalty / UnpPenalty (src/aln.h:275-287); codonk1 is LARGEN unless the gap
 * penalty is double affine; so_params.codonk1 == 0 (a fixture without it) means the same */
static int drv_k1(const so_params* p) { return (p->noll == 3 && p->codonk1 > 0) ? p->codonk1 : INT_MAX; }
static int gap_ext_pen(const so_params* p, int i) { return i > drv_k1(p) ? p->lgep : p->gep; }
static int gap_penalty(const so_params* p, int i)
{
    if (i == 0) return 0;
    return i > drv_k1(p) ? p->lgop + i * p->lgep : p->gop + i * p->gep;
}
static int unp_penalty(const so_params* p, int d)
{
    const int unp = d * p->gep;
    return d <= drv_k1(p) ? unp : unp + (p->lgep - p->gep) * (d - drv_k1(p));
}

static int drv_trcbk(so_drv* d, const so_task* t)
{
    const int width = t->up - t->lw + 3;
    if (width < 0) return NEVSEL32;
    const int m = t->a_right - t->a_left;
    int32_t score = 0;
    int room = d->cap > d->n ? d->cap - d->n : 0;
    if (m < 8 || (d->o.alg & 3) == 0) {
        /* scalar exact-ILD kernel (src/fwd2s1.cc:1676; every trace-back of -A0): needs the intron tables */
        int c = so_trcbk_ng(d->p, t, &score, d->skl + 2 * (d->n < d->cap ? d->n : d->cap), room);
        if (c < 0) { d->unsupported = 1; return NEVSEL32; }
        d->n += c;
        return score;
    }
    int cnt = so_forward_wip(d->p, t, &score, d->skl + 2 * (d->n < d->cap ? d->n : d->cap), room, 0);
    if (cnt < 0) { d->unsupported = 1; return NEVSEL32; }
    d->n += cnt;
    return score;
}

static int drv_diagonal(so_drv* d, const so_task* t)
{
    const so_params* p = d->p;
    const int LocalL = p->local && t->a_exgl && t->b_exgl;
    const int LocalR = p->local && t->a_exgr && t->b_exgr;
    const int dlt = p->local ? 0 : ((t->b_right - t->b_left) - (t->a_right - t->a_left));
    /* dlt < 0 swaps the roles of a and b */
    const uint8_t* as = dlt < 0 ? t->b : t->a;
    const uint8_t* bs = dlt < 0 ? t->a : t->b;
    int al = dlt < 0 ? t->b_left : t->a_left, ar = dlt < 0 ? t->b_right : t->a_right;
    int bl = dlt < 0 ? t->a_left : t->b_left;
    int mL = al, mR = ar;
    int scr = 0, maxh = NEVSEL32;
    for (int m = al, k = 0; m++ < ar; ++k) {
        int x = as[al + k], y = bs[bl + k];
        scr += dlt < 0 ? p->simmtx[y * p->simdim + x] : p->simmtx[x * p->simdim + y];
        if (LocalL && scr < 0) { scr = 0; mL = m; }
        if (LocalR && scr > maxh) { maxh = scr; mR = m; }
    }
    int r = bl - al;
    if (dlt < 0) r -= dlt;
    drv_write(d, mL, mL + r);
    drv_write(d, mR, mR + r);
    return LocalR ? maxh : scr;
}

static int drv_lsp(so_drv* d, so_task* t);

/* band of a block after a Hirschberg pass: re-derived from the ranges by the SIMD modes, the lowest /
 * highest diagonal the pass recorded (cpos[.][8], [9]) under -A0 (src/fwd2s1.cc:1736-1741) */
static void drv_window(so_drv* d, so_task* t, const int32_t* row)
{
    if (d->o.alg & 3) so_stripe(t, d->o.sh);
    else { t->lw = row[8]; t->up = row[9]; }
}

static void drv_mimd_postwork(so_drv* d, so_task* t, const int32_t* cpos, int n_imd)
{
    const int aleft = t->a_left, bleft = t->b_left;
    t->a_exgl = t->b_exgl = t->a_exgr = t->b_exgr = 0;
    int i = n_imd;
    while (--i >= 0 && cpos[10 * i] == END_OF_ULK) ;
    for ( ; i >= 0 && cpos[10 * i] != END_OF_ULK; --i) {
        int c = 0;
        t->a_left = cpos[10 * i + c];
        t->b_exgl = cpos[10 * i + (++c)];
        t->b_left = cpos[10 * i + (++c)];
        if (t->b_left < 0 || t->b_left > t->b_right) break;
        while (cpos[10 * i + (++c)] < END_OF_ULK) drv_write(d, t->a_left, cpos[10 * i + c]);
        drv_window(d, t, cpos + 10 * (i + 1));
        drv_trcbk(d, t);
        t->a_right = t->a_left;
        t->b_right = cpos[10 * i + c - 1];
    }
    if ((i < 0 && cpos[0] != END_OF_ULK) || cpos[2] != END_OF_ULK) {
        t->a_left = aleft;
        t->b_left = bleft;
        drv_window(d, t, cpos);
        drv_trcbk(d, t);
    }
}

static void drv_rcsv_postwork(so_drv* d, so_task* t, const int32_t* cpos)
{
    t->a_exgl = t->b_exgl = t->a_exgr = t->b_exgr = 0;
    int c = 0;
    if (cpos[c++] < END_OF_ULK) {
        while (cpos[++c] < END_OF_ULK) drv_write(d, cpos[0], cpos[c]);
        const int aright = t->a_right, bright = t->b_right;
        t->a_right = cpos[0];
        t->b_right = cpos[c - 1];
        drv_window(d, t, cpos);
        drv_lsp(d, t);
        t->a_left = cpos[0];
        t->b_exgl = cpos[1];
        t->b_left = cpos[2];
        t->a_right = aright;
        t->b_right = bright;
        drv_window(d, t, cpos + 10);
        drv_lsp(d, t);
    } else if (d->p->local) {
        so_stripe(t, d->o.sh);
        drv_trcbk(d, t);
    }
}

static int drv_lsp(so_drv* d, so_task* t)
{
    const so_params* p = d->p;
    const int m = t->a_right - t->a_left;
    const int n = t->b_right - t->b_left;
    if (!m && !n) return 0;
    const int aexgl = t->a_exgl, aexgr = t->a_exgr, bexgl = t->b_exgl, bexgr = t->b_exgr;
    if (!m || !n) {
        drv_write(d, t->a_left, t->b_left);
        drv_write(d, t->a_right, t->b_right);
        if (m) return (aexgl || aexgr) ? gap_ext_pen(p, m) : gap_penalty(p, m);
        return (bexgl || bexgr) ? gap_ext_pen(p, n) : unp_penalty(p, n);
    }
    if (t->up == t->lw) return drv_diagonal(d, t);
    if (abs(n - m) < 8 || m == 1 || n == 1) return drv_trcbk(d, t);
    int n_imd = 1;
    int recursive = d->o.alg & 4;
    const float coef_B = 2.f, coef_C = (float) ((p->noll + 1) * 4);
    const int simd = d->o.alg & 3;
    if (simd == 1) { d->unsupported = 1; return NEVSEL32; }    /* -A1: kernels not restated */
    float cvol = (float) m * (n + m);                   /* rhombic, simd >= 2 */
    if (simd < 2) {                                     /* hexagonal (src/fwd2s1.cc:1830-1833) */
        const float k = (float) (t->lw - t->b_left + t->a_right);
        const float q = (float) (t->b_right - t->a_left - t->up);
        cvol = (float) m * n - (k * k + q * q) / 2;
    }
    if (coef_B * cvol < d->o.max_vmf_space) return drv_trcbk(d, t);
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
            if (n_imd == 0) return drv_trcbk(d, t);
        }
    }
    so_task saved = *t;
    int32_t* cpos = (int32_t*) malloc(sizeof(int32_t) * 10 * (n_imd + 1));
    int32_t ranges[4];
    int32_t scr = 0;
    const int rc = simd ? so_hirschberg_wip(p, t, n_imd, &scr, cpos, ranges)
                        : so_hirschberg_ng(p, t, n_imd, imd_intvl, &scr, cpos, ranges);
    if (rc < 0) { d->unsupported = 1; free(cpos); return NEVSEL32; }
    t->a_left = ranges[0]; t->a_right = ranges[1]; t->b_left = ranges[2]; t->b_right = ranges[3];
    if (scr > NEVSEL32) {
        if (cpos[0] == END_OF_ULK) {
            drv_write(d, t->a_left, t->b_left);
            drv_write(d, t->a_right, t->b_right);
        } else if (recursive)
            drv_rcsv_postwork(d, t, cpos);
        else
            drv_mimd_postwork(d, t, cpos, n_imd);
    }
    *t = saved;
    free(cpos);
    return scr;
}

int so_lsp(const so_params* p, const so_task* t0, const so_lsp_opts* o, int32_t* score,
           int32_t* skl, int cap, int* unsupported)
{
    so_drv d;
    d.p = p; d.o = *o; d.skl = skl; d.cap = cap; d.n = 0; d.unsupported = 0;
    so_task t = *t0;
    *score = drv_lsp(&d, &t);
    if (unsupported) *unsupported = d.unsupported;
    return d.n;
}
