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
