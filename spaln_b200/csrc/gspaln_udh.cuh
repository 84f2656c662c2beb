// gspaln_udh.cuh -- unidirectional-Hirschberg forward pass on sm_100a.
//
// Semantics: SimdAln2s1::hirschbergS1_wip of the reference
// (src/fwd2s1_wip_simd.h:476-864; link initialisation src/fwd2s1_simd.cc:205-238;
// intermediates src/udh_intermediate.h:29-88) for single affine gaps, global,
// semi-global and local (-LS: left-end lanes `hb`, running maximum) modes.  No trace matrix is written: every cell carries a
// link (the diagonal on which its path crossed the previous intermediate row,
// or started) next to H, F, E and the best-donor value.  At the n_imd
// intermediate rows the links are recorded into per-problem arrays and reset;
// a back-walk over those arrays produces the crossing records (`Dim10 cpos[]`)
// from which the host driver cuts the problem into blocks (mimd_postwork).
//
// Mapping: same systolic scheme as gspaln_kernels.cuh (anti-diagonal evaluation
// inside a thread, strips chained through the diagonal-indexed band buffer),
// with NRU = 4 rows per thread (4 threads per strip, 8 strips per pass) to keep
// the doubled per-row state in registers.  The band buffer carries a second
// word pair {H link, F link} per diagonal.
//
// Two classes: one warp per problem (a CTA runs four problems), and -- for long
// queries in batches too small to keep every warp busy, where the largest
// problem on one warp would be the whole launch -- a CTA per problem (NW = 4:
// 32 strips per pass on one chain, the CTA barrier per step; see team_sync).
// The host chooses per batch (gspaln.cu, plan_batch).
#pragma once
#include "gspaln_kernels.cuh"

namespace gspaln {

constexpr int NRU = 4;                  // strip rows per thread in the UDH kernel
constexpr int TPSU = NELEM / NRU;       // threads per strip
constexpr int SPPU = 32 / TPSU;         // strips per pass
constexpr int END_OF_ULK = INT_MAX - 2; // src/aln.h:49
constexpr int NEVSEL32 = INT_MIN / 16 * 7;  // NEVSEL, src/cmn.h:79

// link arrays of the scalar passes (gspaln_xudh.cuh, gspaln_hxudh.cuh), intermediate i of a problem:
// hlnk | vlnk | lwrb | uprb, each noll x width ints, indexed by diagonal from lw - 1
// (UdhIntermediate with bounds, src/udh_intermediate.h:29-66)
struct UxImd {
    int* base; int width, noll, lw;
    __device__ __forceinline__ int& at(int i, int which, int k, int r) const
    {
        return base[((long long) (4 * i + which) * noll + k) * width + (r - (lw - 1))];
    }
};

// per-row event bits handed from the cell loop to the intermediate-row logic
enum : unsigned { EV_HORI = 1, EV_VERT = 2, EV_ACC = 4, EV_DON = 8 };

// left-end rows of local alignments (`hb` lanes of the reference), LOCAL only
struct LocalLanes {
    int BA[NRU], BB[NRU], FB[NRU], EB[NRU], B2[NRU];
};

struct LocalStep {              // per-step inputs / outputs of the local-mode extras
    int up_b, up_fb, up_db;     // hb of the row above: H (prev step), F, H (two steps ago)
    bool localL_now;            // LocalL && !accscr
    bool track;                 // LocalR
    int row_first;              // absolute query row of this thread's row 0 (m coordinate)
    int diag0;                  // diagonal of this thread's row 0 at this step
    int j9_left;                // number of real rows of the strip still below row0 (j9 - row0)
    int bv, bk, bml, bulk;      // best cell of the step: value, row, left end, link
};

template <bool SPJ, bool LOCAL>
__device__ __forceinline__ void strip_step_udh(
    int (&HO)[NRU], const int (&HN)[NRU], int (&F)[NRU], int (&E)[NRU],
    int (&V2)[NRU], int (&NJ)[NRU],
    int (&CO)[NRU], const int (&CN)[NRU], int (&FC)[NRU], int (&EC)[NRU], int (&C2)[NRU],
    int (&BO)[NRU], const int (&BN)[NRU], LocalLanes& LL, LocalStep& ls,
    const int (&arow)[NRU],
    const char* __restrict__ ring_hi, const char* __restrict__ mtx_bytes,
    const int2* __restrict__ pen_tab, int pen_cap, int step,
    int up_h, int up_f, int up_d, int up_c, int up_fc, int up_dc,
    int gn, int ge, unsigned& events)
{
    events = 0u;
    if (LOCAL) { ls.bv = INT_MIN; ls.bk = 0; ls.bml = 0; ls.bulk = 0; }
#pragma unroll
    for (int k = NRU - 1; k >= 0; --k) {
        const RingEntry re = *reinterpret_cast<const RingEntry*>(
            ring_hi - k * (CTA_THREADS * (int) sizeof(RingEntry)));
        const int left = HN[k], lc = CN[k];
        const int uh = k ? HN[k ? k - 1 : 0] : up_h;
        const int uf = k ? F[k ? k - 1 : 0] : up_f;
        const int dg = k ? HO[k ? k - 1 : 0] : up_d;
        const int uc = k ? CN[k ? k - 1 : 0] : up_c;
        const int ufc = k ? FC[k ? k - 1 : 0] : up_fc;
        const int dc = k ? CO[k ? k - 1 : 0] : up_dc;
        unsigned ev = 0;
        int lb = 0, ub = 0, ufb = 0, db = 0;
        if (LOCAL) {
            lb = BN[k];
            ub = k ? BN[k ? k - 1 : 0] : ls.up_b;
            ufb = k ? LL.FB[k ? k - 1 : 0] : ls.up_fb;
            db = k ? BO[k ? k - 1 : 0] : ls.up_db;
        }
        // horizontal
        int x = satlo(left + gn);
        int e = satlo(E[k] + ge);
        if (!(e > x)) { e = x; EC[k] = lc; if (LOCAL) LL.EB[k] = lb; }
        E[k] = e;
        // vertical
        int f = satlo(uf + ge);
        x = satlo(uh + gn);
        int fcl = ufc, fbl = ufb;
        if (!(f > x)) { f = x; fcl = uc; fbl = ub; }
        F[k] = f; FC[k] = fcl;
        if (LOCAL) LL.FB[k] = fbl;
        // diagonal, best of three
        const int pv = *reinterpret_cast<const int*>(mtx_bytes + re.prof + arow[k]);
        int h = sat16(pv + dg);
        int hc = dc, hbv = db;
        if (f > h) { h = f; hc = fcl; hbv = fbl; ev = EV_VERT; }
        if (e > h) { h = e; hc = EC[k]; if (LOCAL) hbv = LL.EB[k]; ev = EV_HORI; }
        if (SPJ) {
            const int q0 = sat16(V2[k] + re.s3);
            const int2 pq = pen_tab[min(step + NJ[k], pen_cap)];
            const int q = min(max(q0 + pq.x, pq.y), 32767);
            if (q > h) { h = q; hc = C2[k]; if (LOCAL) hbv = LL.B2[k]; ev |= EV_ACC; }
        }
        if (LOCAL && ls.localL_now && h < 0) h = 0;
        if (SPJ) {
            // donor (no empty-intron guard in the Hirschberg pass); links as of before the
            // local restart patch below
            const int qd = sat16(h + re.s5);
            if (qd > V2[k]) { V2[k] = qd; C2[k] = hc; if (LOCAL) LL.B2[k] = hbv; NJ[k] = -step; ev |= EV_DON; }
        }
        if (LOCAL) {
            // a cell of score 0 inside the matrix starts a new local alignment
            // (src/fwd2s1_wip_simd.h:719-727)
            const bool inside = re.pad != 0 && k < ls.j9_left;
            if (ls.localL_now && inside && h == 0) { hbv = (short) (ls.row_first + k); hc = ls.diag0 - 2 * k; }
            if (ls.track && k < ls.j9_left && h >= ls.bv) { ls.bv = h; ls.bk = k; ls.bml = hbv; ls.bulk = hc; }
            BO[k] = hbv;
        }
        HO[k] = h;
        CO[k] = hc;
        events |= ev << (4 * k);
    }
}

struct UdhBest { int val, ml, ulk, mr, nr; };     // Rvulmn of the reference (src/fwd2s1_simd.h:49-55)

struct UdhTaskView {
    int* bandc;         // {H link, F link} per diagonal (int2), buf_size entries
    int* bandb;         // {H left end, F left end} per diagonal (int2), LOCAL only
    int* imd;           // n_imd x 4 x width: hlnk0, hlnk1, vlnk0, vlnk1
    int n_imd, n_active, mm0;
};

// NW = warps per problem.  NW == 1: every warp of the CTA runs its own problem, SPPU strips per
// pass.  NW == WARPS_PER_CTA (the "team" class, queries of >= UDH_TEAM_ROWS rows): the CTA runs
// one problem with NW * SPPU strips per pass on one systolic chain -- the strips of different
// warps meet only through the band buffer, exactly as the strips of one warp do, so the only
// change is that the per-step barrier is the CTA's; `tsm` is the team's scratch in shared memory.
template <int NW>
__device__ __forceinline__ void team_sync() { if (NW == 1) __syncwarp(); else __syncthreads(); }

template <bool SPJ, bool LOCAL, int NW>
__device__ void run_pass_udh(const DevParams& P, const SmemLayout& sm,
                             const DevTask& t, const UdhTaskView& uv,
                             const unsigned char* __restrict__ aseq,
                             const ColInfo* __restrict__ cols, unsigned* band,
                             int ml0, int nstr, int& rlst_io,
                             bool localL, bool localL_now, bool localR, int accscr, UdhBest& wbest, int* tsm)
{
    const int lane = NW == 1 ? (threadIdx.x & 31) : (int) threadIdx.x;     // lane of the problem's chain
    const int sidx = lane / TPSU;
    const int sub = lane % TPSU;
    const int row0 = sub * NRU;
    const StripGeom g = strip_geom<false>(t, ml0 + NELEM * sidx);    // `n < n9`
    const int j8 = g.j9 - 1;
    const int nsteps = g.n_last - g.n_start + 1;
    const bool live = sidx < nstr && nsteps > 0;
    const int width = t.up - t.lw + 3;

    int n_start0;
    if (NW == 1) n_start0 = __shfl_sync(0xffffffffu, g.n_start, 0);
    else {
        if (lane == 0) tsm[0] = g.n_start;
        __syncthreads();
        n_start0 = tsm[0];
    }
    const int off = (g.n_start - n_start0) + (NELEM - 1 + LAG) * sidx;
    int niter = live ? off + nsteps : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) niter = max(niter, __shfl_xor_sync(0xffffffffu, niter, o));
    if (NW > 1) {
        if ((lane & 31) == 0) tsm[1 + (lane >> 5)] = niter;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < NW; ++w) niter = max(niter, tsm[1 + w]);
    }
    if (niter == 0) return;

    // Which intermediate row (if any) lies in this strip?  mi_i = a_left + mm0 (i + 1).
    // The reference advances to the next intermediate only after the strip holding the
    // current one, so only the strictly increasing prefix (n_active) is ever visited.
    int imd_i = -1, k8 = -1;
    if (live && uv.n_active > 0) {
        const int lo = g.ml + 1 - t.a_left;                 // first row of the strip, relative
        int i = (lo + uv.mm0 - 1) / uv.mm0 - 1;             // smallest i with mi_i >= first row
        if (i < 0) i = 0;
        const int mi = t.a_left + uv.mm0 * (i + 1);
        if (i < uv.n_active && mi >= g.ml + 1 && mi <= g.ml + NELEM && mi <= t.a_right) {
            imd_i = i;
            k8 = mi - g.ml - 1;                             // strip row of the intermediate
        }
    }
    const bool has_imd = imd_i >= 0 && k8 >= row0 && k8 < row0 + NRU;
    const int k8l = k8 - row0;
    int* hl0 = uv.imd + (long long) max(imd_i, 0) * 4 * width - (t.lw - 1);   // index by diagonal
    int* hl1 = hl0 + width;
    int* vl0 = hl0 + 2 * width;
    int* vl1 = hl0 + 3 * width;
    int donor_r = g.n_start - (g.ml + 1);
    int rlst = rlst_io;

    int HA[NRU], HB[NRU], F[NRU], E[NRU], V2[NRU], NJ[NRU], arow[NRU];
    int CA[NRU], CB[NRU], FC[NRU], EC[NRU], C2[NRU];
    LocalLanes LL;
    LocalStep ls;
    ls.localL_now = localL_now; ls.track = localR;
    ls.row_first = g.ml + row0 + 1;
    ls.j9_left = g.j9 - row0;
    ls.up_b = ls.up_fb = ls.up_db = 0; ls.diag0 = 0;
    int prev_ub = 0;
    int bval = INT_MIN, bstep = 0, bk = 0, bml = 0, bulk = 0;
#pragma unroll
    for (int k = 0; k < NRU; ++k) {
        HA[k] = NEV; HB[k] = NEV; F[k] = NEV; E[k] = NEV; V2[k] = NEV; NJ[k] = 0;
        CA[k] = 0; CB[k] = 0; FC[k] = 0; EC[k] = 0; C2[k] = 0;
        LL.BA[k] = 0; LL.BB[k] = 0; LL.FB[k] = 0; LL.EB[k] = 0; LL.B2[k] = 0;
        arow[k] = (live && row0 + k < g.j9) ? 4 * (int) aseq[(g.ml - t.a_left) + row0 + k] : 4 * ZROW;
    }
    const int gn = P.gn, ge = P.ge;
    int prev_uh = NEV, prev_uc = 0;
    const int band_bias = g.ml + t.lw - 1;
    RingEntry* ring = sm.ring + threadIdx.x;
    const char* mtx_bytes = reinterpret_cast<const char*>(sm.mtx);
    const int ipen = P.ipen;
    const bool owns_bottom = live && j8 >= row0 && j8 < row0 + NRU;
    const int kbot = j8 - row0;
    const int2* bandc = reinterpret_cast<const int2*>(uv.bandc);
    const int2* bandb = reinterpret_cast<const int2*>(uv.bandb);

    auto col_fetch = [&](int c) -> uint2 {
        if (c >= t.b_left && c <= t.b_right)
            return __ldg(reinterpret_cast<const uint2*>(cols + (c - t.b_left)));
        return make_uint2(0u, 0xffffffffu);
    };
    auto col_decode = [&](uint2 ci, int c, bool with_sig) -> RingEntry {
        RingEntry re;
        re.pad = 0;
        re.prof = ZROW * (MTX_LD * 4);
        re.s3 = 0; re.s5 = 0;
        if (ci.y != 0xffffffffu) {
            if (c > t.b_left) { re.prof = (int) (ci.y & 0xffu) * (MTX_LD * 4); re.pad = 1; }
            if (SPJ && with_sig) {
                re.s3 = hi16(ci.x);
                re.s5 = (int) (short) (lo16(ci.x) + ipen);
            }
        }
        return re;
    };

    unsigned nxt_band = 0;
    int2 nxt_bandc = make_int2(0, 0), nxt_bandb = make_int2(0, 0);
    uint2 nxt_col = make_uint2(0u, 0xffffffffu);

    for (int i = -1; i < niter; ++i) {
        const int j = i - off;
        const int sh_h = __shfl_up_sync(0xffffffffu, (i & 1) ? HA[NRU - 1] : HB[NRU - 1], 1);
        const int sh_f = __shfl_up_sync(0xffffffffu, F[NRU - 1], 1);
        const int sh_c = __shfl_up_sync(0xffffffffu, (i & 1) ? CA[NRU - 1] : CB[NRU - 1], 1);
        const int sh_fc = __shfl_up_sync(0xffffffffu, FC[NRU - 1], 1);
        int sh_b = 0, sh_fb = 0;
        if (LOCAL) {
            sh_b = __shfl_up_sync(0xffffffffu, (i & 1) ? LL.BA[NRU - 1] : LL.BB[NRU - 1], 1);
            sh_fb = __shfl_up_sync(0xffffffffu, LL.FB[NRU - 1], 1);
        }
        if (live && j == -1) {
            if (sub == 0) {
                nxt_band = __ldcg(band + (g.n_start - band_bias));
                nxt_bandc = __ldcg(bandc + (g.n_start - band_bias));
                prev_uh = lo16(__ldcg(band + (g.n_start - 1 - band_bias)));
                prev_uc = __ldcg(bandc + (g.n_start - 1 - band_bias)).x;
                if (LOCAL && localL) {
                    nxt_bandb = __ldcg(bandb + (g.n_start - band_bias));
                    prev_ub = __ldcg(bandb + (g.n_start - 1 - band_bias)).x;
                }
            }
            nxt_col = col_fetch(g.n_start);
#pragma unroll 1
            for (int d = 1; d < NELEM; ++d) {
                const int c = g.n_start - d;
                const RingEntry re = col_decode(col_fetch(c), c, false);
                ring[(c & 15) * CTA_THREADS] = re;
                ring[((c & 15) + 16) * CTA_THREADS] = re;
            }
        } else if (live && j >= 0 && j < nsteps) {
            const int n = g.n_start + j;
            const unsigned cur_band = nxt_band;
            const int2 cur_bandc = nxt_bandc, cur_bandb = nxt_bandb;
            const RingEntry cur_col = col_decode(nxt_col, n, n <= t.b_right);
            if (j + 1 < nsteps) {
                if (sub == 0) {
                    nxt_band = __ldcg(band + (n + 1 - band_bias));
                    nxt_bandc = __ldcg(bandc + (n + 1 - band_bias));
                    if (LOCAL && localL) nxt_bandb = __ldcg(bandb + (n + 1 - band_bias));
                }
                nxt_col = col_fetch(n + 1);
            }
            const int slot = n & 15;
            ring[slot * CTA_THREADS] = cur_col;
            ring[(slot + 16) * CTA_THREADS] = cur_col;
            const char* ring_hi = reinterpret_cast<const char*>(ring + (slot + 16 - row0) * CTA_THREADS);
            int up_h, up_f, up_c, up_fc;
            if (sub == 0) {
                up_h = lo16(cur_band); up_f = hi16(cur_band);
                up_c = cur_bandc.x; up_fc = cur_bandc.y;
            } else {
                up_h = sh_h; up_f = sh_f; up_c = sh_c; up_fc = sh_fc;
            }
            const int up_d = prev_uh, up_dc = prev_uc;
            prev_uh = up_h; prev_uc = up_c;
            if (LOCAL) {
                // hb lanes are only maintained when both left ends are free (LocalL)
                if (sub == 0) { ls.up_b = cur_bandb.x; ls.up_fb = cur_bandb.y; }
                else { ls.up_b = sh_b; ls.up_fb = sh_fb; }
                ls.up_db = prev_ub;
                prev_ub = ls.up_b;
                ls.diag0 = (n - row0) - (g.ml + row0 + 1);
            }
            unsigned events;
            if (i & 1)
                strip_step_udh<SPJ, LOCAL>(HB, HA, F, E, V2, NJ, CB, CA, FC, EC, C2, LL.BB, LL.BA, LL, ls,
                                           arow, ring_hi, mtx_bytes, sm.pen, P.pen_cap, j,
                                           up_h, up_f, up_d, up_c, up_fc, up_dc, gn, ge, events);
            else
                strip_step_udh<SPJ, LOCAL>(HA, HB, F, E, V2, NJ, CA, CB, FC, EC, C2, LL.BA, LL.BB, LL, ls,
                                           arow, ring_hi, mtx_bytes, sm.pen, P.pen_cap, j,
                                           up_h, up_f, up_d, up_c, up_fc, up_dc, gn, ge, events);
            if (LOCAL && localR && ls.bv > bval) {      // strictly greater: earlier steps keep ties
                bval = ls.bv; bstep = n; bk = row0 + ls.bk; bml = ls.bml; bulk = ls.bulk;
            }
            // ---- intermediate row (src/fwd2s1_wip_simd.h:694-705, 760-773)
            if (has_imd) {
                const int rj = (n - k8) - (g.ml + k8 + 1);
                if (rj >= t.lw && rj <= t.up) {
                    const unsigned ev = (events >> (4 * k8l)) & 15u;
                    if (ev & EV_ACC) { hl0[rj] = donor_r; hl1[rj] = donor_r + width; rlst = rj; }
                    if (ev & EV_DON) donor_r = rj;
                    if (!(ev & (EV_HORI | EV_VERT))) rlst = rj;         // diagonal
                    if (ev & EV_HORI) hl0[rj] = rlst;
                    int hcv = 0, fcv = 0;
#pragma unroll
                    for (int k = 0; k < NRU; ++k)
                        if (k == k8l) { hcv = (i & 1) ? CB[k] : CA[k]; fcv = FC[k]; }
                    vl0[rj] = hcv;
                    vl1[rj] = fcv;
#pragma unroll
                    for (int k = 0; k < NRU; ++k)
                        if (k == k8l) {
                            if (i & 1) CB[k] = rj; else CA[k] = rj;
                            FC[k] = rj + width;
                        }
                }
            }
            if (owns_bottom) {
                int out_h = (i & 1) ? HB[NRU - 1] : HA[NRU - 1];
                int out_f = F[NRU - 1];
                int out_c = (i & 1) ? CB[NRU - 1] : CA[NRU - 1];
                int out_fc = FC[NRU - 1];
                int out_b = (i & 1) ? LL.BB[NRU - 1] : LL.BA[NRU - 1];
                int out_fb = LL.FB[NRU - 1];
                if (kbot != NRU - 1) {
#pragma unroll
                    for (int k = 0; k < NRU - 1; ++k)
                        if (k == kbot) {
                            out_h = (i & 1) ? HB[k] : HA[k]; out_f = F[k];
                            out_c = (i & 1) ? CB[k] : CA[k]; out_fc = FC[k];
                            out_b = (i & 1) ? LL.BB[k] : LL.BA[k]; out_fb = LL.FB[k];
                        }
                }
                const int cb = n - j8;
                const int r0 = cb - (g.ml + g.j9);
                if (cb > t.b_left && r0 >= t.lw && r0 <= t.up) {
                    __stcg(band + (r0 - t.lw + 1), pack16(out_h, out_f));
                    __stcg(reinterpret_cast<int2*>(uv.bandc) + (r0 - t.lw + 1), make_int2(out_c, out_fc));
                    if (LOCAL && localL)
                        __stcg(reinterpret_cast<int2*>(uv.bandb) + (r0 - t.lw + 1), make_int2(out_b, out_fb));
                }
            }
        } else {
            prev_uh = NEV; prev_uc = 0; prev_ub = 0;
        }
        team_sync<NW>();
    }
    if (LOCAL && localR) {
        // reference order: strips ascending, then step, then lane (first max)
        int bv = (live && bval > INT_MIN) ? bval : INT_MIN;
        int bs = bstep, bkk = bk, bst = sidx, bm = bml, bu = bulk;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const int ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int os = __shfl_xor_sync(0xffffffffu, bs, o);
            const int ok = __shfl_xor_sync(0xffffffffu, bkk, o);
            const int ot = __shfl_xor_sync(0xffffffffu, bst, o);
            const int om = __shfl_xor_sync(0xffffffffu, bm, o);
            const int ou = __shfl_xor_sync(0xffffffffu, bu, o);
            const bool take = ov > bv || (ov == bv && (ot < bst || (ot == bst && (os < bs || (os == bs && ok < bkk)))));
            if (take) { bv = ov; bs = os; bkk = ok; bst = ot; bm = om; bu = ou; }
        }
        if (NW > 1) {
            // the warps of the team, same order
            int* b = tsm + 8 + 6 * (lane >> 5);
            if ((lane & 31) == 0) { b[0] = bv; b[1] = bs; b[2] = bkk; b[3] = bst; b[4] = bm; b[5] = bu; }
            __syncthreads();
            bv = tsm[8]; bs = tsm[9]; bkk = tsm[10]; bst = tsm[11]; bm = tsm[12]; bu = tsm[13];
#pragma unroll
            for (int w = 1; w < NW; ++w) {
                const int* c = tsm + 8 + 6 * w;
                const int ov = c[0], os = c[1], ok = c[2], ot = c[3];
                const bool take = ov > bv || (ov == bv && (ot < bst || (ot == bst && (os < bs || (os == bs && ok < bkk)))));
                if (take) { bv = ov; bs = os; bkk = ok; bst = ot; bm = c[4]; bu = c[5]; }
            }
        }
        if (bv > INT_MIN && bv + accscr > wbest.val) {
            wbest.val = bv + accscr;
            wbest.ml = bm; wbest.ulk = bu;
            wbest.mr = ml0 + NELEM * bst + bkk + 1;
            wbest.nr = bs - bkk;
        }
    }
    // rlst after this pass: the value left by the thread that handled the LAST
    // intermediate of the pass (intermediates of one pass run concurrently; the
    // reference runs them in order -- see DESIGN.md for the tie this leaves open)
    int who = has_imd ? lane : -1;
#pragma unroll
    for (int o = 16; o; o >>= 1) who = max(who, __shfl_xor_sync(0xffffffffu, who, o));
    if (NW == 1) {
        if (who >= 0) rlst_io = __shfl_sync(0xffffffffu, rlst, who);
    } else {
        const int r = __shfl_sync(0xffffffffu, rlst, max(who, 0) & 31);
        if ((lane & 31) == 0) { tsm[32 + 2 * (lane >> 5)] = who; tsm[33 + 2 * (lane >> 5)] = r; }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < NW; ++w)
            if (tsm[32 + 2 * w] >= 0) rlst_io = tsm[33 + 2 * w];     // ascending: the last one stays
    }
}

struct DevUdhOut {                      // per problem
    int score, status;
    int a_left, a_right, b_left, b_right;
    int pad0, pad1;
};

constexpr int UDH_TEAM_ROWS = 512;      // queries with at least this many rows: a CTA (team of warps) per problem

template <bool SPJ, bool LOCAL, int NW = 1>
__global__ void __launch_bounds__(CTA_THREADS, 3)
dp_udh_kernel(const DevParams* __restrict__ gP, const int2* __restrict__ gpen,
              const DevTask* __restrict__ tasks, const int* __restrict__ order, int ntasks, int* ticket,
              const unsigned char* __restrict__ apool, const ColInfo* __restrict__ cpool,
              unsigned* bandpool, long long band_slab, int* udhpool, long long udh_slab,
              int* cpospool, DevUdhOut* results, const int* ready)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ DevParams sP;
    {
        const int* src = reinterpret_cast<const int*>(gP);
        int* dst = reinterpret_cast<int*>(&sP);
        for (int i = threadIdx.x; i < (int) (sizeof(DevParams) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const DevParams& P = sP;
    SmemLayout sm;
    sm.ring = reinterpret_cast<RingEntry*>(smem_raw);
    int2* spen = reinterpret_cast<int2*>(smem_raw + sizeof(RingEntry) * RING * CTA_THREADS);
    for (int i = threadIdx.x; i <= P.pen_cap; i += blockDim.x) spen[i] = gpen[i];
    sm.pen = spen;
    sm.mtx = sP.mtxT;
    __syncthreads();

    static_assert(NW == 1 || NW == WARPS_PER_CTA, "a team is the whole CTA");
    constexpr int NT = 32 * NW;                             // lanes per problem
    __shared__ int s_team[48];
    int* tsm = s_team;
    const int lane = NW == 1 ? (threadIdx.x & 31) : (int) threadIdx.x;
    const long long wslot = (long long) blockIdx.x * WARPS_PER_CTA + (NW == 1 ? (threadIdx.x >> 5) : 0);
    unsigned* band = bandpool + wslot * band_slab;
    int* uslab = udhpool + wslot * udh_slab;

    for (;;) {
        int tk = 0;
        if (NW == 1) {
            if (lane == 0) tk = atomicAdd(ticket, 1);
            tk = __shfl_sync(0xffffffffu, tk, 0);
        } else {
            __syncthreads();
            if (lane == 0) tsm[44] = atomicAdd(ticket, 1);
            __syncthreads();
            tk = tsm[44];
        }
        if (tk >= ntasks) break;
        const int ti = order[tk];
        const DevTask t = tasks[ti];
        if (t.kind != 2 || ((t.flags & 32) != 0) != (NW > 1)) continue;      // another kernel / class runs it
        bool arrived = wait_inputs(ready, tk);
        if (NW > 1) arrived = __syncthreads_and(arrived) != 0;
        if (!arrived) {
            if (lane == 0) {
                // inputs never arrived: no crossing records (the driver skips the post-work)
                DevUdhOut r; memset(&r, 0, sizeof(r)); r.status = 4; r.score = INT_MIN / 16 * 7; results[ti] = r;
                int* cp = cpospool + t.pad1;
                for (int i = 0; i <= t.pad0; ++i) cp[10 * i] = cp[10 * i + 2] = INT_MAX - 2;
            }
            continue;
        }
        const unsigned char* aseq = apool + t.a_off;
        const ColInfo* cols = cpool + t.col_off;
        const int width = t.up - t.lw + 3;
        const int buf_size = width + 2 * NELEM;
        const bool a_exgl = t.flags & 1, a_exgr = t.flags & 2, b_exgl = t.flags & 4, b_exgr = t.flags & 8;
        const int n_imd = t.pad0;
        int* cpos = cpospool + t.pad1;

        const bool LocalL = LOCAL && a_exgl && b_exgl;
        const bool LocalR = LOCAL && a_exgr && b_exgr;
        UdhTaskView uv;
        uv.bandc = uslab;
        uv.bandb = uslab + 2 * (long long) ((buf_size + 1) & ~1);
        uv.imd = uslab + 4 * (long long) ((buf_size + 1) & ~1);
        uv.n_imd = n_imd;
        uv.mm0 = (t.a_right - t.a_left + n_imd) / (n_imd + 1);
        {
            // strictly increasing prefix of the strips that hold an intermediate row
            int na = 0, prev = INT_MIN;
            for (int i = 0; i < n_imd; ++i) {
                const int mi = t.a_left + uv.mm0 * (i + 1);
                const int mm = t.a_left + (mi - t.a_left - 1) / NELEM * NELEM;
                if (mm <= prev || mm >= t.a_right) break;
                prev = mm; ++na;
            }
            uv.n_active = na;
        }

        // ---- fhinitS1: scores (src/fwd2s1_simd.cc:163-184) and links (205-238)
        for (int i = lane; i < buf_size; i += NT) band[i] = pack16(NEV, NEV);
        for (long long i = lane; i < (long long) n_imd * 4 * width; i += NT) uv.imd[i] = END_OF_ULK;
        for (int i = lane; i <= n_imd; i += NT) { cpos[10 * i + 0] = END_OF_ULK; cpos[10 * i + 2] = END_OF_ULK; }
        team_sync<NW>();
        {
            const int rl = t.b_left - t.a_left;
            const int ru = t.up + 2 * NELEM;
            int rr = t.b_right - t.a_left;
            if (t.up < rr) rr = t.up;
            if (b_exgl)
                for (int r = t.lw + lane; r < rl; r += NT) band[r - t.lw + 1] = pack16(0, NEV);
            if (a_exgl) {
                for (int r = rl + lane; r <= rr; r += NT) band[r - t.lw + 1] = pack16(0, NEV);
            } else if (lane == 0) {
                int r = rl;
                int v = 0;
                band[r - t.lw + 1] = pack16(0, NEV);
                ++r;
                v = (short) P.gappen1;
                band[r - t.lw + 1] = pack16(v, NEV);
                if (P.gep) {
                    int x = (NEV - P.gop) / P.gep + rl;
                    if (x < rr) rr = x;
                    while (++r < rr) { v = (short) (v + P.gep); band[r - t.lw + 1] = pack16(v, NEV); }
                } else {
                    for (int q = r; q < rr; ++q) band[q - t.lw + 1] = pack16(v, NEV);
                }
            }
            // links: hc[r] = own diagonal along free ends, else the corner diagonal; fc = hc
            int2* bc = reinterpret_cast<int2*>(uv.bandc);
            for (int i = lane; i < buf_size; i += NT) {
                const int r = t.lw - 1 + i;
                int v;
                if (r < rl) v = b_exgl ? r : rl;
                else if (r == rl) v = rl;
                else v = a_exgl ? (r < ru ? r : rl) : rl;
                if (r > ru) v = 0;
                bc[i] = make_int2(v, v);
                if (LOCAL) {
                    // left-end rows: a.left everywhere; along the free left edge of b the row itself
                    int hbv = (short) t.a_left;
                    if (b_exgl && r >= t.lw && r <= rl) hbv = (short) (t.a_left + (rl - r));
                    reinterpret_cast<int2*>(uv.bandb)[i] = make_int2(hbv, (short) t.a_left);
                }
            }
        }
        __threadfence_block();
        team_sync<NW>();

        int accscr = 0;
        const int md = checkpoint(P.avmch, 0);
        int mc = md + t.a_left;
        int rlst = INT_MAX;
        UdhBest wbest{NEV, END_OF_ULK, t.a_left, t.a_right, t.b_right};
        int ml0 = t.a_left;
        // strips per pass: the chain's length, evened out over the passes the query needs
        int sppu = SPPU * NW;
        if (NW > 1) {
            const int strips = (t.a_right - t.a_left + NELEM - 1) / NELEM;
            const int passes = (strips + sppu - 1) / sppu;
            sppu = (strips + passes - 1) / passes;
        }
        while (ml0 < t.a_right) {
            int nstr = min(sppu, (t.a_right - ml0 + NELEM - 1) / NELEM);
            if (mc >= ml0 && mc < ml0 + nstr * NELEM && ((mc - ml0) % NELEM) == 0)
                nstr = (mc - ml0) / NELEM + 1;
            run_pass_udh<SPJ, LOCAL, NW>(P, sm, t, uv, aseq, cols, band, ml0, nstr, rlst,
                                         LocalL, LocalL && !accscr, LocalR, accscr, wbest, tsm);
            const int last_ml = ml0 + (nstr - 1) * NELEM;
            if (last_ml == mc) {
                team_sync<NW>();
                const int nmax = t.up - t.lw;
                int cm = lo16(__ldcg(band + 1));
                for (int i = lane; i < nmax; i += NT) cm = max(cm, lo16(__ldcg(band + 1 + i)));
#pragma unroll
                for (int o = 16; o; o >>= 1) cm = max(cm, __shfl_xor_sync(0xffffffffu, cm, o));
                if (NW > 1) {
                    if ((lane & 31) == 0) tsm[40 + (lane >> 5)] = cm;
                    __syncthreads();
#pragma unroll
                    for (int w = 0; w < NW; ++w) cm = max(cm, tsm[40 + w]);
                }
                const int d = checkpoint(P.avmch, cm);
                if (d < md / 2) {
                    const int nn = width / NELEM * NELEM;
                    for (int i = lane; i < width; i += NT) {
                        const unsigned w = __ldcg(band + i);
                        int h = lo16(w) - cm, f = hi16(w) - cm;
                        if (i < nn) { h = sat16(h); f = sat16(f); }
                        else { h = (short) h; f = (short) f; }
                        __stcg(band + i, pack16(h, f));
                    }
                    accscr += cm;
                    mc += md;
                } else
                    mc += d;
                team_sync<NW>();
            }
            ml0 += nstr * NELEM;
        }
        __threadfence_block();
        team_sync<NW>();

        // ---- fhlastS1 with links (src/fwd2s1_simd.cc:241-262) ...
        const int rr = t.b_right - t.a_right;
        int maxr = rr;
        auto argmax = [&](int from, int n) -> int {
            int bv = INT_MIN, bi = INT_MAX;
            for (int i = lane; i < n; i += 32) {
                const int v = lo16(__ldcg(band + (from + i - t.lw + 1)));
                if (v > bv) { bv = v; bi = from + i; }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const int ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            return n <= 0 ? from : bi;
        };
        if (!LocalR && (NW == 1 || threadIdx.x < 32)) {     // (a team leaves the end point to its first warp)
            if (a_exgr) {
                const int r = max(t.lw, t.b_left - t.a_right);
                maxr = argmax(r, rr - r);
            }
            if (b_exgr) {
                const int r = min(t.up - 1, t.b_right - t.a_left);
                const int mv = argmax(rr, r - rr);
                if (lo16(__ldcg(band + (mv - t.lw + 1))) > lo16(__ldcg(band + (maxr - t.lw + 1)))) maxr = mv;
            }
        }
        if (lane == 0) {
            // ... and the back-walk over the intermediates (src/fwd2s1_wip_simd.h:812-863)
            const int lw = t.lw, up = t.up;
            int a_left = t.a_left, a_right = t.a_right, b_left = t.b_left, b_right = t.b_right;
            int val, r, maxh_ml;
            if (LocalR) {
                val = wbest.val; r = wbest.ulk; maxh_ml = wbest.ml;
                a_right = wbest.mr; b_right = wbest.nr;
            } else {
                val = lo16(__ldcg(band + (maxr - lw + 1))) + accscr;
                if (maxr > rr) a_right = t.b_right - maxr; else b_right = t.a_right + maxr;
                r = __ldcg(reinterpret_cast<const int2*>(uv.bandc) + (maxr - lw + 1)).x;
                maxh_ml = LocalL ? __ldcg(reinterpret_cast<const int2*>(uv.bandb) + (maxr - lw + 1)).x : a_left;
            }
            auto MI = [&](int i) { return t.a_left + uv.mm0 * (i + 1); };
            auto L = [&](int i, int which, int d, int rr_) -> int& {
                return uv.imd[(long long) i * 4 * width + (which * 2 + d) * width + (rr_ - (lw - 1))];
            };
            int i = n_imd;
            while (--i >= 0 && MI(i) > a_right) ;
            if (i < 0 && MI(0) > a_right) cpos[2] = b_right;
            for ( ; r < END_OF_ULK && i >= 0 && MI(i) > maxh_ml; --i) {
                int cc = 0, d = 0;
                for ( ; r >= up; r -= width) ++d;
                const int v = L(i, 1, d, r);
                if (lw < v && v < up) {
                    cpos[10 * i + cc++] = MI(i);
                    cpos[10 * i + cc++] = d > 0 ? 1 : 0;
                    for (int rp = L(i, 0, d, r); lw <= rp && rp < up && r != rp; rp = L(i, 0, d, r = rp))
                        if (cc < 9) cpos[10 * i + cc++] = r + MI(i);
                    if (cc < 9) cpos[10 * i + cc++] = r + MI(i);
                    cpos[10 * i + cc] = END_OF_ULK;
                    r = L(i, 1, d, r);
                    if (r == END_OF_ULK) break;
                } else
                    cpos[10 * i + 0] = END_OF_ULK;
            }
            for ( ; r > up; r -= width) ;
            if (LocalL) {
                a_left = maxh_ml;
                b_left = r + a_left;
            } else {
                const int rl = b_left - a_left;
                if (b_exgl && rl > r) {
                    a_left = b_left - r;
                    for (int jx = 0; jx < n_imd && MI(jx) < a_left; ++jx) cpos[10 * jx + 0] = END_OF_ULK;
                }
                if (a_exgl && rl < r) b_left = a_left + r;
            }
            ++i;
            bool bad = false;
            if (i >= 0 && i < n_imd && MI(i) < a_left) bad = true;
            if (!bad && cpos[10 * i + 2] < b_left) bad = true;
            DevUdhOut o;
            o.score = bad ? NEVSEL32 : val;
            o.status = 0;
            o.a_left = a_left; o.a_right = a_right; o.b_left = b_left; o.b_right = b_right;
            o.pad0 = o.pad1 = 0;
            results[ti] = o;
        }
        __syncwarp();
    }
}

}   // namespace gspaln
