// gspaln_xudh.cuh -- the scalar unidirectional-Hirschberg pass (exact intron lengths) as a warp kernel.
//
// Semantics: bit-identical to Aln2s1::hirschbergS_ng (src/fwd2s1.cc:764-1104, with hinitS_ng /
// hlastS_ng 701-762 and the intermediates of src/udh_intermediate.h:29-88 in their bounded form),
// the Hirschberg pass of the reference's default mode `-A0`.  No path records: every cell state
// carries {value, highest / lowest diagonal visited since the last intermediate row, start row,
// link to the previous intermediate}; at the intermediate rows the links and bounds are recorded
// and restarted, and a back-walk turns them into the crossing records `Dim10 cpos[]` whose entries
// [8], [9] band the blocks of the post-work (src/fwd2s1.cc:1736-1741).
//
// Mapping: one warp per problem, the 32 lanes own 32 consecutive query rows and sweep them as an
// anti-diagonal wavefront (lane l on column s - l at step s).  The band rows H, F, F2 are the
// reference's diagonal-indexed arrays, kept in global memory (L2) and SHARED by the lanes: row
// m + 1 touches diagonal r two steps after row m did, so with one __syncwarp() per step every load
// sees exactly the history the reference's row-by-row loop would see -- including entries no row
// overwrote.  What a row carries along its columns (the horizontal gap states, the post-splice
// flags, the donor list with its bounds and links) lives in the registers of its lane.  A pass
// holds at most one intermediate row: the "last diagonal on the row" register of the reference
// runs along that row and is handed from pass to pass.
//
// NW = warps per problem.  NW == 1: a CTA runs NG_WARPS independent problems, one per warp.  NW > 1
// (queries with hundreds of rows): the wavefront is 32 NW rows tall, one problem per CTA, and the
// per-step barrier is the CTA's -- the hazard argument above is about steps, not about warps.
#pragma once
#include "gspaln_ng.cuh"
#include "gspaln_udh.cuh"

namespace gspaln {

struct UxCell { int val, upr, lwr, ml, ulk; };          // Rvwml, src/aln.h:130-136
struct __align__(16) UxSlot { int v[8]; };              // a band entry: UxCell padded to 32 bytes

__device__ __forceinline__ UxCell ux_ld(const UxSlot* p)
{
    const int4 a = __ldcg(reinterpret_cast<const int4*>(p));    // through L2: written by other lanes
    const int u = __ldcg(reinterpret_cast<const int*>(p) + 4);
    return UxCell{a.x, a.y, a.z, a.w, u};
}
__device__ __forceinline__ void ux_st(UxSlot* p, const UxCell& c)
{
    *reinterpret_cast<int4*>(p) = make_int4(c.val, c.upr, c.lwr, c.ml);
    reinterpret_cast<int*>(p)[4] = c.ulk;
}

// donor list of one row (see NgList): value, column, gap state | 5' code << 4, and what the path
// carries across the intron
struct UxList {
    int val[NG_NCAND + 1], jnc[NG_NCAND + 1], inf[NG_NCAND + 1];
    int upr[NG_NCAND + 1], lwr[NG_NCAND + 1], ml[NG_NCAND + 1], ulk[NG_NCAND + 1];
    int n;
    __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int l = 0; l <= NG_NCAND; ++l) {
            val[l] = NG_NEVSEL; jnc[l] = 0; inf[l] = 0; upr[l] = INT_MIN; lwr[l] = INT_MAX; ml[l] = 0; ulk[l] = END_OF_ULK;
        }
        n = 0;
    }
    __device__ __forceinline__ bool insert(int x, int j, int info, const UxCell& c, int link)
    {
        if (n > NG_NCAND) n = NG_NCAND;
        int pos = 0;
#pragma unroll
        for (int l = 0; l < NG_NCAND; ++l)
            if (l < n && val[l] >= x) ++pos;
        if (pos >= NG_NCAND) return false;
#pragma unroll
        for (int l = NG_NCAND; l > 0; --l)
            if (l > pos) {
                val[l] = val[l - 1]; jnc[l] = jnc[l - 1]; inf[l] = inf[l - 1];
                upr[l] = upr[l - 1]; lwr[l] = lwr[l - 1]; ml[l] = ml[l - 1]; ulk[l] = ulk[l - 1];
            }
#pragma unroll
        for (int l = 0; l < NG_NCAND; ++l)
            if (l == pos) { val[l] = x; jnc[l] = j; inf[l] = info; upr[l] = c.upr; lwr[l] = c.lwr; ml[l] = c.ml; ulk[l] = link; }
        ++n;
        return true;
    }
};

// st[k].val for a run-time k without indexing the register array
__device__ __forceinline__ int ux_val(const UxCell (&st)[5], int k)
{
    int v = st[0].val;
#pragma unroll
    for (int q = 1; q < 5; ++q) if (q == k) v = st[q].val;
    return v;
}

#ifndef GSPALN_XUDH_MINB
#define GSPALN_XUDH_MINB 2               // CTAs of the wide class per SM the register budget is cut for
#endif
constexpr int XUDH_WIDE = 8;            // warps per problem of the wide class
constexpr int XUDH_WIDE_ROWS = 128;     // queries with at least this many rows run in the wide class

template <int NW>
__global__ void __launch_bounds__(NW == 1 ? NG_THREADS : 32 * NW, NW == 1 ? 1 : GSPALN_XUDH_MINB)
dp_xudh_kernel(const DevParams* __restrict__ gP, const short* __restrict__ tabs, int n_pen,
               const DevTask* __restrict__ tasks, const int* __restrict__ order, int ntasks, int* ticket,
               const unsigned char* __restrict__ apool, const ColInfo* __restrict__ cpool,
               unsigned char* workpool, long long work_slab, long long width_max,
               int* cpospool, DevUdhOut* results, const int* ready)
{
    constexpr int NT = 32 * NW;                         // lanes (= rows of a pass) per problem
    __shared__ DevParams sP;
    __shared__ int s_tk, s_rlst, s_best[NW][7];
    {
        const int* src = reinterpret_cast<const int*>(gP);
        int* dst = reinterpret_cast<int*>(&sP);
        for (int i = threadIdx.x; i < (int) (sizeof(DevParams) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const DevParams& P = sP;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x % NT;                  // lane of the problem's wavefront
    const int slot = threadIdx.x / NT;                  // problem slot of the CTA (NW == 1: its warps)
    auto sync_problem = [] { if (NW == 1) __syncwarp(); else __syncthreads(); };
    const bool dagp = P.noll == 3;
    const int noll = P.noll, nod = 2 * P.noll - 1;
    unsigned char* wbase = workpool + ((long long) blockIdx.x * (NW == 1 ? NG_WARPS : 1) + slot) * work_slab;

    for (;;) {
        int tk = 0;
        if (NW == 1) {
            if (lane == 0) tk = atomicAdd(ticket, 1);
            tk = __shfl_sync(FULL, tk, 0);
        } else {
            __syncthreads();
            if (threadIdx.x == 0) s_tk = atomicAdd(ticket, 1);
            __syncthreads();
            tk = s_tk;
        }
        if (tk >= ntasks) break;
        const int ti = order[tk];
        const DevTask t = tasks[ti];
        if (t.kind != 5 || ((t.flags & 32) != 0) != (NW > 1)) continue;     // another kernel / class runs it
        const int n_req = t.pad0;
        int* cpos = cpospool + t.pad1;
        for (int i = lane; i < 10 * (n_req + 1); i += NT) cpos[i] = (i % 10 == 0 || i % 10 == 2) ? END_OF_ULK : 0;
        bool arrived = wait_inputs(ready, tk);
        if (NW > 1) arrived = __syncthreads_and(arrived) != 0;
        if (!arrived) {
            if (lane == 0) { DevUdhOut r; memset(&r, 0, sizeof(r)); r.status = 4; r.score = NEVSEL32; results[ti] = r; }
            continue;
        }
        const unsigned char* aseq = apool + t.a_off;
        const ColInfo* cols = cpool + t.col_off;
        const int a_left = t.a_left, a_right = t.a_right, b_left = t.b_left, b_right = t.b_right;
        const int lw = t.lw, up = t.up, width = up - lw + 3;
        const bool a_exgl = t.flags & 1, a_exgr = t.flags & 2, b_exgl = t.flags & 4, b_exgr = t.flags & 8;
        const bool LocalL = P.local && a_exgl && b_exgl, LocalR = P.local && a_exgr && b_exgr;
        const int* cip = (t.flags & 16) ? reinterpret_cast<const int*>(aseq + ((a_right - a_left + 1 + 3) & ~3)) : nullptr;
        // spacing and number of the intermediate rows as lspS_ng sets them (src/fwd2s1.cc:1839-1851)
        const int mrows = a_right - a_left;
        const int intvl = (mrows + n_req) / (n_req + 1);
        const int n_im = (intvl * n_req == mrows) ? n_req - 1 : n_req;
        auto MI = [&](int i) { return a_left + intvl * (i + 1); };

        UxSlot* Hs = reinterpret_cast<UxSlot*>(wbase) - (lw - 1);       // by diagonal r in [lw - 1, up + 1]
        UxSlot* Fs = Hs + width_max;
        UxSlot* F2s = Fs + width_max;
        UxImd I;
        I.base = reinterpret_cast<int*>(reinterpret_cast<UxSlot*>(wbase) + 3 * width_max);
        I.width = width; I.noll = noll; I.lw = lw;

        const int r_black = b_left - a_right;
        const UxCell black{NG_NEVSEL, r_black, r_black, 0, END_OF_ULK};
        for (int i = lane; i < width; i += NT) {
            ux_st(Hs + (lw - 1) + i, black); ux_st(Fs + (lw - 1) + i, black);
            if (dagp) ux_st(F2s + (lw - 1) + i, black);
        }
        {
            const long long u = (long long) noll * width;
            for (long long i = lane; i < 4 * u * n_im; i += NT) {
                const int which = (int) ((i / u) & 3);
                I.base[i] = which < 2 ? END_OF_ULK : (which == 2 ? INT_MAX : INT_MIN);
            }
        }
        sync_problem();
        // ---- first row and first column (hinitS_ng)
        {
            const int r0 = b_left - a_left;
            if (lane == 0) ux_st(Hs + r0, UxCell{0, r0, r0, a_left, r0});
            if (a_exgl) {
                const int rr = min(up, b_right - a_left);
                for (int r = r0 + 1 + lane; r <= rr; r += NT) ux_st(Hs + r, UxCell{0, r, r, a_left, r});
            }
            const int rr = max(b_left - a_right, lw);
            if (b_exgl) {
                for (int r = rr + lane; r < r0; r += NT) ux_st(Hs + r, UxCell{0, r, r, a_left + (r0 - r), r});
            } else if (lane == 0) {
                int v = 0;
                for (int i = 1, r = r0 - 1; r >= rr; --r, ++i) {
                    v += i == 1 ? P.gappen1 : (i > P.codonk1 ? P.lgep : P.gep);
                    ux_st(Hs + r, UxCell{v, r0, r, a_left + i, r0});
                }
            }
        }
        __threadfence_block();
        sync_problem();

        int rlst = INT_MAX;                                 // uniform over the problem's lanes between passes
        int bval = NG_NEVSEL, bupr = 0, blwr = 0, bml = a_left, bulk = 0, bmr = a_right, bnr = b_right;   // LocalR
        const int m_first = a_exgl ? a_left + 1 : a_left;
        for (int m0 = m_first; m0 <= a_right; ) {
            // rows m0 .. m9 of this pass, cut so that it holds at most one intermediate row
            int m9 = min(m0 + NT - 1, a_right);
            int ia = (m0 - a_left + intvl - 1) / intvl - 1;     // first intermediate at or below m0
            if (ia < 0) ia = 0;
            const int mi_a = ia < n_im ? MI(ia) : INT_MAX;
            if (ia + 1 < n_im && MI(ia + 1) <= m9) m9 = MI(ia + 1) - 1;
            const int m = m0 + lane;
            const bool row = m <= m9;
            const bool is_imd = row && m == mi_a;
            const int lo = ng_row_lo(t, m), hi = ng_row_hi(t, m);
            const int arow = (m == a_left || !row) ? ZROW : (int) aseq[m - 1 - a_left];
            const int sigB = (cip && row) ? cip[m - a_left] : 0;
            const int last_lane = m9 - m0;
            const int s_begin = ng_row_lo(t, m0) + 1;
            const int s_end = ng_row_hi(t, m9) + last_lane;

            UxCell hleft = black, e1 = black, e2 = black;
            int psp = 0;
            UxList L;
            L.clear();
            int my_rlst = rlst;

            for (int s = s_begin; s <= s_end; ++s) {
                const int n = s - lane;
                if (row && n > lo && n <= hi) {
                    const int r = n - m;
                    const ColInfo col = cols[n - b_left];
                    if (n == lo + 1) hleft = ux_ld(Hs + r - 1);
                    // the five gap states: 0 H, 1 E, 2 F, 3 E2, 4 F2
                    UxCell st[5];
                    st[0] = ux_ld(Hs + r);
                    int mx = 0;
                    if (m != a_left) {
                        st[0].val += P.mtxT[(int) col.code * MTX_LD + arow];
                        const UxCell uh = ux_ld(Hs + r + 1), uf = ux_ld(Fs + r + 1);
                        int x = uh.val + P.gop;
                        if (x >= uf.val) { st[2] = uh; st[2].val = x; } else st[2] = uf;
                        st[2].val += P.gep;
                        if (st[2].val >= st[0].val) mx = 2;
                        if (dagp) {
                            const UxCell uf2 = ux_ld(F2s + r + 1);
                            x = uh.val + P.lgop;
                            if (x >= uf2.val) { st[4] = uh; st[4].val = x; } else st[4] = uf2;
                            st[4].val += P.lgep;
                            if (st[4].val >= ux_val(st, mx)) mx = 4;
                        } else st[4] = black;
                    } else {
                        st[2] = ux_ld(Fs + r);              // untouched band entries
                        st[4] = dagp ? ux_ld(F2s + r) : black;
                    }
                    {
                        int x = hleft.val + P.gop;
                        const int prev_psp = psp;
                        if (x >= e1.val) { e1 = hleft; e1.val = x; psp = psp ? 1 : 0; }
                        else psp &= 3;
                        e1.val += P.gep;
                        st[1] = e1;
                        if (st[1].val >= ux_val(st, mx)) mx = 1;
                        if (dagp) {
                            x = hleft.val + P.lgop;
                            if (x >= e2.val) { e2 = hleft; e2.val = x; if (prev_psp) psp |= 2; }
                            else psp |= prev_psp & 2;
                            e2.val += P.lgep;
                            st[3] = e2;
                            if (st[3].val >= ux_val(st, mx)) mx = 3;
                        } else st[3] = black;
                    }
                    const int cano5 = col.pad[1] & 15, cano3 = col.pad[1] >> 4;
                    const int psp_bit[5] = {4, 1, 8, 2, 16};    // src/aln.h:56
                    // acceptor: strictly better donors only, per gap state
                    bool spj3 = false;
                    if (cano3 && L.n > 0) {
                        int tu[5], tl[5], tm[5], tk_[5];
                        unsigned hit = 0;
#pragma unroll
                        for (int l = 0; l <= NG_NCAND; ++l) {
                            if (l >= L.n || n - L.jnc[l] < P.llmt) continue;
                            const int k = L.inf[l] & 15;
                            const int x = L.val[l] + sigB + ng_spjscr(tabs, n_pen, L.inf[l] >> 4, n - L.jnc[l], col);
#pragma unroll
                            for (int q = 0; q < 5; ++q)
                                if (q == k && x > st[q].val) {
                                    st[q].val = x; tu[q] = L.upr[l]; tl[q] = L.lwr[l]; tm[q] = L.ml[l]; tk_[q] = L.ulk[l];
                                    hit |= 1u << q;
                                }
                        }
                        int maxk = nod;
#pragma unroll
                        for (int q = 0; q < 5; ++q) {
                            if (q >= nod || !(hit >> q & 1u)) continue;
                            psp |= psp_bit[q];
                            if (q == 0) spj3 = true;
                            st[q].upr = max(tu[q], r); st[q].lwr = min(tl[q], r); st[q].ml = tm[q]; st[q].ulk = tk_[q];
                            if (st[q].val > ux_val(st, mx)) { maxk = q; mx = q; }
                        }
                        if (is_imd && maxk < nod) {
                            int link = 0;
#pragma unroll
                            for (int q = 0; q < 5; ++q) if (q == maxk) link = tk_[q];
                            I.at(ia, 0, 0, r) = link;
                            my_rlst = r;
#pragma unroll
                            for (int q = 0; q < 5; ++q) if (q == mx) st[q].ulk = r;
                            if (maxk == 0) {
#pragma unroll
                                for (int c = 1; c < 3; ++c) {
                                    if (c >= noll) continue;
                                    const int d = 2 * c - 1;
                                    const int g = c == 1 ? P.gop : P.lgop;
                                    if ((hit >> d & 1u) && st[d].val > st[0].val + g) {
                                        st[d].ulk = r + c * width;
                                        I.at(ia, 0, c, r) = tk_[d];
                                    }
                                    if ((hit >> (d + 1) & 1u) && st[d + 1].val > st[0].val + g) st[d + 1].ulk = r + c * width;
                                }
                            }
                        }
                    }
                    // best state
                    const int hd = mx;
                    if (mx == 0) {
                        if (LocalR && st[0].val > bval) {
                            bval = st[0].val; bupr = st[0].upr; blwr = st[0].lwr; bml = st[0].ml; bulk = st[0].ulk;
                            bmr = m; bnr = n;
                        }
                    } else {
#pragma unroll
                        for (int q = 1; q < 5; ++q) if (q == mx) st[0] = st[q];
                        if (st[0].upr < r) st[0].upr = r;
                        if (st[0].lwr > r) st[0].lwr = r;
                    }
                    if (LocalL && st[0].val <= 0) { st[0].val = 0; st[0].ml = m; st[0].ulk = st[0].upr = st[0].lwr = r; }
                    const int mxv = ux_val(st, hd);         // (the best state's value as it stands now)
                    // donor: the best (value + 5' signal) of this row by gap state
                    if (cano5) {
                        const int sigJ = col.sig5;
                        const int d5 = col.pad[0] & 15;
#pragma unroll
                        for (int q = 0; q < 5; ++q) {
                            if (q >= nod || q < (hd == 0 ? 0 : 1) || (psp & psp_bit[q])) continue;
                            if (q != hd) {
                                int z = mxv;
                                if (hd == 0 || (q - hd) % 2) z += q / 2 == 0 ? 0 : (q / 2 == 1 ? P.gop : P.lgop);
                                if (st[q].val <= z) continue;
                            }
                            if (L.insert(st[q].val + sigJ, n, q | (d5 << 4), st[q], is_imd ? r : st[q].ulk) && is_imd && q == 1)
                                I.at(ia, 0, 0, r) = my_rlst;
                        }
                    }
                    // intermediate row: record the links and bounds, restart them
                    if (is_imd) {
                        if (hd == 0) my_rlst = r;
                        else if (!spj3 && (hd & 1)) I.at(ia, 0, 0, r) = my_rlst;
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            if (k >= noll) continue;
                            UxCell& c = st[2 * k];
                            I.at(ia, 1, k, r) = c.ulk;
                            I.at(ia, 2, k, r) = min(r, c.lwr);
                            I.at(ia, 3, k, r) = max(r, c.upr);
                            c.lwr = c.upr = r;
                            c.ulk = r + k * width;
                        }
                    }
                    e1 = st[1];
                    if (dagp) e2 = st[3];
                    ux_st(Hs + r, st[0]); ux_st(Fs + r, st[2]);
                    if (dagp) ux_st(F2s + r, st[4]);
                    hleft = st[0];
                }
                sync_problem();
            }
            // hand the intermediate row's last diagonal to the next pass
            if (NW == 1) {
                const unsigned who = __ballot_sync(FULL, is_imd);
                if (who) rlst = __shfl_sync(FULL, my_rlst, __ffs(who) - 1);
            } else {
                if (threadIdx.x == 0) s_rlst = rlst;
                __syncthreads();
                if (is_imd) s_rlst = my_rlst;
                __syncthreads();
                rlst = s_rlst;
            }
            __threadfence_block();
            sync_problem();
            m0 = m9 + 1;
        }

        // ---- end point (hlastS_ng) and the back-walk over the intermediates
        if (LocalR) {
            // row-major order: the first cell that reached the maximum
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const int ov = __shfl_xor_sync(FULL, bval, o), ou = __shfl_xor_sync(FULL, bupr, o);
                const int ol = __shfl_xor_sync(FULL, blwr, o), om = __shfl_xor_sync(FULL, bml, o);
                const int ok = __shfl_xor_sync(FULL, bulk, o), omr = __shfl_xor_sync(FULL, bmr, o);
                const int onr = __shfl_xor_sync(FULL, bnr, o);
                if (ov > bval || (ov == bval && ov > NG_NEVSEL && (omr < bmr || (omr == bmr && onr < bnr)))) {
                    bval = ov; bupr = ou; blwr = ol; bml = om; bulk = ok; bmr = omr; bnr = onr;
                }
            }
            if (NW > 1) {
                // across the warps of the problem, same order
                if ((threadIdx.x & 31) == 0) {
                    int* b = s_best[threadIdx.x >> 5];
                    b[0] = bval; b[1] = bupr; b[2] = blwr; b[3] = bml; b[4] = bulk; b[5] = bmr; b[6] = bnr;
                }
                __syncthreads();
                if (threadIdx.x == 0)
                    for (int w = 1; w < NW; ++w) {
                        const int* b = s_best[w];
                        if (b[0] > bval || (b[0] == bval && b[0] > NG_NEVSEL && (b[5] < bmr || (b[5] == bmr && b[6] < bnr)))) {
                            bval = b[0]; bupr = b[1]; blwr = b[2]; bml = b[3]; bulk = b[4]; bmr = b[5]; bnr = b[6];
                        }
                    }
            }
        }
        if (lane == 0) {
            int A_left = a_left, A_right = a_right, B_left = b_left, B_right = b_right;
            int mval, mupr, mlwr, mml, mulk;
            const int rr = b_right - a_right;
            if (LocalR) {
                int i = n_im;
                while (--i >= 0 && MI(i) > A_right) ;
                A_right = bmr; B_right = bnr;
                if (i < 0) i = 0;
                cpos[10 * i + 8] = blwr; cpos[10 * i + 9] = bupr;
                mval = bval; mupr = bupr; mlwr = blwr; mml = bml; mulk = bulk;
            } else {
                int mxr = rr;
                int best = ux_ld(Hs + rr).val;
                if (b_exgr)
                    for (int r = min(up, b_right - a_left); r > rr; --r) {
                        const int v = ux_ld(Hs + r).val;
                        if (v > best) { best = v; mxr = r; }
                    }
                if (a_exgr)
                    for (int r = max(lw, b_left - a_right); r < rr; ++r) {
                        const int v = ux_ld(Hs + r).val;
                        if (v > best) { best = v; mxr = r; }
                    }
                const UxCell c = ux_ld(Hs + mxr);
                mval = c.val; mupr = c.upr; mlwr = c.lwr; mml = c.ml; mulk = c.ulk;
                if (b_exgr && rr < mxr) A_right = b_right - mxr;
                if (a_exgr && rr > mxr) B_right = a_right + mxr;
            }
            int i = n_im;
            while (--i >= 0 && MI(i) > A_right) ;
            if (i < 0 && MI(0) > A_right) cpos[2] = B_right;
            int r = B_right - A_right;
            cpos[10 * (i + 1) + 8] = min(mlwr, r);
            cpos[10 * (i + 1) + 9] = max(mupr, r);
            r = mulk;
            for ( ; i >= 0 && MI(i) > mml; --i) {
                int c = 0, d = 0;
                if (r > up) { d = (int) (((long long) r - up + width - 1) / width); r -= d * width; }   // for ( ; r > up; r -= width) ++d
                if (d >= noll || r < lw - 1) { cpos[10 * i] = END_OF_ULK; break; }      // (a link no pass wrote)
                if (I.at(i, 1, d, r) < END_OF_ULK) {
                    cpos[10 * i + c++] = MI(i);
                    cpos[10 * i + c++] = d > 0 ? 1 : 0;
                    for (int rp = I.at(i, 0, d, r); lw <= rp && rp < up && r != rp; rp = I.at(i, 0, 0, r = rp))
                        if (c < 7) cpos[10 * i + c++] = r + MI(i);
                    if (c < 8) cpos[10 * i + c++] = r + MI(i);
                    cpos[10 * i + c] = END_OF_ULK;
                    cpos[10 * i + 8] = I.at(i, 2, d, r);
                    cpos[10 * i + 9] = I.at(i, 3, d, r);
                    r = I.at(i, 1, d, r);
                    if (r == END_OF_ULK) break;
                } else
                    cpos[10 * i] = END_OF_ULK;
            }
            if (r > up) r -= (int) (((long long) r - up + width - 1) / width) * width;      // for ( ; r > up; r -= width) ;
            if (LocalL) {
                A_left = mml;
                B_left = r + mml;
            } else {
                const int rl = B_left - A_left;
                if (b_exgl && rl > r) {
                    A_left = B_left - r;
                    for (int j = 0; j < n_im && MI(j) < A_left; ++j) cpos[10 * j] = END_OF_ULK;
                }
                if (a_exgl && rl < r) B_left = A_left + r;
            }
            ++i;
            if ((i < n_im && MI(i) < A_left) || cpos[10 * i + 2] < B_left) mval = NEVSEL32;
            else {
                const int rl = B_left - A_left;
                cpos[10 * i + 8] = min(rl, cpos[10 * i + 8]);
                cpos[10 * i + 9] = max(rl, cpos[10 * i + 9]);
            }
            DevUdhOut o;
            o.score = mval; o.status = 0;
            o.a_left = A_left; o.a_right = A_right; o.b_left = B_left; o.b_right = B_right;
            o.pad0 = o.pad1 = 0;
            results[ti] = o;
        }
        sync_problem();
    }
}

}   // namespace gspaln
