// gspaln_hxudh.cuh -- the scalar unidirectional-Hirschberg pass for protein queries as a warp kernel.
//
// Semantics: bit-identical to Aln2h1::hirschbergH_ng (src/fwd2h1.cc:1085-1520, with hinitH_ng /
// hlastH_ng 941-1083 and the bounded intermediates of src/udh_intermediate.h:29-88), the Hirschberg
// pass of the reference's default mode `-A0`.  The recurrences are forwardH_ng's (codon-wise and
// frame-shift gaps, three splice phases with the split codon translated, coding potential); instead
// of path records every cell state carries {direction, highest / lowest diagonal since the last
// intermediate row, start row, link}.  At the intermediate rows links and bounds are recorded and
// restarted; the back-walk turns them into the crossing records `cpos[]`, entries [8], [9] band
// the blocks of the post-work (src/fwd2h1.cc:2066-2072).
//
// Mapping: as dp_hxild_kernel (gspaln_hng.cuh) -- one warp per problem, 32 consecutive query rows
// in lock step, each lane one genome column behind the lane of the row above, the band rows H, F,
// F2 shared in global memory (32-byte entries here), one __syncwarp() per step.  A pass is cut so
// that it holds at most one intermediate row: its three "last diagonal on the row" registers (by
// column phase) run along that lane and are handed from pass to pass.
#pragma once
#include "gspaln_hng.cuh"
#include "gspaln_udh.cuh"

namespace gspaln {

struct HuCell { int val, dir, upr, lwr, ml, ulk; };         // Rvdwml, src/aln.h:138-145
struct __align__(16) HuSlot { int v[8]; };

__device__ __forceinline__ HuCell hu_ld(const HuSlot* p)
{
    const int4 a = __ldcg(reinterpret_cast<const int4*>(p));
    const int2 b = __ldcg(reinterpret_cast<const int2*>(p) + 2);
    return HuCell{a.x, a.y, a.z, a.w, b.x, b.y};
}
__device__ __forceinline__ void hu_st(HuSlot* p, const HuCell& c)
{
    *reinterpret_cast<int4*>(p) = make_int4(c.val, c.dir, c.upr, c.lwr);
    reinterpret_cast<int2*>(p)[2] = make_int2(c.ml, c.ulk);
}

// donor list of one row and splice phase (see HxList): ties pass, payload = bounds, start row, link
struct HuList {
    int val[HNG_NCAND + 1], jnc[HNG_NCAND + 1], st[HNG_NCAND + 1];
    int upr[HNG_NCAND + 1], lwr[HNG_NCAND + 1], ml[HNG_NCAND + 1], ulk[HNG_NCAND + 1];
    int n;
    __device__ void clear()
    {
        n = 0;
        for (int l = 0; l <= HNG_NCAND; ++l) { val[l] = HNG_NEVSEL; jnc[l] = st[l] = ml[l] = 0; upr[l] = INT_MIN; lwr[l] = INT_MAX; ulk[l] = END_OF_ULK; }
    }
    __device__ bool insert(int x, int state, int j, const HuCell& c, int link)
    {
        if (n > HNG_NCAND) n = HNG_NCAND;
        int pos = 0;
        while (pos < n && pos < HNG_NCAND && val[pos] > x) ++pos;
        if (pos >= HNG_NCAND) return false;
        for (int l = HNG_NCAND; l > pos; --l) {
            val[l] = val[l - 1]; jnc[l] = jnc[l - 1]; st[l] = st[l - 1];
            upr[l] = upr[l - 1]; lwr[l] = lwr[l - 1]; ml[l] = ml[l - 1]; ulk[l] = ulk[l - 1];
        }
        val[pos] = x; jnc[pos] = j; st[pos] = state; upr[pos] = c.upr; lwr[pos] = c.lwr; ml[pos] = c.ml; ulk[pos] = link;
        ++n;
        return true;
    }
};

struct DevUdhHTask {            // DevNgHTask + what the pass needs
    DevNgHTask g;
    int n_req, pad;             // intermediate rows asked for (before the even-division correction)
    long long cpos_off;         // ints, (n_req + 1) x 10
};

template <int NW>                       // warps per problem, as dp_hxild_kernel
__global__ void __launch_bounds__(NW == 1 ? HNG_THREADS : 32 * NW)
dp_hxudh_kernel(const DevNgHParams* __restrict__ gP, const DevUdhHTask* __restrict__ tasks, int ntasks, int* ticket,
                const unsigned char* __restrict__ apool, const unsigned char* __restrict__ bpool,
                const short* __restrict__ sgpool, const unsigned short* __restrict__ i53pool,
                const int* __restrict__ cippool, unsigned char* workpool, int* cpospool, DevUdhOut* results)
{
    constexpr int NT = 32 * NW;
    __shared__ DevNgHParams P;
    __shared__ int s_ti, s_rl[3], s_best[NW][7];
    if (threadIdx.x < sizeof(DevNgHParams) / 4)
        reinterpret_cast<int*>(&P)[threadIdx.x] = reinterpret_cast<const int*>(gP)[threadIdx.x];
    __syncthreads();
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x % NT;                  // lane of the problem's wavefront
    auto sync_problem = [] { if (NW == 1) __syncwarp(); else __syncthreads(); };

    for (;;) {
        int ti = 0;
        if (NW == 1) {
            if (lane == 0) ti = atomicAdd(ticket, 1);
            ti = __shfl_sync(FULL, ti, 0);
        } else {
            __syncthreads();
            if (threadIdx.x == 0) s_ti = atomicAdd(ticket, 1);
            __syncthreads();
            ti = s_ti;
        }
        if (ti >= ntasks) break;
        const DevNgHTask t = tasks[ti].g;
        if ((t.wide != 0) != (NW > 1)) continue;            // the other class runs it
        const int n_req = tasks[ti].n_req;
        int* cpos = cpospool + tasks[ti].cpos_off;
        HngIn T;
        T.a = apool + t.a_off; T.b = bpool + t.b_off; T.sg = sgpool + 8 * t.sg_off; T.i53 = i53pool + t.sg_off;
        T.a_lo = t.a_lo; T.b_lo = t.b_lo; T.b_left = t.b_left; T.b_right = t.b_right;
        T.cip = t.cip_off >= 0 ? cippool + t.cip_off : nullptr; T.cip_lo = 3 * t.a_left - 1;
        const int a_exgl = t.a_exgl, a_exgr = t.a_exgr, b_exgl = t.b_exgl, b_exgr = t.b_exgr;
        const int width = t.up - t.lw + 7;
        const int noll = P.noll, nod = 2 * P.noll - 1;
        const bool dagp = P.noll == 3;
        const bool Local = P.local, LocalL = Local && a_exgl && b_exgl, LocalR = Local && a_exgr && b_exgr;
        const int a_left = t.a_left, a_right = t.a_right, b_left = t.b_left, b_right = t.b_right;
        const int lw = t.lw, up = t.up;
        const int mrows = a_right - a_left;
        const int intvl = (mrows + n_req) / (n_req + 1);
        const int n_im = (intvl * n_req == mrows) ? n_req - 1 : n_req;
        auto MI = [&](int i) { return a_left + intvl * (i + 1); };

        HuSlot* buf = reinterpret_cast<HuSlot*>(workpool + t.work_off);
        HuSlot* Hb = buf - lw + 3;              // by diagonal r = n - 3 m in [lw - 3, up + 3]
        HuSlot* Fb = Hb + width;
        HuSlot* F2b = Fb + width;
        UxImd I;                                // hlnk | vlnk | lwrb | uprb per intermediate (biased by lw - 1)
        I.base = reinterpret_cast<int*>(buf + 3 * (width + 8));
        I.width = width; I.noll = noll; I.lw = lw;

        const int r_black = b_left - 3 * a_right;
        const HuCell black{HNG_NEVSEL, 0, r_black, r_black, 0, END_OF_ULK};
        for (int i = lane; i < 3 * (width + 8); i += NT) hu_st(buf + i, black);
        {
            const long long u = (long long) noll * width;
            for (long long i = lane; i < 4 * u * n_im; i += NT) {
                const int which = (int) ((i / u) & 3);
                I.base[i] = which < 2 ? END_OF_ULK : (which == 2 ? INT_MAX : INT_MIN);
            }
        }
        for (int i = lane; i < 10 * (n_req + 1); i += NT) cpos[i] = END_OF_ULK;
        sync_problem();

        // ---- start row and start column (hinitH_ng): serial, lane 0
        if (lane == 0) {
            auto sigS = [&](int n) { const int s = T.sgf(n, F_SIGS); return s > 0 ? s : 0; };
            const int r0 = b_left - 3 * a_left;
            const int dir0 = a_exgl ? DEAD : DIAG;
            hu_st(Hb + r0, HuCell{a_exgl ? sigS(b_left + 1) : 0, dir0, r0, r0, a_left, r0});
            if (a_exgl) {
                int jnc = b_left;
                const int rr = min(up, b_right - 3 * a_left);
                for (int i = 1, r = r0 + 1; r <= rr; ++r, ++i) {
                    const int n = b_left + i;
                    HuCell h;
                    if (i < 3) h = HuCell{sigS(n + 1), dir0, r, r, a_left, r};     // (upr is set below)
                    else {
                        h = hu_ld(Hb + r - 3);
                        const int d = n - jnc;
                        if (!(a_exgl & 1) && d == 3) h.val += P.gop;
                        if (!(a_exgl & 2)) h.val += gap_ext3(P, d);
                        h.val += T.sgf(n + 1 - 3, F_SIGE);
                        h.dir = HORI;
                        const HuCell h1 = hu_ld(Hb + r - 1), h2 = hu_ld(Hb + r - 2);
                        if (h1.val + P.gw1 > h.val) { h = h1; h.val += P.gw1; h.dir = HOR1; }
                        if (h2.val + P.gw2 > h.val) { h = h2; h.val += P.gw2; h.dir = HOR2; }
                    }
                    const int xs = sigS(n + 1);
                    if (h.val < xs) { h.val = xs; h.dir = DEAD; jnc = n; h.lwr = h.ulk = r; }
                    h.upr = r;
                    hu_st(Hb + r, h);
                }
            }
            const int rr = max(lw, b_left - 3 * a_right);
            for (int i = 1, r = r0 - 1; r >= rr; --r, ++i) {
                HuCell h;
                if (b_exgl == 1) h = HuCell{0, DEAD, r, r, a_left + i / 3, r};
                else if (i <= 3) {
                    h = hu_ld(Hb + r + i);
                    if (!(b_exgl & 2)) h.val += P.gep;
                    if (!(b_exgl & 1)) h.val += P.gop;
                    if (i < 3) h.val += P.extragop;
                    h.dir = VERT;
                    h.ml += i / 3;
                    h.lwr = h.ulk = r;
                } else {
                    h = hu_ld(Hb + r + 3);
                    if (!(b_exgl & 2)) h.val += gap_ext3(P, i);
                    h.lwr = h.ulk = r;
                    ++h.ml;
                }
                hu_st(Hb + r, h);
            }
        }
        __threadfence_block();
        sync_problem();

        int rl0 = INT_MAX, rl1 = INT_MAX, rl2 = INT_MAX;       // rlst[3], uniform over the problem's lanes between passes
        int bval = HNG_NEVSEL, bupr = 0, blwr = 0, bml = a_left, bulk = 0, bmr = a_right, bnr = b_right;   // LocalR
        const int m_first = a_exgl ? a_left + 1 : a_left;
        for (int m0 = m_first; m0 <= a_right; ) {
            int m9 = min(m0 + NT - 1, a_right);
            int ia = (m0 - a_left + intvl - 1) / intvl - 1;
            if (ia < 0) ia = 0;
            const int mi_a = ia < n_im ? MI(ia) : INT_MAX;
            if (ia + 1 < n_im && MI(ia + 1) <= m9) m9 = MI(ia + 1) - 1;
            const int m = m0 + lane;
            const bool row = m <= m9;
            const bool is_imd = row && m == mi_a;
            const int n0 = max(3 * m + lw - 1, b_left), n9 = min(3 * m + up, b_right);
            const int last_lane = m9 - m0;
            const int s_begin = max(3 * m0 + lw - 1, b_left);
            const int s_end = min(3 * m9 + up, b_right) + last_lane;
            HuCell e1[3] = {black, black, black}, e2[3] = {black, black, black};
            int q = 0;
            HuList don[3];
            don[0].clear(); don[1].clear(); don[2].clear();
            const int* prof_prev = P.mtx + (row ? T.aa(m > 0 ? m - 1 : 0) : 0) * P.simdim;
            const int* prof_next = P.mtx + (row ? T.aa(m) : 0) * P.simdim;
            bool started = false;
            int sigB[3] = {0, 0, 0};
            if (T.cip && row)
                for (int phs = -1; phs < 2; ++phs) sigB[phs + 1] = T.cip[3 * m - phs - T.cip_lo];
            int rl[3] = {rl0, rl1, rl2};

            for (int s = s_begin; s <= s_end; ++s) {
                const int n = s - lane;
                if (row && n >= n0 && n <= n9) {
                    const int r = n - 3 * m;
                    if (!started) {
                        started = true;
                        if (!b_exgl && m == a_left) {
                            const HuCell c = hu_ld(Hb + r);
                            e1[2] = c; e1[2].val += P.gw3;
                            e2[2] = c; e2[2].val += P.gw3l;
                        }
                    }
                    const int sigE = n > b_left ? T.sgf(n - 2, F_SIGE) : 0;
                    const HuCell hq = hu_ld(Hb + r);
                    HuCell st[5];                           // 0 H, 1 E, 2 F, 3 E2, 4 F2
                    st[0] = hq; st[1] = e1[q]; st[2] = hu_ld(Fb + r); st[3] = e2[q];
                    st[4] = dagp ? hu_ld(F2b + r) : black;
                    int mx = 0;
                    if (m != a_left) {
                        if (n < b_left + 3) st[0] = black;
                        else {
                            st[0].val += prof_prev[T.tron(n - 2)] + sigE;
                            st[0].dir = (hq.dir & DIAG) ? DIAG : NEWD;      // a bit test in this function
                        }
                        const HuCell u1 = hu_ld(Hb + r + 1), u2 = hu_ld(Hb + r + 2), u3 = hu_ld(Hb + r + 3);
                        const HuCell fu = hu_ld(Fb + r + 3);
                        const int ext = fu.val + P.gep;
                        int x = u1.val + (h_is_vert(u1.dir) ? P.gape1 : P.gw1);
                        if (x > ext) { st[2] = u1; st[2].val = x; st[2].dir = SLA2; } else st[2].val = ext;
                        x = u2.val + (h_is_vert(u2.dir) ? P.gape2 : P.gw2);
                        if (x > st[2].val) { st[2] = u2; st[2].val = x; st[2].dir = SLA1; }
                        x = u3.val + P.gw3;
                        if (x >= st[2].val) { st[2] = u3; st[2].val = x; st[2].dir = VERT; }
                        else if (ext >= st[2].val) { st[2] = fu; st[2].val = ext; st[2].dir = VERT; }
                        if (st[2].val >= st[mx].val) mx = 2;
                        if (dagp) {
                            const HuCell f2u = hu_ld(F2b + r + 3);
                            x = u3.val + P.gw3l;
                            const int ext2 = f2u.val + P.lgep;
                            if (x >= ext2) { st[4] = u3; st[4].val = x; st[4].dir = VERL; }
                            else { st[4] = f2u; st[4].val = ext2; }
                            if (st[4].val >= st[mx].val) mx = 4;
                        }
                    }
                    if (n > n0 + 2) {
                        const HuCell l3 = hu_ld(Hb + r - 3);
                        int x = l3.val + P.gw3;
                        st[1].val += P.gep;
                        if (x > st[1].val) { st[1] = l3; st[1].val = x; }
                        st[1].val += sigE;
                        st[1].dir = (st[1].dir & SPIN) + HORI;
                        if (dagp) {
                            x = l3.val + P.gw3l;
                            st[3].val += P.lgep;
                            if (x > st[3].val) { st[3] = l3; st[3].val = x; }
                            st[3].val += sigE;
                            st[3].dir = (st[3].dir & SPIN) + HORL;
                            if (st[3].val > st[mx].val) mx = 3;
                        }
                    }
                    if (n > n0 + 1) {
                        const HuCell l2 = hu_ld(Hb + r - 2);
                        const int x = l2.val + P.gw2;
                        if (x > st[1].val) { st[1] = l2; st[1].val = x; st[1].dir = HOR2; }
                    }
                    {
                        const HuCell l1 = hu_ld(Hb + r - 1);
                        const int x = l1.val + P.gw1;
                        if (x > st[1].val) { st[1] = l1; st[1].val = x; st[1].dir = HOR1; }
                    }
                    if (st[1].val > st[mx].val) mx = 1;
                    const int qn = q == 2 ? 0 : q + 1;      // the ring index after this column (rlst's index)

                    // acceptor
                    bool spj3 = false;
                    const int phs3 = T.sgf(n, F_PHS3);
                    if (P.spj && phs3 > -2) {
                        for (int phs = phs3 == 2 ? -1 : phs3; ; phs = 1) {
                            const int nb = n - phs;
                            const HuList& L = don[phs + 1];
                            int tu[5], tl[5], tm[5], tk[5], td[5];
                            unsigned hit = 0;
                            for (int l = 0; l < L.n; ++l) {
                                const int k = L.st[l];
                                if ((phs == 1 && k == 2) || nb - L.jnc[l] < P.minl) continue;
                                int x = L.val[l] + sigB[phs + 1] + hx_spjscr(P, T, L.jnc[l], nb);
                                if (k == 0 && phs) {
                                    const unsigned char* cs = hx_spjseq(P, T, L.jnc[l], nb);
                                    if (phs == 1) x += prof_prev[cs[0]];
                                    else x += prof_next[cs[1]] - prof_next[T.tron(n + 1)] - T.sgf(n + 1, F_SIGE);
                                }
                                if (x > st[k].val) {
                                    st[k].val = x; tu[k] = L.upr[l]; tl[k] = L.lwr[l]; tm[k] = L.ml[l]; tk[k] = L.ulk[l]; td[k] = k;
                                    hit |= 1u << k;
                                }
                            }
                            int maxk = nod;
                            for (int k = 0; k < nod; ++k) {
                                if (!(hit >> k & 1u)) continue;
                                st[k].dir = c_nod2dir[td[k]] | SPIN;
                                st[k].upr = max(tu[k], r); st[k].lwr = min(tl[k], r); st[k].ml = tm[k]; st[k].ulk = tk[k];
                                if (st[k].val >= st[mx].val) { maxk = k; mx = k; }
                            }
                            if (is_imd && maxk < nod) {
                                I.at(ia, 0, 0, r) = tk[maxk];
                                st[mx].ulk = rl[qn] = r;
                                spj3 = true;
                                if (maxk == 0) {
                                    for (int c = 1, d = 1; c < noll; ++c, d += 2) {
                                        const int g = c == 1 ? P.gop : P.lgop;
                                        if ((hit >> d & 1u) && st[d].val > st[0].val + g) {
                                            st[d].ulk = r + c * width;
                                            I.at(ia, 0, c, r) = tk[d];
                                        }
                                        if ((hit >> (d + 1) & 1u) && st[d + 1].val > st[0].val + g) st[d + 1].ulk = r + c * width;
                                    }
                                }
                            }
                            if (phs3 - phs != 3) break;         // AGAG: both phases
                        }
                    }

                    // best state (the source state is widened, then copied)
                    if (mx == 0) {
                        if (LocalR && st[0].val > bval) {
                            bval = st[0].val; bupr = st[0].upr; blwr = st[0].lwr; bml = st[0].ml; bulk = st[0].ulk;
                            bmr = m; bnr = n;
                        }
                    } else {
                        if (st[mx].upr < r) st[mx].upr = r;
                        if (st[mx].lwr > r) st[mx].lwr = r;
                        st[0] = st[mx];
                    }
                    if (LocalL && st[0].val <= 0) { st[0].val = 0; st[0].dir = 0; st[0].ml = m; st[0].ulk = st[0].upr = st[0].lwr = r; }
                    const int hd = c_dir2nod[st[mx].dir & 15];
                    const int mxval = st[mx].val;

                    // donor
                    const int phs5 = T.sgf(n, F_PHS5);
                    if (P.spj && phs5 > -2) {
                        for (int phs = phs5 == 2 ? -1 : phs5; ; phs = 1) {
                            const int nb = n - phs;
                            const int sigJ = T.sgf(nb, F_SIG5);
                            for (int k = (hd == 0 || phs == 1) ? 0 : 1; k < nod; ++k) {
                                const bool cross = phs == 1 && k == 0;
                                const HuCell from = cross ? hq : st[k];
                                if (!from.dir || (from.dir & SPIN)) continue;
                                if (!cross && k != hd && hd >= 0) {
                                    int z = mxval;
                                    if (hd == 0 || (k - hd) % 2) z += k / 2 == 0 ? 0 : (k / 2 == 1 ? P.gop : P.lgop);
                                    if (from.val <= z) continue;
                                }
                                if (don[phs + 1].insert(from.val + sigJ, k, nb, from, is_imd ? r : from.ulk) && is_imd && k == 1)
                                    I.at(ia, 0, 0, r) = rl[qn];
                            }
                            if (phs5 - phs != 3) break;         // GTGT: both phases
                        }
                    }
                    // intermediate row: record links and bounds, restart them
                    if (is_imd) {
                        if (hd == 0) rl[qn] = r;
                        else if (!spj3 && hd % 2) I.at(ia, 0, 0, r) = rl[qn];
                        for (int k = 0; k < noll; ++k) {
                            HuCell& c = st[2 * k];
                            I.at(ia, 1, k, r) = c.ulk;
                            I.at(ia, 2, k, r) = min(r, c.lwr);
                            I.at(ia, 3, k, r) = max(r, c.upr);
                            c.lwr = c.upr = r;
                            c.ulk = r + k * width;
                        }
                    }
                    hu_st(Hb + r, st[0]);
                    hu_st(Fb + r, st[2]);
                    if (dagp) hu_st(F2b + r, st[4]);
                    e1[q] = st[1]; e2[q] = st[3];
                    q = qn;
                }
                sync_problem();
            }
            if (NW == 1) {
                const unsigned who = __ballot_sync(FULL, is_imd);
                if (who) {
                    const int src = __ffs(who) - 1;
                    rl0 = __shfl_sync(FULL, rl[0], src); rl1 = __shfl_sync(FULL, rl[1], src); rl2 = __shfl_sync(FULL, rl[2], src);
                }
            } else {
                if (threadIdx.x == 0) { s_rl[0] = rl0; s_rl[1] = rl1; s_rl[2] = rl2; }
                __syncthreads();
                if (is_imd) { s_rl[0] = rl[0]; s_rl[1] = rl[1]; s_rl[2] = rl[2]; }
                __syncthreads();
                rl0 = s_rl[0]; rl1 = s_rl[1]; rl2 = s_rl[2];
            }
            __threadfence_block();
            sync_problem();
            m0 = m9 + 1;
        }

        // ---- end point (hlastH_ng) and the back-walk over the intermediates
        if (LocalR) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const int ov = __shfl_xor_sync(FULL, bval, o), ou = __shfl_xor_sync(FULL, bupr, o);
                const int ol = __shfl_xor_sync(FULL, blwr, o), om = __shfl_xor_sync(FULL, bml, o);
                const int ok = __shfl_xor_sync(FULL, bulk, o), omr = __shfl_xor_sync(FULL, bmr, o);
                const int onr = __shfl_xor_sync(FULL, bnr, o);
                if (ov > bval || (ov == bval && ov > HNG_NEVSEL && (omr < bmr || (omr == bmr && onr < bnr)))) {
                    bval = ov; bupr = ou; blwr = ol; bml = om; bulk = ok; bmr = omr; bnr = onr;
                }
            }
            if (NW > 1) {
                if ((threadIdx.x & 31) == 0) {
                    int* b = s_best[threadIdx.x >> 5];
                    b[0] = bval; b[1] = bupr; b[2] = blwr; b[3] = bml; b[4] = bulk; b[5] = bmr; b[6] = bnr;
                }
                __syncthreads();
                if (threadIdx.x == 0)
                    for (int w = 1; w < NW; ++w) {
                        const int* b = s_best[w];
                        if (b[0] > bval || (b[0] == bval && b[0] > HNG_NEVSEL && (b[5] < bmr || (b[5] == bmr && b[6] < bnr)))) {
                            bval = b[0]; bupr = b[1]; blwr = b[2]; bml = b[3]; bulk = b[4]; bmr = b[5]; bnr = b[6];
                        }
                    }
            }
        }
        if (lane == 0) {
            int A_left = a_left, A_right = a_right, B_left = b_left, B_right = b_right;
            int mval, mupr, mlwr, mml, mulk;
            const int rr = b_right - 3 * a_right;
            if (LocalR) {
                int i = n_im;
                while (--i >= 0 && MI(i) > A_right) ;
                A_right = bmr; B_right = bnr;
                if (i < 0) i = 0;
                cpos[10 * i + 8] = blwr; cpos[10 * i + 9] = bupr;
                mval = bval; mupr = bupr; mlwr = blwr; mml = bml; mulk = bulk;
            } else {
                const int m3 = 3 * a_right;
                const int rw0 = max(lw, b_left - m3);
                const int r9 = b_right - m3;
                int mxr = r9, mxrow = 0;        // best cell so far: diagonal, band row (0 H, 1 F)
                if (a_exgr) {
                    int glen[3] = {0, 0, 0};
                    int ph = 0;
                    for (int r = rw0; r <= r9; ++r, ph = ph == 2 ? 0 : ph + 1) {
                        const int n = r + m3;
                        HuCell h = hu_ld(Hb + r);
                        glen[ph] += 3;
                        int c0 = h.val, c1 = HNG_NEVSEL, c2 = HNG_NEVSEL;
                        if (r - rw0 >= 3) {
                            const HuCell h3 = hu_ld(Hb + r - 3);
                            if (h3.dir != DEAD) {
                                c1 = h3.val + T.sgf(n - 2, F_SIGE);
                                if (!(a_exgr & 2)) c1 += gap_ext3(P, glen[ph]);
                                if (glen[ph] == 3 && !(a_exgr & 1)) c1 += P.gop;
                                if (P.lcl2 && !(h.dir & SPIN)) c2 = h3.val + T.sgf(n - 2, F_SIGT);
                            }
                        }
                        const int s5 = (Local && T.sgf(n, F_SIG5) > 0) ? T.sgf(n, F_SIG5) : 0;
                        c0 += s5; c1 += s5;
                        const bool self = mxr == r;         // (a cell never beats itself)
                        const int mxv = hu_ld(Hb + mxr).val;
                        int k = 0;
                        if (c1 > c0) k = 1;
                        if (c2 > (k ? c1 : c0)) k = 2;
                        if (k == 0) { if (!h_is_hori(h.dir)) glen[ph] = 0; }
                        else if (k == 1) { h = hu_ld(Hb + r - 3); h.dir = HORI; h.val = c1 - s5; hu_st(Hb + r, h); }
                        else {
                            h = hu_ld(Hb + r - 3);
                            h.dir = DEAD; h.val = c2; h.upr = max(r, h.upr);
                            hu_st(Hb + r, h);
                        }
                        if (!self && h.val > mxv) mxr = r;
                    }
                } else {
                    const HuCell h3 = hu_ld(Hb + r9 - 3);
                    const int y = h3.val + T.sgf(b_right - 2, F_SIGT);
                    if (y > hu_ld(Hb + r9).val) {
                        HuCell h = h3;
                        h.val = y; h.dir = HORI; h.upr = max(r9, h.upr);
                        hu_st(Hb + r9, h);
                    }
                }
                if (b_exgr == 1) {
                    for (int r = min(up, b_right - 3 * a_left); r > r9; --r) {
                        HuCell h = hu_ld(Hb + r);
                        const int x = h.val + (r % 3 ? P.extragop : 0);
                        if (x > hu_ld(Hb + mxr).val) { mxr = r; h.val = x; hu_st(Hb + r, h); }
                    }
                } else if (b_exgr == 2) { mxr = r9; mxrow = 1; }
                const HuCell c = hu_ld((mxrow ? Fb : Hb) + mxr);
                mval = c.val; mupr = c.upr; mlwr = c.lwr; mml = c.ml; mulk = c.ulk;
                const int r = mxrow ? width + mxr : mxr;    // (the reference takes a pointer difference)
                if (b_exgr && rr < r) A_right = (b_right - r) / 3;
                if (a_exgr && rr > r) B_right = 3 * a_right + r;
            }
            int i = n_im;
            while (--i >= 0 && MI(i) > A_right) ;
            if (i < 0 && MI(0) > A_right) cpos[2] = B_right;
            int r = B_right - 3 * A_right;
            cpos[10 * (i + 1) + 8] = min(mlwr, r);
            cpos[10 * (i + 1) + 9] = max(mupr, r);
            r = mulk;
            for ( ; i >= 0 && MI(i) > mml; --i) {
                int c = 0, d = 0;
                if (r > up) { d = (int) (((long long) r - up + width - 1) / width); r -= d * width; }
                if (d >= noll || r < lw - 1) { cpos[10 * i] = END_OF_ULK; break; }      // (a link no pass wrote)
                if (I.at(i, 1, d, r) < END_OF_ULK) {
                    cpos[10 * i + c++] = MI(i);
                    cpos[10 * i + c++] = d > 0 ? 1 : 0;
                    const int mm3 = 3 * MI(i);
                    for (int rp = I.at(i, 0, d, r); lw <= rp && rp < up && r != rp; rp = I.at(i, 0, 0, r = rp))
                        if (c < 7) cpos[10 * i + c++] = r + mm3;
                    if (c < 8) cpos[10 * i + c++] = r + mm3;
                    cpos[10 * i + c] = END_OF_ULK;
                    cpos[10 * i + 8] = I.at(i, 2, d, r);
                    cpos[10 * i + 9] = I.at(i, 3, d, r);
                    r = I.at(i, 1, d, r);
                    if (r == END_OF_ULK) break;
                } else
                    cpos[10 * i] = END_OF_ULK;
            }
            if (r > up) r -= (int) (((long long) r - up + width - 1) / width) * width;
            if (LocalL) {
                A_left = mml;
                B_left = r + 3 * mml;
            } else {
                const int rl_ = B_left - 3 * A_left;
                if (b_exgl && rl_ > r) {
                    A_left = (B_left - r) / 3;
                    for (int j = 0; j < n_im && MI(j) < A_left; ++j) cpos[10 * j] = END_OF_ULK;
                }
                if (a_exgl && rl_ < r) B_left = 3 * A_left + r;
            }
            ++i;
            if ((i < n_im && MI(i) < A_left) || cpos[10 * i + 2] < B_left) mval = NEVSEL32;
            else {
                r = B_left - 3 * A_left;
                cpos[10 * i + 8] = min(r, cpos[10 * i + 8]);
                cpos[10 * i + 9] = max(r, cpos[10 * i + 9]);
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
