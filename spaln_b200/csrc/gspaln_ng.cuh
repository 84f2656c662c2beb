// gspaln_ng.cuh -- the scalar spliced DP kernel with exact intron scoring on the device.
//
// Reference: Aln2s1::trcbkalignS_ng on its scalar branch (src/fwd2s1.cc:1667-1710), i.e.
// forwardS_ng (217-444) with initS_ng / lastS_ng (141-215), the Vmf record store and walk
// (src/vmf.cc:66-140) and the end-point adjustment; intron score SpJunc::spjscr
// (src/codepot.cc:74-77) = IntronPenalty::Penalty(length) + pair-corrected 3' signal
// (Exinon::sig53(IE53), src/codepot.cc:410-415).  The reference takes this branch for every
// block with fewer than 8 query rows, also under -A2 / -A3 (src/fwd2s1.cc:1676); the lsp driver
// meets such blocks between the intermediate rows of a Hirschberg pass.
//
// Shape of the work: at most a handful of query rows against a band, int32, with a sorted list
// of the NCAND best donors per row and a linked list of path records -- sequential in the
// column index by construction.  One THREAD per problem (problems of this kind come in
// hundreds per driver level and are a few thousand cells each); band rows, direction bytes and
// the record store live in a per-thread HBM workspace.  This is the exactness kernel of the
// path, not the throughput kernel (that is dp_wip_kernel).
#pragma once
#include "gspaln_kernels.cuh"

namespace gspaln {

constexpr int NG_NCAND = 4;             // NCAND, src/aln.h:55
constexpr int NG_NEWD = 8;              // Newd, src/fwd2s1.cc:48
constexpr int NG_THREADS = 32;          // threads per CTA (one problem each)
constexpr int NG_NEVSEL = INT_MIN / 16 * 7;     // NEVSEL, src/cmn.h:79

struct NgRvp { int val, ptr; };
struct NgCand { int val, ptr, dir, jnc; };

struct NgWork {                         // per-thread workspace (device pointers)
    NgRvp* band;                        // 3 x width: H | F | F2, index 0 <-> diagonal lw - 1
    unsigned char* dirs;                // width
    int* rec;                           // Vmf records {m, n, prev} x rec_cap
    int rec_cap;
    int n_rec;
    bool overflow;
    __device__ __forceinline__ int add(int m, int n, int p)
    {
        if (n_rec >= rec_cap) { overflow = true; return 0; }
        int* r = rec + 3 * (long long) n_rec;
        r[0] = m; r[1] = n; r[2] = p;
        return n_rec++;
    }
};

// tabs: [0, 544) Exinon::sig53tab, [544, 544 + n_pen) IntronPenalty::Penalty(length)
__device__ __forceinline__ int ng_spjscr(const short* __restrict__ tabs, int n_pen, const ColInfo* cols,
                                         int b_left, int n5, int n3)
{
    const int len = n3 - n5;
    const int pen = tabs[544 + min(len, n_pen - 1)];
    const ColInfo& c5 = cols[n5 - b_left];
    const ColInfo& c3 = cols[n3 - b_left];
    const int d5 = c5.pad[0] & 15, d3 = c3.pad[0] >> 4;
    const short sig = (short) (c3.sig3 - tabs[16 + d3] + tabs[32 + 16 * d5 + d3]);
    return pen + sig;
}

__global__ void __launch_bounds__(NG_THREADS)
dp_ng_kernel(const DevParams* __restrict__ gP, const short* __restrict__ tabs, int n_pen,
             const DevTask* __restrict__ tasks, const int* __restrict__ order, int ntasks, int* ticket,
             const unsigned char* __restrict__ apool, const ColInfo* __restrict__ cpool,
             unsigned char* workpool, long long work_slab, long long width_max, int rec_cap,
             int2* sklpool, DevResult* results, const int* ready)
{
    __shared__ DevParams sP;
    {
        const int* src = reinterpret_cast<const int*>(gP);
        int* dst = reinterpret_cast<int*>(&sP);
        for (int i = threadIdx.x; i < (int) (sizeof(DevParams) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const DevParams& P = sP;
    const bool dagp = P.noll == 3, spj = P.spj != 0;
    const int nod = 2 * P.noll - 1;
    const int gop_k[3] = {0, P.gop, P.lgop};                // PwdB::GOP, src/aln2.cc:111
    const int psp_bit[5] = {4, 1, 8, 2, 16};                // src/aln.h:56

    NgWork W;
    {
        unsigned char* base = workpool + ((long long) blockIdx.x * NG_THREADS + threadIdx.x) * work_slab;
        W.band = reinterpret_cast<NgRvp*>(base);
        W.dirs = base + 3 * width_max * (long long) sizeof(NgRvp);
        W.rec = reinterpret_cast<int*>(W.dirs + ((width_max + 15) / 16) * 16);
        W.rec_cap = rec_cap;
    }

    for (;;) {
        const int tk = atomicAdd(ticket, 1);
        if (tk >= ntasks) break;
        const int ti = order[tk];
        const DevTask t = tasks[ti];
        if (t.kind != 3) continue;                          // handled by another kernel
        if (ready) {                                        // streamed batch: wait for the inputs
            const volatile int* r = ready;
            const long long t0 = clock64();
            bool ok = true;
            while (*r <= tk) {
                __nanosleep(256);
                if (clock64() - t0 > (1ll << 33)) { ok = false; break; }
            }
            if (!ok) { DevResult rr; rr.score = 0; rr.status = 4; rr.n_skl = 0; rr.pad = 0; results[ti] = rr; continue; }
        }
        const unsigned char* aseq = apool + t.a_off;        // aseq[i] pairs query row a_left + i + 1
        const ColInfo* cols = cpool + t.col_off;            // cols[j] is column b_left + j
        const int width = t.up - t.lw + 3;
        const bool a_exgl = t.flags & 1, a_exgr = t.flags & 2, b_exgl = t.flags & 4, b_exgr = t.flags & 8;
        const bool LocalL = P.local && a_exgl && b_exgl, LocalR = P.local && a_exgr && b_exgr;
        const int a_left = t.a_left, a_right = t.a_right, b_left = t.b_left, b_right = t.b_right;
        const int lw = t.lw, up = t.up;
        NgRvp* H = W.band - lw + 1;                         // by diagonal r = n - m in [lw - 1, up + 1]
        NgRvp* F = H + width;
        NgRvp* F2 = F + width;
        unsigned char* dirs = W.dirs - lw + 1;
        for (int i = 0; i < 3 * width; ++i) W.band[i] = NgRvp{NG_NEVSEL, 0};
        for (int i = 0; i < width; ++i) W.dirs[i] = 0;
        W.n_rec = 0; W.overflow = false;
        W.add(0, 0, 0);                                     // record 0 is never a path node

        // ---- initS_ng (src/fwd2s1.cc:141-184)
        {
            int r = b_left - a_left, rr = b_right - a_left;
            H[r].val = 0; dirs[r] = 0;
            H[r].ptr = W.add(a_left, b_left, 0);
            if (a_exgl) {
                if (up < rr) rr = up;
                while (++r <= rr) { H[r] = NgRvp{0, 0}; dirs[r] = 1; }
            }
            r = b_left - a_left;
            rr = max(b_left - a_right, lw);
            for (int i = 1; --r >= rr; ++i) {
                dirs[r] = 2;
                if (b_exgl) H[r] = NgRvp{0, 0};
                else {
                    NgRvp v = H[r + 1];
                    v.val += i == 1 ? P.gappen1 : (i > P.codonk1 ? P.lgep : P.gep);
                    H[r] = v;
                }
            }
        }

        int best_val = NG_NEVSEL, best_m = a_left, best_n = b_left, best_p = 0;     // LocalR
        for (int m = a_exgl ? a_left + 1 : a_left; m <= a_right; ++m) {
            const bool internal = spj && (!a_exgr || m < a_right);
            const bool first = m == a_left;                 // global start row: horizontal moves only
            int n = max((m - 1) + lw, b_left);
            const int n9 = min((m - 1) + up + 1, b_right);
            const int arow = first ? ZROW : (int) aseq[m - 1 - a_left];
            NgRvp e1{NG_NEVSEL, 0}, e2{NG_NEVSEL, 0};
            NgCand rcd[NG_NCAND + 1];
            int idx[NG_NCAND + 1];
#pragma unroll
            for (int l = 0; l <= NG_NCAND; ++l) { rcd[l] = NgCand{NG_NEVSEL, 0, 0, 0}; idx[l] = l; }
            int ncand = -1, psp = 0;
            NgRvp hleft = H[n - m];                         // H[m][n] of the previous column
            while (++n <= n9) {
                const int r = n - m;
                const ColInfo col = cols[n - b_left];
                // cell state: 0 H, 1 E1, 2 F, 3 E2, 4 F2 (the reference's hf[] order)
                NgRvp st[5];
                st[0] = H[r]; st[1] = e1; st[2] = F[r]; st[3] = e2; st[4] = dagp ? F2[r] : NgRvp{NG_NEVSEL, 0};
                int mx = 0;
                const int diag = st[0].val;
                int dir = dirs[r];
                if (!first) {
                    st[0].val += P.mtxT[(int) col.code * MTX_LD + arow];
                    dir = (dir % NG_NEWD) ? NG_NEWD : 0;
                    const NgRvp up_h = H[r + 1];
                    const NgRvp up_f = F[r + 1];
                    int x = up_h.val + P.gop;
                    if (x >= up_f.val) st[2] = NgRvp{x, up_h.ptr}; else st[2] = up_f;
                    st[2].val += P.gep;
                    if (st[2].val > st[mx].val) mx = 2;
                    if (dagp) {
                        const NgRvp up_f2 = F2[r + 1];
                        x = up_h.val + P.lgop;
                        if (x >= up_f2.val) st[4] = NgRvp{x, up_h.ptr}; else st[4] = up_f2;
                        st[4].val += P.lgep;
                        if (st[4].val > st[mx].val) mx = 4;
                    }
                }
                {
                    int x = hleft.val + P.gop;
                    const int prev_psp = psp;
                    if (x >= st[1].val) { st[1] = NgRvp{x, hleft.ptr}; psp = psp ? 1 : 0; }
                    else psp &= 1;
                    st[1].val += P.gep;
                    if (st[1].val >= st[mx].val) mx = 1;
                    if (dagp) {
                        x = hleft.val + P.lgop;
                        if (x >= st[3].val) { st[3] = NgRvp{x, hleft.ptr}; if (prev_psp) psp |= 2; }
                        else psp |= prev_psp & 2;
                        st[3].val += P.lgep;
                        if (st[3].val >= st[mx].val) mx = 3;
                    }
                }
                const int cano5 = col.pad[1] & 15, cano3 = col.pad[1] >> 4;
                // acceptor: every stored donor of this row, per gap state
                if (internal && cano3) {
                    int top[5] = {-1, -1, -1, -1, -1};
                    for (int l = 0; l <= ncand; ++l) {
                        const NgCand& c = rcd[idx[l]];
                        if (n - c.jnc < P.llmt) continue;
                        const int x = c.val + ng_spjscr(tabs, n_pen, cols, b_left, c.jnc, n);
                        if (x >= st[c.dir].val) { st[c.dir].val = x; top[c.dir] = idx[l]; }
                    }
                    for (int k = 0; k < nod; ++k) {
                        if (top[k] < 0) continue;
                        const NgCand& c = rcd[top[k]];
                        psp |= psp_bit[k];
                        const int inner = W.add(m, c.jnc, c.ptr);
                        st[k].ptr = W.add(m, n, inner);
                        if (st[k].val >= st[mx].val) mx = k;
                    }
                }
                // best state
                const int hd = mx;
                const int mxval = st[mx].val;               // the reference reads mx->val later on
                if (mx != 0) {
                    st[0] = st[mx];
                    dir = hd;
                } else if (P.local && st[0].val > diag) {
                    if (LocalL && diag == 0) st[0].ptr = W.add(m - 1, n - 1, 0);
                    else if (LocalR && st[0].val > best_val) {
                        best_val = st[0].val; best_p = st[0].ptr; best_m = m; best_n = n;
                    }
                }
                int mx_now = mxval;
                if (LocalL && st[0].val <= 0) { st[0].val = 0; dir = 1; if (mx == 0) mx_now = 0; }
                else if (dir == NG_NEWD && !(psp & psp_bit[0])) st[0].ptr = W.add(m - 1, n - 1, st[0].ptr);
                // donor: keep the NCAND best (value + 5' signal) of this row
                if (internal && cano5) {
                    const int sigJ = col.sig5;
                    for (int k = hd == 0 ? 0 : 1; k < nod; ++k) {
                        if (psp & psp_bit[k]) continue;
                        const NgRvp from = st[k];
                        if (k != hd) {
                            int z = mx_now;
                            if (hd == 0 || (k - hd) % 2) z += gop_k[k / 2];
                            if (from.val <= z) continue;
                        }
                        const int x = from.val + sigJ;
                        int l = ncand < NG_NCAND ? ++ncand : NG_NCAND;
                        while (--l >= 0) {
                            if (x > rcd[idx[l]].val) { const int s = idx[l]; idx[l] = idx[l + 1]; idx[l + 1] = s; }
                            else break;
                        }
                        if (++l < NG_NCAND) rcd[idx[l]] = NgCand{x, from.ptr, k, n};
                        else --ncand;
                    }
                }
                H[r] = st[0]; F[r] = st[2];
                if (dagp) F2[r] = st[4];
                dirs[r] = (unsigned char) dir;
                e1 = st[1]; e2 = st[3];
                hleft = st[0];
            }
        }

        int ptr, val;
        if (LocalR) {
            ptr = W.add(best_m, best_n, best_p);
            val = best_val;
        } else {
            // lastS_ng (src/fwd2s1.cc:186-215)
            int rw = max(lw, b_left - a_right);
            const int r9 = b_right - a_right;
            int mxr = r9;
            if (a_exgr)
                for (int r = rw; r <= r9; ++r) if (H[r].val > H[mxr].val) mxr = r;
            if (b_exgr) {
                rw = min(up, b_right - a_left);
                for (int r = rw; r > r9; --r) if (H[r].val > H[mxr].val) mxr = r;
            }
            const int i = mxr - r9;
            int m9 = a_right, n9 = b_right;
            if (i > 0) m9 -= i;
            if (i < 0) n9 += i;
            ptr = W.add(m9, n9, H[mxr].ptr);
            val = H[mxr].val;
        }

        // ---- Vmf::traceback + the start-point adjustment of trcbkalignS_ng
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
            const int rd = P.local ? 0 : (n_last - m_last) - b_left + a_left;
            if (rd) {
                if (cnt < t.skl_cap)
                    skl[cnt] = rd > 0 ? make_int2(a_left, b_left + rd) : make_int2(a_left - rd, b_left);
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


// ---------------------------------------------------------------------------
// Aln2s1::scorealoneS_ng (src/fwd2s1.cc:1163-1336) with sinitS_ng / slastS_ng (1112-1161): the
// scalar score-only kernel (HomScoreS_ng under -A0 and for queries shorter than 4 residues,
// src/fwd2s1.cc:2704-2705).  Same frame as dp_ng_kernel (one thread per problem) but no path
// records: the workspace is three int rows of the band width, so it also takes full-size problems.
// Its tie rules differ from forwardS_ng's (strict comparisons for the gap states and the
// acceptors, ties accepted in the donor list) and so do its initial rows.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(NG_THREADS)
dp_ng_score_kernel(const DevParams* __restrict__ gP, const short* __restrict__ tabs, int n_pen,
                   const DevTask* __restrict__ tasks, const int* __restrict__ order, int ntasks, int* ticket,
                   const unsigned char* __restrict__ apool, const ColInfo* __restrict__ cpool,
                   int* workpool, long long width_max, DevResult* results, const int* ready)
{
    __shared__ DevParams sP;
    {
        const int* src = reinterpret_cast<const int*>(gP);
        int* dst = reinterpret_cast<int*>(&sP);
        for (int i = threadIdx.x; i < (int) (sizeof(DevParams) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const DevParams& P = sP;
    const bool dagp = P.noll == 3;
    const int nod = 2 * P.noll - 1;
    const int gop_k[3] = {0, P.gop, P.lgop};
    const int psp_bit[5] = {4, 1, 8, 2, 16};
    int* wbase = workpool + ((long long) blockIdx.x * NG_THREADS + threadIdx.x) * 3 * width_max;

    for (;;) {
        const int tk = atomicAdd(ticket, 1);
        if (tk >= ntasks) break;
        const int ti = order[tk];
        const DevTask t = tasks[ti];
        if (t.kind != 4) continue;
        if (ready) {
            const volatile int* r = ready;
            const long long t0 = clock64();
            bool ok = true;
            while (*r <= tk) {
                __nanosleep(256);
                if (clock64() - t0 > (1ll << 33)) { ok = false; break; }
            }
            if (!ok) { DevResult rr; rr.score = 0; rr.status = 4; rr.n_skl = 0; rr.pad = 0; results[ti] = rr; continue; }
        }
        const unsigned char* aseq = apool + t.a_off;
        const ColInfo* cols = cpool + t.col_off;
        const int width = t.up - t.lw + 3;
        const bool a_exgl = t.flags & 1, a_exgr = t.flags & 2, b_exgl = t.flags & 4, b_exgr = t.flags & 8;
        const bool LocalL = P.local && a_exgl && b_exgl, LocalR = P.local && a_exgr && b_exgr;
        const int a_left = t.a_left, a_right = t.a_right, b_left = t.b_left, b_right = t.b_right;
        const int lw = t.lw, up = t.up;
        int* H = wbase - lw + 1;
        int* F = H + width;
        int* F2 = F + width;
        for (int i = 0; i < 3 * width; ++i) wbase[i] = NG_NEVSEL;
        // ---- sinitS_ng
        {
            int r = b_left - a_left, rr = b_right - a_left;
            H[r] = 0;
            if (a_exgl) {
                if (up < rr) rr = up;
                for (int q = r + 1; q <= rr; ++q) H[q] = 0;
            }
            rr = max(b_left - a_right, lw);
            if (b_exgl) {
                for (int q = rr; q < r; ++q) H[q] = 0;
            } else {
                for (int i = 1; --r >= rr; ++i) {
                    int h = H[r + 1];
                    if (i == 1) { h += P.gappen1; F[r] = h; }
                    else { h += i > P.codonk1 ? P.lgep : P.gep; F[r] = F[r + 1] + P.gep; }
                    H[r] = h;
                }
            }
        }
        int maxh = NG_NEVSEL;
        for (int m = a_exgl ? a_left + 1 : a_left; m <= a_right; ++m) {
            const bool first = m == a_left;
            int n = max((m - 1) + lw, b_left);
            const int n9 = min((m - 1) + up + 1, b_right);
            const int arow = first ? ZROW : (int) aseq[m - 1 - a_left];
            int e1 = NG_NEVSEL, e2 = NG_NEVSEL;
            int cval[NG_NCAND + 1], cdir[NG_NCAND + 1], cjnc[NG_NCAND + 1], idx[NG_NCAND + 1];
#pragma unroll
            for (int l = 0; l <= NG_NCAND; ++l) { cval[l] = NG_NEVSEL; cdir[l] = 0; cjnc[l] = 0; idx[l] = l; }
            int ncand = -1, psp = 0;
            int hleft = H[n - m];
            while (++n <= n9) {
                const int r = n - m;
                const ColInfo col = cols[n - b_left];
                int st[5];                              // H, E1, F, E2, F2 (the reference's hf[] order)
                st[0] = H[r]; st[1] = e1; st[2] = F[r]; st[3] = e2; st[4] = dagp ? F2[r] : NG_NEVSEL;
                int mx = 0;
                if (!first) {
                    st[0] += P.mtxT[(int) col.code * MTX_LD + arow];
                    const int up_h = H[r + 1];
                    st[2] = max(up_h + P.gop, F[r + 1]) + P.gep;
                    if (st[2] > st[mx]) mx = 2;
                    if (dagp) {
                        st[4] = max(up_h + P.lgop, F2[r + 1]) + P.lgep;
                        if (st[4] > st[mx]) mx = 4;
                    }
                }
                {
                    int x = hleft + P.gop;
                    const int prev_psp = psp;
                    if (x > st[1]) { st[1] = x; psp = psp ? 1 : 0; }
                    else psp &= 1;
                    st[1] += P.gep;
                    if (st[1] > st[mx]) mx = 1;
                    if (dagp) {
                        x = hleft + P.lgop;
                        if (x > st[3]) { st[3] = x; if (prev_psp) psp |= 2; }
                        else psp |= prev_psp & 2;
                        st[3] += P.lgep;
                        if (st[3] > st[mx]) mx = 3;
                    }
                }
                const int cano5 = col.pad[1] & 15, cano3 = col.pad[1] >> 4;
                if (cano3) {
                    unsigned top = 0;
                    for (int l = 0; l <= ncand; ++l) {
                        const int j = idx[l];
                        if (n - cjnc[j] < P.llmt) continue;
                        const int x = cval[j] + ng_spjscr(tabs, n_pen, cols, b_left, cjnc[j], n);
                        if (x > st[cdir[j]]) { st[cdir[j]] = x; top |= 1u << cdir[j]; }
                    }
                    for (int k = 0; k < nod; ++k) {
                        if (!(top >> k & 1u)) continue;
                        psp |= psp_bit[k];
                        if (st[k] > st[mx]) mx = k;
                    }
                }
                const int y = st[0];
                const int mxval = st[mx];
                if (mx != 0) st[0] = mxval;
                else if (LocalR && y > maxh) maxh = y;
                int mx_now = mxval;
                if (LocalL && st[0] < 0) { st[0] = 0; if (mx == 0) mx_now = 0; }
                const int hd = mx;
                if (cano5) {
                    const int sigJ = col.sig5;
                    for (int k = hd == 0 ? 0 : 1; k < nod; ++k) {
                        if (psp & psp_bit[k]) continue;
                        const int from = k == 0 ? st[0] : st[k];
                        if (k != hd) {
                            int z = mx_now;
                            if (hd == 0 || (k - hd) % 2) z += gop_k[k / 2];
                            if (from <= z) continue;
                        }
                        const int x = from + sigJ;
                        int l = ncand < NG_NCAND ? ++ncand : NG_NCAND;
                        while (--l >= 0) {
                            if (x >= cval[idx[l]]) { const int s = idx[l]; idx[l] = idx[l + 1]; idx[l + 1] = s; }
                            else break;
                        }
                        if (++l < NG_NCAND) { cval[idx[l]] = x; cjnc[idx[l]] = n; cdir[idx[l]] = k; }
                        else --ncand;
                    }
                }
                H[r] = st[0]; F[r] = st[2];
                if (dagp) F2[r] = st[4];
                e1 = st[1]; e2 = st[3];
                hleft = st[0];
            }
        }
        if (!LocalR) {
            // slastS_ng
            const int r9 = b_right - a_right;
            int mxv = H[r9];
            if (b_exgr) {
                const int rw = min(up, b_right - a_left);
                for (int r = rw; r > r9; --r) mxv = max(mxv, H[r]);
            }
            if (a_exgr) {
                const int rw = max(lw, b_left - a_right);
                for (int r = rw; r < r9; ++r) mxv = max(mxv, H[r]);
            }
            maxh = mxv;
        }
        DevResult res;
        res.score = maxh; res.status = 0; res.n_skl = 0; res.pad = 0;
        results[ti] = res;
    }
}

}   // namespace gspaln
