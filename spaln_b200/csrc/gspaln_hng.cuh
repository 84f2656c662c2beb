// gspaln_hng.cuh -- exact intron-length DP for protein queries (the reference's scalar
// formulation) as a warp kernel.
//
// Semantics: bit-identical to Aln2h1::trcbkalignH_ng on its scalar branch (src/fwd2h1.cc:1997-2041):
// forwardH_ng (294-617) with initH_ng / lastH_ng (143-292), the Vmf path records
// (src/vmf.cc:66-140), exact intron scoring SpJunc::spjscr and the split-codon translation
// SpJunc::spjseq (src/codepot.cc:74-102).  The reference takes this branch for every block with
// fewer than 8 query rows (src/fwd2h1.cc:2007).
//
// Mapping (NOT the reference's row-by-row loop): one warp per problem; the 32 lanes own 32
// consecutive query rows and advance in lock step, each lane ONE genome column behind the lane of
// the row above.  With that skew the band rows in global memory -- {value, record, direction} per
// diagonal r = n - 3m for H, F and F2, the reference's hh[0..2][] arrays -- can be shared by all
// lanes without any hazard: at a step the lane of row m writes diagonal r only, reads r .. r + 3 of
// the row above (written one to four steps earlier) and r - 1 .. r - 3 of its own row, while the
// lane below is at r - 4 and the lane above at r + 4; one __syncwarp() per step orders the
// stores.  Every quirk of the reference that comes from re-reading band entries nobody overwrote is
// reproduced for free, because the arrays ARE the reference's arrays.  What a row carries along its
// columns (three-slot rings of the two horizontal gap states, three donor lists by splice phase)
// lives in the lane; path records go to a per-problem store in lane-private chunks.  The start
// row / column and the end-point search are short serial loops on lane 0.
#pragma once
#include "gspaln_kernels.cuh"

namespace gspaln {

constexpr int HNG_WARPS = 4;                        // problems in flight per CTA
constexpr int HNG_THREADS = 32 * HNG_WARPS;
constexpr int HNG_NEVSEL = INT_MIN / 16 * 7;        // NEVSEL, src/cmn.h:79
constexpr int HNG_NCAND = 4;                        // NCAND, src/aln.h:55
constexpr int HNG_CHUNK = 32;                       // path records a lane reserves at a time
// direction codes, src/aln.h:30-35
enum { DEAD, RSRV, DIAG, NEWD, VERT, SLA1, SLA2, VERL, HORI, HOR1, HOR2, HORL, NEWV, NEWH, SPIN = 16 };
// gap state of a direction code (0 H, 1 E, 2 F, 3 E2, 4 F2; -1 none) and back
__constant__ int c_dir2nod[16] = {-1, -1, 0, 0, 2, 2, 2, 4, 1, 1, 1, 3, 2, 1, -1, -1};
__constant__ int c_nod2dir[5] = {DIAG, HORI, VERT, HORL, VERL};
__constant__ unsigned char c_hncred[17] = {15, 15, 0, 1, 4, 2, 5, 6, 10, 3, 7, 8, 10, 9, 12, 13, 14};

__device__ __forceinline__ bool h_is_diag(int d) { d &= 15; return d == DIAG || d == NEWD; }
__device__ __forceinline__ bool h_is_vert(int d) { d &= 15; return (d >= VERT && d <= VERL) || d == NEWV; }
__device__ __forceinline__ bool h_is_hori(int d) { d &= 15; return (d >= HORI && d <= HORL) || d == NEWH; }

struct __align__(16) HCell { int val, ptr, dir, pad; };     // RVPD: value, path record, direction

struct DevNgHParams {           // frozen scalars + device pointers of the tables
    int gop, gep, lgop, lgep, codonk1, gw1, gw2, gw3, gw3l, gape1, gape2, extragop;
    int local, spj, noll, minl, simdim, n_penalty;
    int lcl2, pad_;             // algmode.lcl & 2 (hlastH_ng's termination-codon candidate)
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
    long long cip_off;          // Cip_score words from coding position 3 a_left - 1 on; -1: none
    int wide, pad_;             // 1: the problem runs on a CTA of HNG_WIDE warps
};

enum { F_SIG5, F_SIG3, F_SIGS, F_SIGT, F_SIGE, F_SIGI, F_PHS5, F_PHS3 };

// the inputs of one problem with the reference's indexing (positions, not offsets)
struct HngIn {
    const unsigned char* a; const unsigned char* b; const short* sg; const unsigned short* i53;
    const int* cip; int cip_lo;     // Cip_score::cip_score(c) = cip[c - cip_lo] (nullptr: none)
    int a_lo, b_lo, b_left, b_right;
    __device__ __forceinline__ int aa(int m) const { return a[m - a_lo]; }
    __device__ __forceinline__ int tron(int n) const { return b[n - b_lo]; }
    __device__ __forceinline__ int sgf(int n, int f) const { return sg[8 * ((long long) n - b_lo) + f]; }
    __device__ __forceinline__ int int53(int n) const { return i53[n - b_lo]; }
};

// lane-private slice of the problem's record store
struct HngAlloc {
    int* rec; int cap;
    int* next;                  // shared counter of the warp
    int cur, end;
    bool overflow;
    __device__ __forceinline__ int add(int m, int n, int prev)
    {
        if (overflow) return 0;
        if (cur == end) { cur = atomicAdd(next, HNG_CHUNK); end = cur + HNG_CHUNK; }
        if (end > cap) { overflow = true; return 0; }
        int* r = rec + 3 * (long long) cur;
        r[0] = m; r[1] = n; r[2] = prev;
        return cur++;
    }
};

__device__ __forceinline__ HCell hx_ld(const HCell* p)
{
    const int4 v = __ldcg(reinterpret_cast<const int4*>(p));    // through L2: written by other lanes
    return HCell{v.x, v.y, v.z, 0};
}
__device__ __forceinline__ void hx_st(HCell* p, const HCell& c)
{
    *reinterpret_cast<int4*>(p) = make_int4(c.val, c.ptr, c.dir, 0);
}

__device__ __forceinline__ int gap_ext3(const DevNgHParams& P, int i) { return i > P.codonk1 ? P.lgep : P.gep; }

// SpJunc::spjseq (src/codepot.cc:79-102): the two residues a codon split by the intron (n5, n3)
// translates to, from the two nucleotides before the donor and the two after the acceptor
__device__ const unsigned char* hx_spjseq(const DevNgHParams& P, const HngIn& T, int n5, int n3)
{
    const unsigned char* tab = P.spj_tabs, *amb_tron = tab + 514, *tron_amb = tab + 514 + 128, *aa2nuc = tab + 514 + 256;
    if (n5 < T.b_left || n3 >= T.b_right) return tab + 2 * 256;
    auto nuc = [&](int pos) -> int { const unsigned c = T.tron(pos); const unsigned v = aa2nuc[c < 26 ? c : 0]; return c_hncred[v < 17 ? v : 0]; };
    int amb = 0;
    int c = nuc(n5 - 2);
    if (c >= 4) { amb = 1; c = 0; }
    unsigned w = (unsigned) c;
    if ((c = nuc(n5 - 1)) < 4) {
        w = 4 * w + c;
        if ((c = nuc(n3)) < 4) {
            w = 4 * w + c;
            if ((c = nuc(n3 + 1)) < 4) w = 4 * w + c;
            else if (amb) w = 256;
            else amb = 2;
        } else w = 256;
    } else w = 256;
    if (amb == 0 || w == 256) return tab + 2 * w;
    return (amb == 1 ? amb_tron : tron_amb) + 2 * w;
}

// SpJunc::spjscr: length penalty + pair-corrected 3' signal
__device__ __forceinline__ int hx_spjscr(const DevNgHParams& P, const HngIn& T, int n5, int n3)
{
    const int len = n3 - n5;
    const int pen = P.penalty[len < 0 ? 0 : (len < P.n_penalty ? len : P.n_penalty - 1)];
    const int d5 = T.int53(n5) & 15, d3 = (T.int53(n3) >> 4) & 15;
    return pen + (short) (T.sgf(n3, F_SIG3) - P.sig53tab[16 + d3] + P.sig53tab[32 + 16 * d5 + d3]);
}

// donor list of one row and splice phase: sorted best first, an equal newcomer passes the entries
// it ties with; the entry pushed out of the best NCAND survives in slot NCAND until the next
// insertion attempt (src/fwd2h1.cc:553-566 keeps NCAND + 1 slots behind an index permutation)
struct HxList {
    int val[HNG_NCAND + 1], ptr[HNG_NCAND + 1], jnc[HNG_NCAND + 1], st[HNG_NCAND + 1];
    int n;
    __device__ void clear() { n = 0; for (int l = 0; l <= HNG_NCAND; ++l) { val[l] = HNG_NEVSEL; ptr[l] = jnc[l] = st[l] = 0; } }
    __device__ void insert(int x, int p, int state, int j)
    {
        if (n > HNG_NCAND) n = HNG_NCAND;
        int pos = 0;
        while (pos < n && pos < HNG_NCAND && val[pos] > x) ++pos;
        if (pos >= HNG_NCAND) return;
        for (int l = HNG_NCAND; l > pos; --l) { val[l] = val[l - 1]; ptr[l] = ptr[l - 1]; jnc[l] = jnc[l - 1]; st[l] = st[l - 1]; }
        val[pos] = x; ptr[pos] = p; jnc[pos] = j; st[pos] = state;
        ++n;
    }
};

// NW = warps per problem: 1 (a CTA runs HNG_WARPS independent problems) or HNG_WIDE (queries of
// HNG_WIDE_ROWS residues and more: one problem per CTA, the skewed wavefront 32 NW rows tall, the
// per-step barrier the CTA's -- the lanes only ever meet through the band rows, so nothing else changes)
constexpr int HNG_WIDE = 8;
constexpr int HNG_WIDE_ROWS = 96;

template <int NW>
__global__ void __launch_bounds__(NW == 1 ? HNG_THREADS : 32 * NW)
dp_hxild_kernel(const DevNgHParams* __restrict__ gP, const DevNgHTask* __restrict__ tasks, int ntasks, int* ticket,
                const unsigned char* __restrict__ apool, const unsigned char* __restrict__ bpool,
                const short* __restrict__ sgpool, const unsigned short* __restrict__ i53pool,
                const int* __restrict__ cippool, unsigned char* workpool, int2* sklpool, DevResult* results)
{
    constexpr int NT = 32 * NW;
    constexpr int NWARP = NW == 1 ? HNG_WARPS : NW;
    __shared__ DevNgHParams P;
    __shared__ int warp_next[NWARP];
    __shared__ int s_ti, s_best[NWARP][4];
    if (threadIdx.x < sizeof(DevNgHParams) / 4)
        reinterpret_cast<int*>(&P)[threadIdx.x] = reinterpret_cast<const int*>(gP)[threadIdx.x];
    __syncthreads();
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x % NT;                  // lane of the problem's wavefront
    const int wid = NW == 1 ? (threadIdx.x >> 5) : 0;   // problem slot of the CTA
    auto sync_problem = [] { if (NW == 1) __syncwarp(); else __syncthreads(); };
    const HCell black = {HNG_NEVSEL, 0, 0, 0};

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
        const DevNgHTask t = tasks[ti];
        if ((t.wide != 0) != (NW > 1)) continue;            // the other class runs it
        HngIn T;
        T.a = apool + t.a_off; T.b = bpool + t.b_off; T.sg = sgpool + 8 * t.sg_off; T.i53 = i53pool + t.sg_off;
        T.a_lo = t.a_lo; T.b_lo = t.b_lo; T.b_left = t.b_left; T.b_right = t.b_right;
        T.cip = t.cip_off >= 0 ? cippool + t.cip_off : nullptr; T.cip_lo = 3 * t.a_left - 1;
        const int a_exgl = t.a_exgl, a_exgr = t.a_exgr, b_exgl = t.b_exgl, b_exgr = t.b_exgr;
        const int width = t.up - t.lw + 7;
        const int nod = 2 * P.noll - 1;
        const bool dagp = P.noll == 3;
        const bool Local = P.local, LocalL = Local && a_exgl && b_exgl, LocalR = Local && a_exgr && b_exgr;
        const int a_left = t.a_left, a_right = t.a_right, b_left = t.b_left, b_right = t.b_right;
        const int lw = t.lw, up = t.up;
        // band rows by diagonal r = n - 3 m in [lw - 3, up + 3]
        HCell* buf = reinterpret_cast<HCell*>(workpool + t.work_off);
        HCell* Hb = buf - lw + 3;
        HCell* Fb = Hb + width;
        HCell* F2b = Fb + width;
        HngAlloc A;
        A.rec = reinterpret_cast<int*>(buf + 3 * (width + 8));
        A.cap = t.rec_cap; A.next = &warp_next[wid]; A.cur = A.end = 0; A.overflow = false;
        if (lane == 0) warp_next[wid] = HNG_CHUNK;      // record 0 is never a path node
        for (int i = lane; i < 3 * (width + 8); i += NT) hx_st(buf + i, black);
        sync_problem();

        // ---- start row and start column (initH_ng, src/fwd2h1.cc:143-222): serial, lane 0
        if (lane == 0) {
            auto sigS = [&](int n) { const int s = T.sgf(n, F_SIGS); return s > 0 ? s : 0; };
            const int r0 = b_left - 3 * a_left;
            const int dir0 = a_exgl ? DEAD : DIAG;
            hx_st(Hb + r0, HCell{a_exgl ? sigS(b_left + 1) : 0, A.add(a_left, b_left, 0), dir0, 0});
            if (a_exgl) {
                // free start anywhere on the first row: per reading frame either a codon-wise
                // extension of the frame's last start (gap open only for the first codon unless the
                // end is free) or a new start at this column, whichever scores more with its
                // start-codon signal
                int jnc[3] = {b_left, 0, 0};
                const int rr = min(up, b_right - 3 * a_left);
                for (int i = 1, r = r0 + 1; r <= rr; ++r, ++i) {
                    const int n = b_left + i;
                    HCell h;
                    if (i < 3) {
                        h = HCell{sigS(n + 1), A.add(a_left, n, 0), dir0, 0};
                        jnc[i] = n;
                    } else {
                        h = *(Hb + r - 3);
                        const int k = n - jnc[i % 3];
                        if (k == 3 && !(a_exgl & 1)) h.val += P.gop;
                        if (!(a_exgl & 2)) h.val += gap_ext3(P, k);
                        h.val += T.sgf(n + 1 - 3, F_SIGE);
                        h.dir = HORI;
                        const HCell h1 = *(Hb + r - 1), h2 = *(Hb + r - 2);
                        if (h1.val + P.gw1 > h.val) { h = h1; h.val += P.gw1; h.dir = HOR1; }
                        if (h2.val + P.gw2 > h.val) { h = h2; h.val += P.gw2; h.dir = HOR2; }
                    }
                    const int xs = sigS(n + 1);
                    if (h.val < xs) { h = HCell{xs, A.add(a_left, n, 0), DEAD, 0}; jnc[i % 3] = n; }
                    hx_st(Hb + r, h);
                }
            }
            // leading query residues against nothing: free, or a gap opened at the corner (the
            // first three rows come from the corner itself and pay the frame-shift surcharge)
            const int rr = max(lw, b_left - 3 * a_right);
            for (int i = 1, r = r0 - 1; r >= rr; --r, ++i) {
                HCell h;
                if (b_exgl == 1) h = HCell{0, 0, DEAD, 0};
                else if (i <= 3) {
                    h = *(Hb + r + i);
                    if (!(b_exgl & 2)) h.val += P.gep;
                    if (!(b_exgl & 1)) h.val += P.gop;
                    if (i < 3) h.val += P.extragop;
                    h.dir = VERT;
                } else {
                    h = *(Hb + r + 3);
                    if (!(b_exgl & 2)) h.val += gap_ext3(P, i);
                }
                hx_st(Hb + r, h);
            }
        }
        __threadfence_block();
        sync_problem();

        int best_val = HNG_NEVSEL, best_m = a_left, best_n = b_left, best_p = 0;    // LocalR (lane-local)
        const int m_first = a_exgl ? a_left + 1 : a_left;
        for (int m0 = m_first; m0 <= a_right; m0 += NT) {
            const int m = m0 + lane;
            const bool row = m <= a_right;
            const int n0 = max(3 * m + lw - 1, b_left), n9 = min(3 * m + up, b_right);
            const int last_lane = min(NT - 1, a_right - m0);
            const int s_begin = max(3 * m0 + lw - 1, b_left);
            const int s_end = min(3 * (m0 + last_lane) + up, b_right) + last_lane;
            // horizontal gap states by column phase (three-slot rings), donor lists by splice phase
            HCell e1[3] = {black, black, black}, e2[3] = {black, black, black};
            int q = 0;
            HxList don[3];
            don[0].clear(); don[1].clear(); don[2].clear();
            // (lanes past the last row own nothing: they must not touch the query array)
            const int* prof_prev = P.mtx + (row ? T.aa(m > 0 ? m - 1 : 0) : 0) * P.simdim;  // residue the row pairs
            const int* prof_next = P.mtx + (row ? T.aa(m) : 0) * P.simdim;                  // the one after it
            bool started = false;
            // bonus of an intron conserved with the query's annotation, by splice phase -1, 0, 1
            int sigB[3] = {0, 0, 0};
            if (T.cip && row)
                for (int phs = -1; phs < 2; ++phs) sigB[phs + 1] = T.cip[3 * m - phs - T.cip_lo];

            for (int s = s_begin; s <= s_end; ++s) {
                const int n = s - lane;
                if (row && n >= n0 && n <= n9) {
                    const int r = n - 3 * m;
                    if (!started) {
                        started = true;
                        if (!b_exgl && m == a_left) {
                            // global start row: the horizontal states continue the corner's gap
                            const HCell c = hx_ld(Hb + r);
                            e1[2] = HCell{P.gw3, c.ptr, c.dir, 0};
                            e2[2] = HCell{P.gw3l, c.ptr, c.dir, 0};
                        }
                    }
                    const int sigE = n > b_left ? T.sgf(n - 2, F_SIGE) : 0;
                    const HCell hq = hx_ld(Hb + r);             // cell (m - 1, n - 3) or a start value
                    // the five gap states of the cell: 0 H, 1 E, 2 F, 3 E2, 4 F2
                    HCell st[5];
                    st[0] = hq; st[1] = e1[q]; st[2] = hx_ld(Fb + r); st[3] = e2[q];
                    st[4] = dagp ? hx_ld(F2b + r) : black;
                    int mx = 0;
                    if (m != a_left) {
                        if (n < b_left + 3) st[0] = black;
                        else {
                            st[0].val += prof_prev[T.tron(n - 2)] + sigE;
                            st[0].dir = h_is_diag(hq.dir) ? DIAG : NEWD;
                        }
                        // query residue against a gap: whole codon (open / extend) or one / two
                        // nucleotides of the genome skipped with it (frame shifts)
                        const HCell u1 = hx_ld(Hb + r + 1), u2 = hx_ld(Hb + r + 2), u3 = hx_ld(Hb + r + 3);
                        const HCell fu = hx_ld(Fb + r + 3);
                        const int ext = fu.val + P.gep;
                        int x = u1.val + (h_is_vert(u1.dir) ? P.gape1 : P.gw1);
                        if (x > ext) st[2] = HCell{x, u1.ptr, SLA2, 0}; else st[2].val = ext;
                        x = u2.val + (h_is_vert(u2.dir) ? P.gape2 : P.gw2);
                        if (x > st[2].val) st[2] = HCell{x, u2.ptr, SLA1, 0};
                        x = u3.val + P.gw3;
                        if (x >= st[2].val) st[2] = HCell{x, u3.ptr, VERT, 0};
                        else if (ext >= st[2].val) st[2] = HCell{ext, fu.ptr, VERT, 0};
                        if (st[2].val > st[mx].val) mx = 2;
                        if (dagp) {
                            const HCell f2u = hx_ld(F2b + r + 3);
                            x = u3.val + P.gw3l;
                            const int ext2 = f2u.val + P.lgep;
                            if (x >= ext2) st[4] = HCell{x, u3.ptr, VERL, 0};
                            else { st[4] = f2u; st[4].val = ext2; }
                            if (st[4].val > st[mx].val) mx = 4;
                        }
                    }
                    // genome against a gap: a codon (from three columns back, open or extend) or a
                    // frame shift of two / one nucleotides
                    if (n > n0 + 2) {
                        const HCell l3 = hx_ld(Hb + r - 3);
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
                        const HCell l2 = hx_ld(Hb + r - 2);
                        const int x = l2.val + P.gw2;
                        if (x > st[1].val) { st[1] = l2; st[1].val = x; st[1].dir = (st[1].dir & SPIN) + HOR2; }
                    }
                    {
                        const HCell l1 = hx_ld(Hb + r - 1);
                        const int x = l1.val + P.gw1;
                        if (x > st[1].val) { st[1] = l1; st[1].val = x; st[1].dir = (st[1].dir & SPIN) + HOR1; }
                    }
                    if (st[1].val > st[mx].val) mx = 1;

                    // acceptor: the stored donors of the matching splice phase(s)
                    const int phs3 = T.sgf(n, F_PHS3);
                    if (P.spj && phs3 > -2) {
                        for (int phs = phs3 == 2 ? -1 : phs3; ; phs = 1) {
                            const int nb = n - phs;
                            const HxList& L = don[phs + 1];
                            int tj[5], tp[5];
                            unsigned hit = 0;
                            for (int l = 0; l < L.n; ++l) {
                                const int k = L.st[l];
                                if ((phs == 1 && k == 2) || nb - L.jnc[l] < P.minl) continue;
                                int x = L.val[l] + sigB[phs + 1] + hx_spjscr(P, T, L.jnc[l], nb);
                                if (k == 0 && phs) {
                                    // the codon the intron splits is scored with its true translation
                                    const unsigned char* cs = hx_spjseq(P, T, L.jnc[l], nb);
                                    if (phs == 1) x += prof_prev[cs[0]];
                                    else x += prof_next[cs[1]] - prof_next[T.tron(n + 1)] - T.sgf(n + 1, F_SIGE);
                                }
                                if (x > st[k].val) { st[k].val = x; tj[k] = L.jnc[l]; tp[k] = L.ptr[l]; hit |= 1u << k; }
                            }
                            for (int k = 0; k < nod; ++k) {
                                if (!(hit >> k & 1u)) continue;
                                st[k].ptr = A.add(m, n, A.add(m, tj[k] + phs, tp[k]));
                                st[k].dir = c_nod2dir[k] | SPIN;
                                if (st[k].val > st[mx].val) mx = k;
                            }
                            if (phs3 - phs != 3) break;         // AGAG: both phases
                        }
                    }

                    // best state
                    const int y = st[0].val;
                    const int mxdir = st[mx].dir;
                    if (mx != 0) st[0] = st[mx];
                    else if (Local && y > hq.val) {
                        if (LocalL && hq.dir == 0 && !(st[0].dir & SPIN)) st[0].ptr = A.add(m - 1, n - 3, 0);
                        else if (LocalR && y > best_val) { best_val = y; best_p = st[0].ptr; best_m = m; best_n = n; }
                    }
                    int hd_dir = mxdir;
                    if (LocalL && st[0].val <= 0) { st[0].val = 0; st[0].dir = 0; if (mx == 0) hd_dir = 0; }
                    else if (st[0].dir == NEWD) st[0].ptr = A.add(m - 1, n - 3, st[0].ptr);
                    const int mxval = mx == 0 ? st[0].val : st[mx].val;

                    // donor: per splice phase the states that may start an intron here
                    const int phs5 = T.sgf(n, F_PHS5);
                    if (P.spj && phs5 > -2) {
                        for (int phs = phs5 == 2 ? -1 : phs5; ; phs = 1) {
                            const int nb = n - phs;
                            const int sigJ = T.sgf(nb, F_SIG5);
                            const int hd = c_dir2nod[hd_dir & 15];
                            for (int k = (hd == 0 || phs == 1) ? 0 : 1; k < nod; ++k) {
                                const bool cross = phs == 1 && k == 0;  // phase +1 leaves from the cell above-left
                                const HCell from = cross ? hq : st[k];
                                if (!from.dir || (from.dir & SPIN)) continue;
                                if (!cross && k != hd && hd >= 0) {
                                    int z = mxval;
                                    if (hd == 0 || (k - hd) % 2) z += k / 2 == 0 ? 0 : (k / 2 == 1 ? P.gop : P.lgop);
                                    if (from.val <= z) continue;
                                }
                                don[phs + 1].insert(from.val + sigJ, from.ptr, k, nb);
                            }
                            if (phs5 - phs != 3) break;         // GTGT: both phases
                        }
                    }
                    hx_st(Hb + r, st[0]);
                    hx_st(Fb + r, st[2]);
                    if (dagp) hx_st(F2b + r, st[4]);
                    e1[q] = st[1]; e2[q] = st[3];
                    if (++q == 3) q = 0;
                }
                sync_problem();
            }
            __threadfence_block();
            sync_problem();
        }

        // ---- end point
        int bv = best_val, bm = best_m, bn = best_n, bp = best_p;
        if (LocalR) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const int ov = __shfl_xor_sync(FULL, bv, o), om = __shfl_xor_sync(FULL, bm, o);
                const int on = __shfl_xor_sync(FULL, bn, o), op = __shfl_xor_sync(FULL, bp, o);
                if (ov > bv || (ov == bv && ov > HNG_NEVSEL && (om < bm || (om == bm && on < bn)))) { bv = ov; bm = om; bn = on; bp = op; }
            }
            if (NW > 1) {
                // across the warps of the problem, same order
                const int w = threadIdx.x >> 5;
                if ((threadIdx.x & 31) == 0) { s_best[w][0] = bv; s_best[w][1] = bm; s_best[w][2] = bn; s_best[w][3] = bp; }
                __syncthreads();
                if (threadIdx.x == 0)
                    for (int k = 1; k < NW; ++k) {
                        const int ov = s_best[k][0], om = s_best[k][1], on = s_best[k][2];
                        if (ov > bv || (ov == bv && ov > HNG_NEVSEL && (om < bm || (om == bm && on < bn)))) { bv = ov; bm = om; bn = on; bp = s_best[k][3]; }
                    }
            }
        }
        int ptr = 0, val = HNG_NEVSEL;
        if (lane == 0) {
            if (!LocalR || bm == a_right) {
                // lastH_ng (src/fwd2h1.cc:224-292) on the last row of the band, serial
                const int m3 = 3 * a_right;
                const int rw0 = max(lw, b_left - m3);
                const int r9 = b_right - m3;
                int mxr = r9, mxrow = 0;                    // best cell: diagonal, band row (0 H, 1 F)
                (void) rw0;
                auto cell = [&](int r) { return hx_ld(Hb + r); };
                if (a_exgr) {
                    // free end on the genome: per reading frame, extend the frame's trailing gap
                    // codon-wise or stop at a termination codon, whichever scores more
                    int glen[3] = {0, 0, 0};
                    int ph = 0;
                    for (int r = rw0; r <= r9; ++r, ph = ph == 2 ? 0 : ph + 1) {
                        const int n = r + m3;
                        HCell h = cell(r);
                        glen[ph] += 3;
                        int c0 = h.val, c1 = HNG_NEVSEL, c2 = HNG_NEVSEL;
                        if (r - rw0 >= 3) {
                            const HCell h3 = cell(r - 3);
                            if (h3.dir != DEAD) {
                                c1 = h3.val + T.sgf(n - 2, F_SIGE);
                                if (!(a_exgr & 2)) c1 += gap_ext3(P, glen[ph]);
                                if (!(a_exgr & 1) && glen[ph] == 3) c1 += P.gop;
                                if (T.sgf(n - 2, F_SIGT) > 0 && !(h.dir & SPIN)) c2 = h3.val + T.sgf(n - 2, F_SIGT);
                            }
                        }
                        const int s5 = (Local && T.sgf(n, F_SIG5) > 0) ? T.sgf(n, F_SIG5) : 0;
                        c0 += s5; c1 += s5;
                        // (the reference compares through a pointer to the best cell so far: a cell
                        // never beats itself, also after it has just been rewritten)
                        const bool self = mxr == r;
                        const int mxv = cell(mxr).val;
                        int k = 0;
                        if (c1 > c0) k = 1;
                        if (c2 > (k ? c1 : c0)) k = 2;
                        if (k == 0) { if (!h_is_hori(h.dir)) glen[ph] = 0; }
                        else if (k == 1) { h = cell(r - 3); h.dir = HORI; h.val = c1 - s5; hx_st(Hb + r, h); }
                        else {
                            h = cell(r - 3);
                            h.dir = DEAD; h.val = c2;
                            if (!self && h.val > mxv) h.ptr = A.add(a_right, n - 3, h.ptr);
                            hx_st(Hb + r, h);
                        }
                        if (!self && h.val > mxv) mxr = r;
                    }
                } else {
                    const HCell h3 = cell(r9 - 3);
                    const int y = h3.val + T.sgf(b_right - 2, F_SIGT);
                    if (y > cell(r9).val) { HCell h = h3; h.val = y; h.dir = HORI; hx_st(Hb + r9, h); }
                }
                bool done = false;
                if (b_exgr == 1) {
                    // free end on the query: trailing residues unpaired, per reading frame
                    int g[3] = {HNG_NEVSEL, HNG_NEVSEL, HNG_NEVSEL};
                    int ph = 0;
                    for (int r = min(up, b_right - 3 * a_left) - 3; r >= r9; --r) {
                        int x = cell(r + 3).val;
                        if (!(b_exgr & 1)) x += P.gop;
                        if (x > g[ph]) g[ph] = x;
                        if (!(b_exgr & 2)) g[ph] += P.gep;
                        const int mxv = cell(mxr).val;
                        if (cell(r).val > g[ph]) g[ph] = HNG_NEVSEL;
                        else if (g[ph] > mxv) { mxr = r; HCell h = cell(r); h.val = g[ph]; hx_st(Hb + r, h); }
                        if (++ph == 3) ph = 0;
                    }
                } else if (b_exgr == 2) {
                    mxr = r9; mxrow = 1;
                    HCell f = hx_ld(Fb + r9);
                    f.ptr = A.add(a_right, b_right, f.ptr);
                    hx_st(Fb + r9, f);
                    done = true;
                }
                HCell best = hx_ld((mxrow ? Fb : Hb) + mxr);
                if (!done) {
                    int pp = mxr - r9;
                    int m9 = a_right, n9 = b_right;
                    if (pp > 0) { m9 -= (pp + 2) / 3; if (pp %= 3) n9 -= 3 - pp; }
                    else if (pp < 0) n9 += pp;
                    best.ptr = A.add(m9, n9, best.ptr);
                }
                val = best.val;
                ptr = best.ptr;
            } else {
                ptr = A.add(bm, bn, bp);
                val = bv;
            }
        }
        const bool overflow = NW == 1 ? __any_sync(FULL, A.overflow) != 0 : __syncthreads_or(A.overflow) != 0;
        __threadfence_block();
        sync_problem();

        if (lane == 0) {
            // Vmf::traceback + the start-point adjustment of trcbkalignH_ng (src/fwd2h1.cc:2021-2037)
            int2* skl = sklpool + t.skl_off;
            int cnt = 0;
            if (!overflow && ptr) {
                int m_last = 0, n_last = 0;
                for (int p = ptr; p; ) {
                    const int* rr = A.rec + 3 * (long long) p;
                    m_last = rr[0]; n_last = rr[1];
                    if (cnt < t.skl_cap) skl[cnt] = make_int2(m_last, n_last);
                    ++cnt;
                    p = rr[2];
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
            res.status = overflow ? 5 : (cnt > t.skl_cap ? 1 : 0);
            res.n_skl = cnt; res.pad = 0;
            results[ti] = res;
        }
        sync_problem();
    }
}

}   // namespace gspaln
