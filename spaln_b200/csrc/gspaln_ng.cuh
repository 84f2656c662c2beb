// gspaln_ng.cuh -- exact intron-length DP (the reference's scalar formulation) as a warp kernel.
//
// Semantics: bit-identical to Aln2s1::trcbkalignS_ng on its scalar branch (src/fwd2s1.cc:1667-1710:
// forwardS_ng 217-444 with initS_ng / lastS_ng 141-215, the Vmf path records src/vmf.cc:66-140 and
// the start-point adjustment) and to Aln2s1::scorealoneS_ng (1112-1336); intron score
// SpJunc::spjscr (src/codepot.cc:74-77) = IntronPenalty::Penalty(length) + pair-corrected 3'
// signal (Exinon::sig53(IE53), src/codepot.cc:410-415).  int32 cells, a sorted list of the best
// donors of every query row, path records instead of a trace matrix.
//
// Mapping (NOT the reference's row-by-row loop): one warp per problem, the 32 lanes own 32
// consecutive query rows and sweep them as an anti-diagonal wavefront -- at step s lane l sits on
// column s - l of row m0 + l -- so the 32 cells of a step are independent.  A lane receives the
// (H, F, F2, direction) of the row above from its neighbour by warp shuffle; lane 0 takes them
// from the problem's diagonal-indexed band rows in global memory (exactly the reference's
// hh[0..2][r] arrays, every lane stores its cell there too), staged 32 diagonals at a time through
// shared memory with one coalesced load.  Everything a row carries along its columns -- the two
// horizontal gap states, the post-splice flags, the donor list (5 entries x {value, record, state,
// column}) -- lives in the registers of its lane.  Path records are appended to a per-warp store
// in lane-private chunks (the record numbering differs from the reference's, the linked list does
// not); lane 0 walks the list at the end and writes the corner array.
#pragma once
#include "gspaln_kernels.cuh"

namespace gspaln {

constexpr int NG_NCAND = 4;             // NCAND, src/aln.h:55 (the list holds one more: see NgList)
constexpr int NG_NEWD = 8;              // Newd, src/fwd2s1.cc:48
constexpr int NG_WARPS = 4;             // warps (problems in flight) per CTA
constexpr int NG_THREADS = 32 * NG_WARPS;
constexpr int NG_NEVSEL = INT_MIN / 16 * 7;     // NEVSEL, src/cmn.h:79
constexpr int NG_CHUNK = 32;            // path records a lane reserves at a time

struct NgRvp { int val, ptr; };         // value + path record (the reference's RVP)

// per-warp workspace in global memory: band rows by diagonal (index 0 <-> diagonal lw - 1)
struct NgBand {
    NgRvp* H; NgRvp* F; NgRvp* F2;
    unsigned char* dirs;
    int* rec;                           // path records {m, n, previous} x rec_cap
    int rec_cap;
};

// lane-private slice of the warp's record store
struct NgAlloc {
    int cur = 0, end = 0;
    bool overflow = false;
};

__device__ __forceinline__ int ng_add(const NgBand& W, NgAlloc& A, int* warp_next, int m, int n, int prev)
{
    if (A.overflow) return 0;
    if (A.cur == A.end) {
        A.cur = atomicAdd(warp_next, NG_CHUNK);
        A.end = A.cur + NG_CHUNK;
    }
    if (A.end > W.rec_cap) { A.overflow = true; A.cur = A.end = 0; return 0; }
    int* r = W.rec + 3 * (long long) A.cur;
    r[0] = m; r[1] = n; r[2] = prev;
    return A.cur++;
}

// tabs: [0, 544) Exinon::sig53tab, [544, 544 + n_pen) IntronPenalty::Penalty(length)
__device__ __forceinline__ int ng_spjscr(const short* __restrict__ tabs, int n_pen, int d5, int len,
                                         const ColInfo& c3)
{
    const int pen = tabs[544 + min(len, n_pen - 1)];
    const int d3 = c3.pad[0] >> 4;
    const short sig = (short) (c3.sig3 - tabs[16 + d3] + tabs[32 + 16 * d5 + d3]);
    return pen + sig;
}

// The donor list of one row.  The reference keeps NCAND + 1 slots behind an index permutation
// (src/fwd2s1.cc:395-404); what that code does is: the list is sorted by value, best first; a new
// donor enters behind the entries that are at least as good (STRICT: better ones only); the entry
// it pushes out of the best NCAND survives in slot NCAND until the next insertion attempt, which
// drops it whether or not the newcomer gets in.
struct NgList {
    int val[NG_NCAND + 1], ptr[NG_NCAND + 1], jnc[NG_NCAND + 1];
    int inf[NG_NCAND + 1];              // gap state (0..4) | 5' dinucleotide code << 4
    int n;
    __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int l = 0; l <= NG_NCAND; ++l) { val[l] = NG_NEVSEL; ptr[l] = 0; jnc[l] = 0; inf[l] = 0; }
        n = 0;
    }
    template <bool TIES_AHEAD>          // scorealoneS_ng lets an equal newcomer pass (src/fwd2s1.cc:1311)
    __device__ __forceinline__ void insert(int x, int p, int info, int j)
    {
        if (n > NG_NCAND) n = NG_NCAND;
        int pos = 0;
#pragma unroll
        for (int l = 0; l < NG_NCAND; ++l)
            if (l < n && (TIES_AHEAD ? val[l] > x : val[l] >= x)) ++pos;
        if (pos >= NG_NCAND) return;
#pragma unroll
        for (int l = NG_NCAND; l > 0; --l)
            if (l > pos) { val[l] = val[l - 1]; ptr[l] = ptr[l - 1]; jnc[l] = jnc[l - 1]; inf[l] = inf[l - 1]; }
#pragma unroll
        for (int l = 0; l < NG_NCAND; ++l)
            if (l == pos) { val[l] = x; ptr[l] = p; jnc[l] = j; inf[l] = info; }
        ++n;
    }
};

// band rows are written by one lane and read by another within a step or two: read them through
// L2 (ld.global.cg) so that a stale L1 line can never be seen
__device__ __forceinline__ NgRvp ng_ld(const NgRvp* p)
{
    const int2 v = __ldcg(reinterpret_cast<const int2*>(p));
    return NgRvp{v.x, v.y};
}
__device__ __forceinline__ int ng_ldd(const unsigned char* p) { return (int) __ldcg(p); }

// first / last column (exclusive / inclusive) of query row m: max(m - 1 + lw, b_left) < n <= min(m + up, b_right)
__device__ __forceinline__ int ng_row_lo(const DevTask& t, int m) { return max(m - 1 + t.lw, t.b_left); }
__device__ __forceinline__ int ng_row_hi(const DevTask& t, int m) { return min(m + t.up, t.b_right); }

// ---------------------------------------------------------------------------
// SCORE = false: forwardS_ng + path records + walk (task kind GSPALN_FORWARD_NG)
// SCORE = true : scorealoneS_ng (GSPALN_SCOREALONE_NG); its tie rules differ: strict comparisons
//                for the gap states and acceptors, ties accepted in the donor list
// ---------------------------------------------------------------------------
// NW = warps per problem.  NW == 1: a CTA runs NG_WARPS independent problems, one per warp.  NW > 1
// (queries of NG_WIDE_ROWS rows and more): the wavefront is 32 NW rows tall, one problem per CTA,
// barrier per step = the CTA's; the first lane of every warp but the first reads the row above from
// the band rows (its neighbour sits in another warp), written one step earlier.
constexpr int NG_WIDE = 8;
constexpr int NG_WIDE_ROWS = 128;

template <bool SCORE, int NW>
__global__ void __launch_bounds__(NW == 1 ? NG_THREADS : 32 * NW)
dp_xild_kernel(const DevParams* __restrict__ gP, const short* __restrict__ tabs, int n_pen,
               const DevTask* __restrict__ tasks, const int* __restrict__ order, int ntasks, int* ticket,
               const unsigned char* __restrict__ apool, const ColInfo* __restrict__ cpool,
               unsigned char* workpool, long long work_slab, long long width_max, int rec_cap,
               int2* sklpool, DevResult* results, const int* ready)
{
    __shared__ DevParams sP;
    constexpr int NT = 32 * NW;                         // lanes (= rows of a pass) per problem
    constexpr int NWARP = NW == 1 ? NG_WARPS : NW;      // warps of the CTA
    __shared__ NgRvp stageH[NWARP][32], stageF[NWARP][32], stageF2[NWARP][32];
    __shared__ int stageD[NWARP][32];
    __shared__ int warp_next[NWARP];
    __shared__ int s_tk, s_best[NWARP][4];
    {
        const int* src = reinterpret_cast<const int*>(gP);
        int* dst = reinterpret_cast<int*>(&sP);
        for (int i = threadIdx.x; i < (int) (sizeof(DevParams) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const DevParams& P = sP;
    const unsigned FULL = 0xffffffffu;
    const int wl = threadIdx.x & 31, wid = threadIdx.x >> 5;    // lane of the warp (shuffles), warp of the CTA
    const int lane = threadIdx.x % NT;                  // lane of the problem's wavefront
    const int slot = NW == 1 ? wid : 0;                 // problem slot of the CTA
    auto sync_problem = [] { if (NW == 1) __syncwarp(); else __syncthreads(); };
    const bool dagp = P.noll == 3, spj = P.spj != 0;
    const int nod = 2 * P.noll - 1;
    constexpr int WANT = SCORE ? 4 : 3;

    NgBand W;
    {
        unsigned char* base = workpool + ((long long) blockIdx.x * (NW == 1 ? NG_WARPS : 1) + slot) * work_slab;
        W.H = reinterpret_cast<NgRvp*>(base);
        W.F = W.H + width_max;
        W.F2 = W.F + width_max;
        W.dirs = reinterpret_cast<unsigned char*>(W.F2 + width_max);
        W.rec = reinterpret_cast<int*>(W.dirs + ((width_max + 15) / 16) * 16);
        W.rec_cap = rec_cap;
    }

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
        if (t.kind != WANT || ((t.flags & 32) != 0) != (NW > 1)) continue;  // another kernel / class runs it
        bool arrived = wait_inputs(ready, tk);
        if (NW > 1) arrived = __syncthreads_and(arrived) != 0;
        if (!arrived) {
            if (lane == 0) { DevResult rr; rr.score = 0; rr.status = 4; rr.n_skl = 0; rr.pad = 0; results[ti] = rr; }
            continue;
        }
        const unsigned char* aseq = apool + t.a_off;        // aseq[i] pairs query row a_left + i + 1
        // Cip_score of the rows (src/gsinfo.h:127-139), word i = row a_left + i; absent: no bonus
        const int* cip = t.pad1 ? reinterpret_cast<const int*>(aseq + t.pad1) : nullptr;
        const ColInfo* cols = cpool + t.col_off;            // cols[j] is column b_left + j
        const int width = t.up - t.lw + 3;
        const bool a_exgl = t.flags & 1, a_exgr = t.flags & 2, b_exgl = t.flags & 4, b_exgr = t.flags & 8;
        const bool LocalL = P.local && a_exgl && b_exgl, LocalR = P.local && a_exgr && b_exgr;
        const int a_left = t.a_left, a_right = t.a_right, b_left = t.b_left, b_right = t.b_right;
        const int lw = t.lw, up = t.up;
        // band rows by diagonal r = n - m in [lw - 1, up + 1]
        NgRvp* H = W.H - (lw - 1);
        NgRvp* F = W.F - (lw - 1);
        NgRvp* F2 = W.F2 - (lw - 1);
        unsigned char* dirs = W.dirs - (lw - 1);
        NgAlloc A;
        if (lane == 0) warp_next[slot] = NG_CHUNK;          // record 0 is never a path node
        for (int i = lane; i < width; i += NT) {
            W.H[i] = NgRvp{NG_NEVSEL, 0}; W.F[i] = NgRvp{NG_NEVSEL, 0}; W.F2[i] = NgRvp{NG_NEVSEL, 0};
            W.dirs[i] = 0;
        }
        sync_problem();

        // ---- first row and first column (initS_ng src/fwd2s1.cc:141-184, sinitS_ng 1112-1140)
        {
            const int r0 = b_left - a_left;
            int p0 = 0;
            if (!SCORE && lane == 0) p0 = ng_add(W, A, &warp_next[slot], a_left, b_left, 0);
            p0 = __shfl_sync(FULL, p0, 0);
            if (lane == 0) H[r0] = NgRvp{0, p0};
            if (a_exgl) {
                const int rr = min(up, b_right - a_left);
                for (int r = r0 + 1 + lane; r <= rr; r += NT) { H[r] = NgRvp{0, 0}; dirs[r] = 1; }
            }
            const int rr = max(b_left - a_right, lw);
            if (b_exgl) {
                for (int r = rr + lane; r < r0; r += NT) { H[r] = NgRvp{0, 0}; dirs[r] = 2; }
            } else if (lane == 0) {
                // a leading gap in the genome: opened once, extended per residue (long-gap slope
                // beyond codonk1); the score-only kernel seeds the vertical state as well
                int v = 0, f = NG_NEVSEL;
                for (int i = 1, r = r0 - 1; r >= rr; --r, ++i) {
                    if (i == 1) { v += P.gappen1; f = v; }
                    else { v += i > P.codonk1 ? P.lgep : P.gep; f += P.gep; }
                    H[r] = NgRvp{v, p0};
                    dirs[r] = 2;
                    if (SCORE) F[r] = NgRvp{f, 0};
                }
            }
        }
        __threadfence_block();
        sync_problem();

        int best_val = NG_NEVSEL, best_m = a_left, best_n = b_left, best_p = 0;     // LocalR (lane-local)
        const int m_first = a_exgl ? a_left + 1 : a_left;
        for (int m0 = m_first; m0 <= a_right; m0 += NT) {
            const int m = m0 + lane;
            const bool row = m <= a_right;
            const bool first = m == a_left;                 // global start row: horizontal moves only
            const bool internal = spj && (!a_exgr || m < a_right);
            const int lo = ng_row_lo(t, m), hi = ng_row_hi(t, m);
            const int lo_up = ng_row_lo(t, m - 1), hi_up = ng_row_hi(t, m - 1);     // row above
            const bool has_up = lane > 0;                   // the row above belongs to this pass
            const int arow = (first || !row) ? ZROW : (int) aseq[m - 1 - a_left];
            const int sigB = (cip && row) ? cip[m - a_left] : 0;    // bonus of an intron conserved at this row
            const int last_lane = min(NT - 1, a_right - m0);
            const int s_begin = ng_row_lo(t, m0) + 1;
            const int s_end = ng_row_hi(t, m0 + last_lane) + last_lane;

            NgRvp hleft{NG_NEVSEL, 0}, e1{NG_NEVSEL, 0}, e2{NG_NEVSEL, 0};
            NgRvp oH{NG_NEVSEL, 0}, oF{NG_NEVSEL, 0}, oF2{NG_NEVSEL, 0};           // this lane's last cell
            int oD = 0;
            NgRvp dH{NG_NEVSEL, 0};                         // the cell above-left (received one step ago)
            int dD = 0;
            int psp = 0;
            NgList L;
            L.clear();

            for (int s = s_begin; s <= s_end; ++s) {
                const int k = s - s_begin;
                // lane 0's row above: 32 diagonals of the band rows per coalesced load
                if ((k & 31) == 0 && (NW == 1 || wid == 0)) {
                    __syncwarp();
                    const int r = (s - m0 + 1) + lane;      // diagonal lane 0 reads `lane` steps from now
                    const bool in = r >= lw - 1 && r <= up + 1;
                    stageH[wid][lane] = in ? ng_ld(H + r) : NgRvp{NG_NEVSEL, 0};
                    stageF[wid][lane] = in ? ng_ld(F + r) : NgRvp{NG_NEVSEL, 0};
                    if (dagp) stageF2[wid][lane] = in ? ng_ld(F2 + r) : NgRvp{NG_NEVSEL, 0};
                    stageD[wid][lane] = in ? ng_ldd(dirs + r) : 0;
                    __syncwarp();
                }
                // the cell above: neighbour's result of the previous step
                NgRvp uH, uF, uF2;
                int uD;
                uH.val = __shfl_up_sync(FULL, oH.val, 1); uH.ptr = __shfl_up_sync(FULL, oH.ptr, 1);
                uF.val = __shfl_up_sync(FULL, oF.val, 1); uF.ptr = __shfl_up_sync(FULL, oF.ptr, 1);
                uF2.val = dagp ? __shfl_up_sync(FULL, oF2.val, 1) : NG_NEVSEL;
                uF2.ptr = dagp ? __shfl_up_sync(FULL, oF2.ptr, 1) : 0;
                uD = __shfl_up_sync(FULL, oD, 1);
                const int n = s - lane;
                const int r = n - m;
                if (lane == 0) {
                    uH = stageH[wid][k & 31]; uF = stageF[wid][k & 31];
                    if (dagp) uF2 = stageF2[wid][k & 31];
                    uD = stageD[wid][k & 31];
                } else if (wl == 0 || !(n > lo_up && n <= hi_up)) {
                    // the neighbour sits in another warp (its cell went to the band rows one step ago), or
                    // the row above never evaluated column n: what the band rows hold there
                    const bool in = r + 1 <= up + 1 && r + 1 >= lw - 1;
                    uH = in ? ng_ld(H + r + 1) : NgRvp{NG_NEVSEL, 0};
                    uF = in ? ng_ld(F + r + 1) : NgRvp{NG_NEVSEL, 0};
                    if (dagp) uF2 = in ? ng_ld(F2 + r + 1) : NgRvp{NG_NEVSEL, 0};
                    uD = in ? ng_ldd(dirs + r + 1) : 0;
                }
                const bool act = row && n > lo && n <= hi;
                if (act) {
                    const ColInfo col = cols[n - b_left];
                    if (n == lo + 1) {
                        // row start: the band entry left of the first cell, and the cell above-left
                        // if the row above did not evaluate it
                        hleft = ng_ld(H + (lo - m));
                        if (!(has_up && lo > lo_up && lo <= hi_up)) { dH = ng_ld(H + r); dD = ng_ldd(dirs + r); }
                    }
                    // cell states in the reference's order: 0 H, 1 E, 2 F, 3 E2, 4 F2
                    NgRvp h = dH;
                    const int diag = h.val;
                    int dir = dD;
                    NgRvp f{NG_NEVSEL, 0}, f2{NG_NEVSEL, 0};
                    int mx = 0, mxv;
                    if (!first) {
                        h.val += P.mtxT[(int) col.code * MTX_LD + arow];
                        dir = (dir % NG_NEWD) ? NG_NEWD : 0;
                        int x = uH.val + P.gop;
                        f = (x >= uF.val) ? NgRvp{x, uH.ptr} : uF;
                        f.val += P.gep;
                        mxv = h.val;
                        if (f.val > mxv) { mx = 2; mxv = f.val; }
                        if (dagp) {
                            x = uH.val + P.lgop;
                            f2 = (x >= uF2.val) ? NgRvp{x, uH.ptr} : uF2;
                            f2.val += P.lgep;
                            if (f2.val > mxv) { mx = 4; mxv = f2.val; }
                        }
                    } else {
                        f = ng_ld(F + r); if (dagp) f2 = ng_ld(F2 + r);     // untouched band entries
                        mxv = h.val;
                    }
                    {
                        int x = hleft.val + P.gop;
                        const int prev_psp = psp;
                        if (SCORE ? x > e1.val : x >= e1.val) { e1 = NgRvp{x, hleft.ptr}; psp = psp ? 1 : 0; }
                        else psp &= 1;
                        e1.val += P.gep;
                        if (SCORE ? e1.val > mxv : e1.val >= mxv) { mx = 1; mxv = e1.val; }
                        if (dagp) {
                            x = hleft.val + P.lgop;
                            if (SCORE ? x > e2.val : x >= e2.val) { e2 = NgRvp{x, hleft.ptr}; if (prev_psp) psp |= 2; }
                            else psp |= prev_psp & 2;
                            e2.val += P.lgep;
                            if (SCORE ? e2.val > mxv : e2.val >= mxv) { mx = 3; mxv = e2.val; }
                        }
                    }
                    const int cano5 = col.pad[1] & 15, cano3 = col.pad[1] >> 4;
                    // acceptor: every stored donor of this row, per gap state (the last entry that is
                    // at least as good as the state wins)
                    if ((SCORE || internal) && cano3 && L.n > 0) {
                        int tj[5], tp[5];
                        unsigned hit = 0;
#pragma unroll
                        for (int l = 0; l <= NG_NCAND; ++l) {
                            if (l >= L.n || n - L.jnc[l] < P.llmt) continue;
                            const int st = L.inf[l] & 15;
                            const int x = L.val[l] + sigB + ng_spjscr(tabs, n_pen, L.inf[l] >> 4, n - L.jnc[l], col);
#pragma unroll
                            for (int q = 0; q < 5; ++q) {
                                if (q != st) continue;
                                int& sv = q == 0 ? h.val : q == 1 ? e1.val : q == 2 ? f.val : q == 3 ? e2.val : f2.val;
                                if (SCORE ? x > sv : x >= sv) { sv = x; tj[q] = L.jnc[l]; tp[q] = L.ptr[l]; hit |= 1u << q; }
                            }
                        }
                        // the running best may itself have been raised by a donor
                        mxv = mx == 0 ? h.val : mx == 1 ? e1.val : mx == 2 ? f.val : mx == 3 ? e2.val : f2.val;
                        const int psp_bit[5] = {4, 1, 8, 2, 16};    // src/aln.h:56
#pragma unroll
                        for (int q = 0; q < 5; ++q) {
                            if (q >= nod || !(hit >> q & 1u)) continue;
                            NgRvp& sq = q == 0 ? h : q == 1 ? e1 : q == 2 ? f : q == 3 ? e2 : f2;
                            psp |= psp_bit[q];
                            if (!SCORE) {
                                const int inner = ng_add(W, A, &warp_next[slot], m, tj[q], tp[q]);
                                sq.ptr = ng_add(W, A, &warp_next[slot], m, n, inner);
                            }
                            if (SCORE ? sq.val > mxv : sq.val >= mxv) { mx = q; mxv = sq.val; }
                        }
                    }
                    // best state
                    const int hd = mx;
                    const int y = h.val;
                    if (mx != 0) {
                        h = mx == 1 ? e1 : mx == 2 ? f : mx == 3 ? e2 : f2;
                        dir = hd;
                    } else if (SCORE) {
                        if (LocalR && y > best_val) best_val = y;
                    } else if (P.local && h.val > diag) {
                        if (LocalL && diag == 0) h.ptr = ng_add(W, A, &warp_next[slot], m - 1, n - 1, 0);
                        else if (LocalR && h.val > best_val) { best_val = h.val; best_p = h.ptr; best_m = m; best_n = n; }
                    }
                    int mx_now = mxv;
                    if (LocalL && (SCORE ? h.val < 0 : h.val <= 0)) { h.val = 0; dir = 1; if (mx == 0) mx_now = 0; }
                    else if (!SCORE && dir == NG_NEWD && !(psp & 4))
                        h.ptr = ng_add(W, A, &warp_next[slot], m - 1, n - 1, h.ptr);
                    // donor: the best (value + 5' signal) of this row by gap state
                    if ((SCORE || internal) && cano5) {
                        const int sigJ = col.sig5;
                        const int d5 = col.pad[0] & 15;
                        const int psp_bit[5] = {4, 1, 8, 2, 16};
#pragma unroll
                        for (int q = 0; q < 5; ++q) {
                            if (q >= nod || q < (hd == 0 ? 0 : 1) || (psp & psp_bit[q])) continue;
                            const NgRvp from = q == 0 ? h : q == 1 ? e1 : q == 2 ? f : q == 3 ? e2 : f2;
                            if (q != hd) {
                                int z = mx_now;
                                if (hd == 0 || (q - hd) % 2) z += q / 2 == 0 ? 0 : (q / 2 == 1 ? P.gop : P.lgop);
                                if (from.val <= z) continue;
                            }
                            L.template insert<SCORE>(from.val + sigJ, from.ptr, q | (d5 << 4), n);
                        }
                    }
                    // results: to the band rows (what the rows of the next pass and the end-point
                    // search read) and to the neighbour
                    H[r] = h; F[r] = f;
                    if (dagp) F2[r] = f2;
                    dirs[r] = (unsigned char) dir;
                    oH = h; oF = f; oF2 = f2; oD = dir;
                    hleft = h;
                }
                dH = uH; dD = uD;
                sync_problem();
            }
            __threadfence_block();
            sync_problem();
        }

        // ---- end point (lastS_ng src/fwd2s1.cc:186-215, slastS_ng 1142-1161)
        int val = NG_NEVSEL, ptr = 0;
        if (LocalR) {
            // row-major order: the first cell that reached the maximum
            int bv = best_val, bm = best_m, bn = best_n, bp = best_p;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const int ov = __shfl_xor_sync(FULL, bv, o), om = __shfl_xor_sync(FULL, bm, o);
                const int on = __shfl_xor_sync(FULL, bn, o), op = __shfl_xor_sync(FULL, bp, o);
                if (ov > bv || (ov == bv && ov > NG_NEVSEL && (om < bm || (om == bm && on < bn)))) { bv = ov; bm = om; bn = on; bp = op; }
            }
            if (NW > 1) {
                // across the warps of the problem, same order
                if (wl == 0) { s_best[wid][0] = bv; s_best[wid][1] = bm; s_best[wid][2] = bn; s_best[wid][3] = bp; }
                __syncthreads();
                if (threadIdx.x == 0)
                    for (int w = 1; w < NW; ++w) {
                        const int ov = s_best[w][0], om = s_best[w][1], on = s_best[w][2];
                        if (ov > bv || (ov == bv && ov > NG_NEVSEL && (om < bm || (om == bm && on < bn)))) { bv = ov; bm = om; bn = on; bp = s_best[w][3]; }
                    }
            }
            val = bv;
            if (!SCORE && lane == 0) ptr = ng_add(W, A, &warp_next[slot], bm, bn, bp);
        } else if (lane == 0) {
            const int r9 = b_right - a_right;
            int mxr = r9;
            if (SCORE) {
                int mxv = ng_ld(H + r9).val;
                if (b_exgr) for (int r = min(up, b_right - a_left); r > r9; --r) mxv = max(mxv, ng_ld(H + r).val);
                if (a_exgr) for (int r = max(lw, b_left - a_right); r < r9; ++r) mxv = max(mxv, ng_ld(H + r).val);
                val = mxv;
            } else {
                int mxval = ng_ld(H + r9).val;
                if (a_exgr)
                    for (int r = max(lw, b_left - a_right); r <= r9; ++r) {
                        const int v = ng_ld(H + r).val;
                        if (v > mxval) { mxr = r; mxval = v; }
                    }
                if (b_exgr)
                    for (int r = min(up, b_right - a_left); r > r9; --r) {
                        const int v = ng_ld(H + r).val;
                        if (v > mxval) { mxr = r; mxval = v; }
                    }
                const int i = mxr - r9;
                ptr = ng_add(W, A, &warp_next[slot], a_right - max(i, 0), b_right + min(i, 0), ng_ld(H + mxr).ptr);
                val = mxval;
            }
        }
        const bool overflow = NW == 1 ? __any_sync(FULL, A.overflow) != 0 : __syncthreads_or(A.overflow) != 0;
        __threadfence_block();
        sync_problem();

        if (lane == 0) {
            int cnt = 0;
            if (!SCORE && !overflow && ptr) {
                // Vmf::traceback + the start-point adjustment of trcbkalignS_ng (src/fwd2s1.cc:1690-1706)
                int2* skl = sklpool + t.skl_off;
                int m_last = 0, n_last = 0;
                for (int q = ptr; q; ) {
                    const int* rr = W.rec + 3 * (long long) q;
                    m_last = rr[0]; n_last = rr[1];
                    if (cnt < t.skl_cap) skl[cnt] = make_int2(m_last, n_last);
                    ++cnt;
                    q = rr[2];
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
            res.status = overflow ? 5 : (cnt > t.skl_cap ? 1 : 0);
            res.n_skl = cnt; res.pad = 0;
            results[ti] = res;
        }
        sync_problem();
    }
}

}   // namespace gspaln
