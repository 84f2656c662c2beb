// gspaln_kernels.cuh -- sm_100a device code of the spliced-alignment DP engine.
//
// Semantics: bit-identical to SimdAln2s1::forwardS1_wip / scoreonlyS1_wip of
// the reference (src/fwd2s1_wip_simd.h:42-474, src/fwd2s1_simd.cc:163-262,
// src/rhomb_coord.h:65-235) at the AVX2 lane count (strips of 16 query rows).
//
// Mapping (NOT the reference's): one warp owns one DP problem.  A lane owns one
// 16-row strip and walks it column by column with the four per-row state
// words (H, E, best-donor value, intron length) of all 16 rows in registers;
// the vertical dependency runs down the column inside the thread.  The 32
// strips of a pass form a systolic chain: lane t is LAG columns behind lane
// t-1 and receives the (H, F) of the row above through the task's
// diagonal-indexed band buffer in global memory (one packed int16x2 word per
// diagonal, L2 resident, read with ld.global.cg one iteration ahead).  The
// band buffer has exactly the reference's hv[]/fv[] semantics, including
// entries that persist because a strip did not overwrite them.
// Trace codes leave as one 16-byte store per lane per column.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gspaln {

constexpr int NELEM = 16;               // rows per strip == reference nelem (AVX2)
constexpr int NEV = -32768 + 1024;      // nevsel, src/fwd2s1_simd.h:47,202
constexpr int CHECK_SCR = 29490;        // int(0.9 * SHRT_MAX), src/fwd2s1_simd.h:44
constexpr int LAG = 2;                  // systolic lag between neighbouring strips (columns)
constexpr int TRACE_PAD = 46;           // trace columns per strip = width + TRACE_PAD
constexpr int MAXQ = 8;
constexpr int MTX_LD = 32;              // leading dimension of the substitution table in smem
constexpr int ZROW = 31;                // all-zero row: lanes outside the matrix score 0
constexpr int WARPS_PER_CTA = 4;

// TraceBackCode, src/rhomb_coord.h:36-61
enum : unsigned { TB_DIAG = 1, TB_HORI = 2, TB_VERT = 8, TB_ACCR = 14,
                  TB_NHOR = 16, TB_NVER = 32, TB_DONR = 128 };

struct DevParams {
    int gn, ge;                 // (short)(gep + gop), (short) gep
    int ipen, mil, nquant;
    int quant[MAXQ], mean[MAXQ];
    int avmch, local, spj, simdim, gappen1, gop, gep;
    short mtxT[MTX_LD * MTX_LD];    // [genome code][query code], row ZROW == 0
};

struct ColInfo {                // one genome column (8 B)
    short sig5, sig3;           // Exinon::data_n[n]
    unsigned char code;         // *b->at(n - 1)
    unsigned char pad[3];
};

struct DevTask {
    int kind;
    int a_left, a_right, b_left, b_right;
    int lw, up;
    int flags;                  // bit0 a_exgl, bit1 a_exgr, bit2 b_exgl, bit3 b_exgr
    int skl_cap;
    int pad0;
    long long a_off;            // into query-code pool; element 0 == a->at(a_left)
    long long col_off;          // into ColInfo pool; element 0 == column b_left
    long long skl_off;          // into corner pool (int2)
    long long pad1;
};

struct DevResult {
    int score, status, n_skl, pad;
};

__device__ __forceinline__ int sat16(int x) { return max(min(x, 32767), -32768); }
__device__ __forceinline__ int lo16(unsigned w) { return (int) (short) (w & 0xffffu); }
__device__ __forceinline__ int hi16(unsigned w) { return (int) (short) (w >> 16); }
__device__ __forceinline__ unsigned pack16(int lo, int hi)
{
    return ((unsigned) lo & 0xffffu) | ((unsigned) hi << 16);
}

// SimdAln2s1::checkpoint, src/fwd2s1_simd.h:179-182
__device__ __forceinline__ int checkpoint(int avmch, int pv)
{
    return (CHECK_SCR - abs(pv)) / avmch / NELEM * NELEM;
}

struct StripGeom {              // per (task, strip) loop bounds, src/fwd2s1_wip_simd.h:283-289
    int ml, j9, n_start, n_last;    // n_last: last step (inclusive)
};

template <bool TRACE>
__device__ __forceinline__ StripGeom strip_geom(const DevTask& t, int ml)
{
    StripGeom g;
    g.ml = ml;
    g.j9 = min(NELEM, t.a_right - ml);
    g.n_start = max(t.b_left, t.lw + ml);
    const int n9 = min(t.b_right, t.up + (ml + g.j9) + 1) + g.j9;
    g.n_last = TRACE ? n9 : n9 - 1;     // `n <= n9` (forward) vs `n < n9` (score only)
    return g;
}

struct WarpMax { int val, mr, nr; };

// ---------------------------------------------------------------------------
// one pass: strips ml0, ml0+16, ... (nstr <= 32), lane t owns strip t
// ---------------------------------------------------------------------------
template <bool TRACE>
__device__ void run_pass(const DevParams& P, const short* __restrict__ smtx,
                         const DevTask& t, const unsigned char* __restrict__ aseq,
                         const ColInfo* __restrict__ cols, unsigned* band,
                         unsigned char* trace, int ml0, int nstr, bool localL_now,
                         bool localR, int accscr, WarpMax& wmax)
{
    const int lane = threadIdx.x & 31;
    const bool mine = lane < nstr;
    const StripGeom g = strip_geom<TRACE>(t, ml0 + NELEM * lane);
    const int j8 = g.j9 - 1;
    const int span = g.n_last - g.n_start;          // steps - 1
    const bool live = mine && span >= 0;
    const int clo = g.n_start - j8;                 // first column touched by any row
    const int chi = g.n_last;                       // last column touched
    const int width = t.up - t.lw + 3;

    // Systolic schedule: in iteration i lane t works on column
    //     c = c0 + i - LAG * t,
    // i.e. lane t reaches a column LAG iterations after lane t-1 did.  c0 is
    // chosen so that every lane meets its first column at some i >= 0.
    int c0 = live ? clo + LAG * lane : INT_MAX;
    int cend = live ? chi + LAG * lane : INT_MIN;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        c0 = min(c0, __shfl_xor_sync(0xffffffffu, c0, o));
        cend = max(cend, __shfl_xor_sync(0xffffffffu, cend, o));
    }
    if (c0 == INT_MAX) return;                      // nothing to do in this pass
    const int niter = cend - c0 + 1;

    int H[NELEM], E[NELEM], V2[NELEM], IL[NELEM], arow[NELEM];
#pragma unroll
    for (int k = 0; k < NELEM; ++k) {
        H[k] = NEV; E[k] = NEV; V2[k] = NEV; IL[k] = 0;
        arow[k] = (live && k < g.j9) ? aseq[(g.ml - t.a_left) + k] : 0;
    }
    const int gn = P.gn, ge = P.ge, mil = P.mil;
    const int floorL = localL_now ? 0 : INT_MIN;
    int prev_uh = NEV;
    // best local-mode cell of this lane: value, step (n = c + k), row index
    int bval = INT_MIN, bstep = 0, bk = 0;

    unsigned char* tr_base = trace + ((long long) (lane + (ml0 - t.a_left) / NELEM) * (width + TRACE_PAD)) * NELEM;
    const int tr_c0 = t.lw + g.ml - (NELEM - 1);
    const int band_bias = g.ml + t.lw - 1;          // entry index of diagonal d = c - ml: c - band_bias

    // software pipeline registers
    unsigned nxt_band = 0;
    ColInfo nxt_col = ColInfo{0, 0, 0, {0, 0, 0}};
    {
        const int c = c0 - LAG * lane;              // column of iteration 0
        if (live && c >= g.n_start && c <= g.n_last) nxt_band = __ldcg(band + (c - band_bias));
        if (live && c >= t.b_left && c <= t.b_right) nxt_col = cols[c - t.b_left];
    }

    for (int i = 0; i < niter; ++i) {
        const int c = c0 + i - LAG * lane;          // this lane's column in iteration i
        const unsigned cur_band = nxt_band;
        const ColInfo cur_col = nxt_col;
        const bool in_box = live && c >= clo && c <= chi;
        {
            const int cn = c + 1;
            if (live && cn >= g.n_start && cn <= g.n_last) nxt_band = __ldcg(band + (cn - band_bias));
            if (live && cn >= t.b_left && cn <= t.b_right) nxt_col = cols[cn - t.b_left];
        }
        if (in_box) {
            const bool bvalid = c > t.b_left && c <= t.b_right;
            const bool svalid = P.spj && c >= g.n_start && c <= t.b_right;
            const short* prof = smtx + (bvalid ? (int) cur_col.code : ZROW) * MTX_LD;
            const int s3 = svalid ? (int) cur_col.sig3 : 0;
            const int s5 = svalid ? (int) (short) ((int) cur_col.sig5 + P.ipen) : 0;
            // row above the strip: band buffer (hv[r+1], fv[r+1], hv[r])
            const bool top = c >= g.n_start;        // row 0 active (c <= n_last holds: c <= chi)
            int up_h = NEV, up_f = NEV, diag = NEV;
            if (top) {
                up_h = lo16(cur_band);
                up_f = hi16(cur_band);
                diag = (c == g.n_start) ? lo16(__ldcg(band + (c - 1 - band_bias))) : prev_uh;
                prev_uh = up_h;
            }
            unsigned tw[4] = {0u, 0u, 0u, 0u};
            int out_h = NEV, out_f = NEV;
            const int rel0 = c - g.n_start;         // rel0 + k in [0, span] <=> row k active
#pragma unroll
            for (int k = 0; k < NELEM; ++k) {
                const bool act = (k < g.j9) && ((unsigned) (rel0 + k) <= (unsigned) span);
                if (act) {
                    const int left = H[k];
                    unsigned hb = 0;
                    // horizontal: genome residue against a gap
                    int x = sat16(left + gn);
                    int e = sat16(E[k] + ge);
                    if (!(e > x)) { e = x; hb = TB_NHOR; }
                    E[k] = e;
                    // vertical: query residue against a gap
                    int f = sat16(up_f + ge);
                    x = sat16(up_h + gn);
                    if (!(f > x)) { f = x; hb |= TB_NVER; }
                    // diagonal
                    int h = sat16((int) prof[arow[k]] + diag);
                    unsigned pb = TB_DIAG;
                    if (f > h) { h = f; pb = TB_VERT; }
                    if (e > h) { h = e; pb = TB_HORI; }
                    bool acc = false;
                    if (P.spj) {
                        // acceptor: best donor so far + 3' signal + binned length penalty
                        int q = sat16(V2[k] + s3);
                        int pen = P.mean[0];
#pragma unroll
                        for (int j = 1; j < MAXQ; ++j)
                            if (j < P.nquant && IL[k] > P.quant[j - 1]) pen = P.mean[j];
                        q = sat16(q + pen);
                        if (!(IL[k] > mil)) q = NEV;
                        if (q > h) { h = q; pb = TB_ACCR; acc = true; }
                    }
                    if (h < floorL) { h = floorL; hb = 0; }
                    if (P.spj) {
                        // donor
                        int q = sat16(h + s5);
                        if (TRACE && acc) q = NEV;
                        if (q > V2[k]) { V2[k] = q; IL[k] = 0; if (TRACE) hb |= TB_DONR; }
                        IL[k] = min(IL[k] + 1, 32767);
                    }
                    if (localR) {
                        const int step = c + k;
                        if (h > bval || (h == bval && (step < bstep || (step == bstep && k < bk)))) {
                            bval = h; bstep = step; bk = k;
                        }
                    }
                    diag = left;
                    H[k] = h;
                    up_h = h;
                    up_f = f;
                    if (TRACE) tw[k >> 2] |= (hb | pb) << (8 * (k & 3));
                    if (k == j8) { out_h = h; out_f = f; }
                } else {
                    // row not started yet (or finished): the row below sees the
                    // initial lane contents
                    diag = H[k];
                    up_h = H[k];
                    up_f = NEV;
                }
            }
            if (TRACE) {
                *reinterpret_cast<uint4*>(tr_base + (long long) (c - tr_c0) * NELEM) =
                    make_uint4(tw[0], tw[1], tw[2], tw[3]);
            }
            // bottom row of the strip -> band buffer (src/fwd2s1_wip_simd.h:438-442)
            const int rb = c - g.n_start + j8;      // bottom row active?
            if ((unsigned) rb <= (unsigned) span && c > t.b_left) {
                const int r0 = c - (g.ml + g.j9);
                if (r0 >= t.lw && r0 <= t.up)
                    __stcg(band + (r0 - t.lw + 1), pack16(out_h, out_f));
            }
        }
        __syncwarp();
    }

    if (localR) {
        // reference order: strips ascending, then step, then lane (first max)
        int v = (live && bval > INT_MIN) ? bval : INT_MIN;
        int best = v, who = lane;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            int ov = __shfl_xor_sync(0xffffffffu, best, o);
            int ow = __shfl_xor_sync(0xffffffffu, who, o);
            if (ov > best || (ov == best && ow < who)) { best = ov; who = ow; }
        }
        // reference: k1 = lane + 1; mr = ml + k1; nr = n - k1 + 1 with n = step
        const int mr = __shfl_sync(0xffffffffu, g.ml + bk + 1, who);
        const int nr = __shfl_sync(0xffffffffu, bstep - bk, who);
        if (best > INT_MIN && best + accscr > wmax.val) {
            wmax.val = best + accscr;
            wmax.mr = mr;
            wmax.nr = nr;
        }
    }
}

// ---------------------------------------------------------------------------
// trace-code lookup for the walk (cells never evaluated read as STOP = 0)
// ---------------------------------------------------------------------------
struct TraceView {
    const DevTask* t;
    const unsigned char* trace;
    int width;
    __device__ __forceinline__ unsigned code(int cur_m, int cur_n) const
    {
        const DevTask& T = *t;
        if (cur_m == 0)                         // initialize_m0(HORI), src/fwd2s1_wip_simd.h:269
            return (!(T.flags & 1) && cur_n >= 1) ? TB_HORI : 0u;
        const int s = (cur_m - 1) / NELEM, k = (cur_m - 1) % NELEM;
        const StripGeom g = strip_geom<true>(T, T.a_left + s * NELEM);
        const int c = cur_n + T.b_left;
        if (c < g.n_start - k || c > g.n_last - k) return 0u;
        const long long off = ((long long) s * (width + TRACE_PAD) + (c - (T.lw + g.ml - (NELEM - 1)))) * NELEM + k;
        return trace[off];
    }
};

// Anti_rhomb_coord<CHAR>::traceback + go_back (src/rhomb_coord.h:142-235), step = 1
__device__ int walk_trace(const DevTask& t, const unsigned char* trace, int m_abs, int n_abs,
                          int2* skl, int cap, int* status)
{
    TraceView tv{&t, trace, t.up - t.lw + 3};
    int m = m_abs - t.a_left, n = n_abs - t.b_left;
    unsigned code = tv.code(m, n);
    int cnt = 0;
    auto to_left = [&](int s) -> unsigned {
        n -= s;
        if (n < 0) { n = 0; return 0u; }
        return tv.code(m, n);
    };
    auto to_upper = [&](int s) -> unsigned {
        --m; n -= s;
        if (m < 0) { m = 0; n += s; return 0u; }
        if (n < 0) { if (s > 0) m -= n / s; n = 0; return 0u; }
        return tv.code(m, n);
    };
    while (code) {
        if (cnt < cap) skl[cnt] = make_int2(m + t.a_left, n + t.b_left);
        ++cnt;
        const unsigned dir = code & 15u;
        if (dir == TB_DIAG) {
            do { code = to_upper(1); } while (code && (code & 15u) == TB_DIAG);
        } else if (dir == TB_HORI) {
            bool stop = false;
            while (!(code & TB_NHOR)) { code = to_left(1); if (!code) { stop = true; break; } }
            if (!stop) code = to_left(1);
        } else if (dir == TB_VERT) {
            bool stop = false;
            while (!(code & TB_NVER)) { code = to_upper(0); if (!code) { stop = true; break; } }
            if (!stop) code = to_upper(0);
        } else if (dir == TB_ACCR) {
            do { code = to_left(1); } while (code && !(code & TB_DONR));
        } else {
            *status = 2;
            break;
        }
    }
    if (cnt < cap) skl[cnt] = make_int2(m + t.a_left, n + t.b_left);
    ++cnt;
    return cnt;
}

// ---------------------------------------------------------------------------
// persistent kernel: each warp pulls problems from a global ticket counter
// ---------------------------------------------------------------------------
template <bool TRACE>
__global__ void __launch_bounds__(32 * WARPS_PER_CTA)
dp_wip_kernel(const DevParams* __restrict__ gP, const DevTask* __restrict__ tasks,
              const int* __restrict__ order, int ntasks, int* ticket,
              const unsigned char* __restrict__ apool, const ColInfo* __restrict__ cpool,
              unsigned* bandpool, long long band_slab, unsigned char* tracepool,
              long long trace_slab, int2* sklpool, DevResult* results)
{
    __shared__ DevParams sP;
    {
        const int* src = reinterpret_cast<const int*>(gP);
        int* dst = reinterpret_cast<int*>(&sP);
        for (int i = threadIdx.x; i < (int) (sizeof(DevParams) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const DevParams& P = sP;
    const int lane = threadIdx.x & 31;
    // per-warp workspace: band rows and trace matrix are reused by every
    // problem this warp picks up (the walk runs before the next forward pass)
    const long long wslot = (long long) blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    unsigned* band = bandpool + wslot * band_slab;
    unsigned char* trace = TRACE ? tracepool + wslot * trace_slab : nullptr;

    for (;;) {
        int tk = 0;
        if (lane == 0) tk = atomicAdd(ticket, 1);
        tk = __shfl_sync(0xffffffffu, tk, 0);
        if (tk >= ntasks) break;
        const int ti = order[tk];
        const DevTask t = tasks[ti];
        if ((t.kind == 0) != TRACE) continue;       // handled by the other instantiation
        const unsigned char* aseq = apool + t.a_off;
        const ColInfo* cols = cpool + t.col_off;
        const int width = t.up - t.lw + 3;
        const int buf_size = width + 2 * NELEM;
        const bool a_exgl = t.flags & 1, a_exgr = t.flags & 2, b_exgl = t.flags & 4, b_exgr = t.flags & 8;
        const bool LocalL = P.local && a_exgl && b_exgl;
        const bool LocalR = P.local && a_exgr && b_exgr;

        // ---- fhinitS1 (src/fwd2s1_simd.cc:163-184); entry i <-> diagonal lw - 1 + i
        for (int i = lane; i < buf_size; i += 32) band[i] = pack16(NEV, NEV);
        __syncwarp();
        {
            const int rl = t.b_left - t.a_left;
            int rr = t.b_right - t.a_left;
            if (t.up < rr) rr = t.up;
            if (b_exgl)
                for (int r = t.lw + lane; r < rl; r += 32) band[r - t.lw + 1] = pack16(0, NEV);
            if (a_exgl) {
                for (int r = rl + lane; r <= rr; r += 32) band[r - t.lw + 1] = pack16(0, NEV);
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
        }
        __threadfence_block();
        __syncwarp();

        // ---- strips in passes of <= 32, cut at re-basing check points
        int accscr = 0;
        const int md = checkpoint(P.avmch, 0);
        int mc = md + t.a_left;
        WarpMax wmax{NEV, t.a_right, t.b_right};
        int ml0 = t.a_left;
        while (ml0 < t.a_right) {
            int nstr = min(32, (t.a_right - ml0 + NELEM - 1) / NELEM);
            if (mc >= ml0 && mc < ml0 + nstr * NELEM && ((mc - ml0) % NELEM) == 0)
                nstr = (mc - ml0) / NELEM + 1;
            run_pass<TRACE>(P, P.mtxT, t, aseq, cols, band, trace, ml0, nstr,
                            LocalL && !accscr, LocalR, accscr, wmax);
            const int last_ml = ml0 + (nstr - 1) * NELEM;
            if (last_ml == mc) {
                // src/fwd2s1_wip_simd.h:454-465
                const int nmax = t.up - t.lw;
                int cm = lo16(__ldcg(band + 1));
                for (int i = lane; i < nmax; i += 32) cm = max(cm, lo16(__ldcg(band + 1 + i)));
#pragma unroll
                for (int o = 16; o; o >>= 1) cm = max(cm, __shfl_xor_sync(0xffffffffu, cm, o));
                const int d = checkpoint(P.avmch, cm);
                if (d < md / 2) {
                    const int nn = width / NELEM * NELEM;   // saturating part; tail wraps
                    for (int i = lane; i < width; i += 32) {
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
                __syncwarp();
            }
            ml0 += nstr * NELEM;
        }

        // ---- fhlastS1 (src/fwd2s1_simd.cc:241-262)
        if (!LocalR) {
            const int rr = t.b_right - t.a_right;
            int maxr = rr;
            auto argmax = [&](int from, int n) -> int {     // first maximum; `from` if n <= 0
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
            if (a_exgr) {
                const int r = max(t.lw, t.b_left - t.a_right);
                maxr = argmax(r, rr - r);
            }
            if (b_exgr) {
                const int r = min(t.up - 1, t.b_right - t.a_left);
                const int mv = argmax(rr, r - rr);
                if (lo16(__ldcg(band + (mv - t.lw + 1))) > lo16(__ldcg(band + (maxr - t.lw + 1)))) maxr = mv;
            }
            wmax.val = lo16(__ldcg(band + (maxr - t.lw + 1))) + accscr;
            if (maxr > rr) wmax.mr = t.b_right - maxr;
            else wmax.nr = t.a_right + maxr;
        }

        int status = 0, n_skl = 0;
        if (TRACE) {
            __threadfence_block();
            __syncwarp();
            if (lane == 0)
                n_skl = walk_trace(t, trace, wmax.mr, wmax.nr, sklpool + t.skl_off, t.skl_cap, &status);
            if (lane == 0 && n_skl > t.skl_cap && status == 0) status = 1;
        }
        if (lane == 0) {
            DevResult r;
            r.score = wmax.val; r.status = status; r.n_skl = n_skl; r.pad = 0;
            results[ti] = r;
        }
        __syncwarp();
    }
}

}   // namespace gspaln
