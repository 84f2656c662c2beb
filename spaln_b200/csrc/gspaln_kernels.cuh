// gspaln_kernels.cuh -- sm_100a device code of the spliced-alignment DP engine.
//
// Semantics: bit-identical to SimdAln2s1::forwardS1_wip / scoreonlyS1_wip of
// the reference (src/fwd2s1_wip_simd.h:42-474, src/fwd2s1_simd.cc:163-262,
// src/rhomb_coord.h:65-235) at the AVX2 lane count (strips of 16 query rows).
//
// Mapping (NOT the reference's): one warp owns one DP problem.  TPS = 2
// neighbouring threads own one 16-row strip (NR = 8 rows each, lock step,
// joined by two warp shuffles per step) and evaluate it along anti-diagonals:
// at step n strip row k sits on column n - k, so the cell updates of a step
// are independent (no serial chain) and all rows of a strip start and stop together,
// exactly like the lanes of the reference's vector -- including the cells
// that lie outside the matrix.  All per-row state (H of the last two steps,
// F, E, best donor value, intron length) lives in registers.  The 16 strips
// of a pass form a systolic chain: strip s runs 15 + LAG steps behind
// strip s-1 and receives the (H, F) of the row above it through the
// problem's diagonal-indexed band buffer in global memory (one packed
// int16x2 word per diagonal, L2 resident, fetched with ld.global.cg one
// iteration ahead).  The band buffer has exactly the reference's hv[]/fv[]
// semantics, including entries that persist because nobody overwrote them.
// Per-column inputs (substitution-table row, 3'/5' signals) sit in a small
// per-thread shared-memory ring so that row k can read column n - k with an
// immediate offset.  Trace codes leave as one 16-byte store per thread/step.
#pragma once
#include <climits>
#include <cstdint>
#include <cuda_runtime.h>
#include "gspaln_packed.cuh"

namespace gspaln {

constexpr int NELEM = 16;               // rows per strip == reference nelem (AVX2)
#ifndef GSPALN_NR
#define GSPALN_NR 8
#endif
constexpr int NR = GSPALN_NR;           // strip rows owned by one thread
constexpr int TPS = NELEM / NR;         // threads per strip (run in lock step)
constexpr int SPP = 32 / TPS;           // strips per pass of one warp
constexpr int NEV = -32768 + 1024;      // nevsel, src/fwd2s1_simd.h:47,202
constexpr int CHECK_SCR = 29490;        // int(0.9 * SHRT_MAX), src/fwd2s1_simd.h:44
constexpr int PF = 1;                   // band entries are loaded PF steps before their use
constexpr int LAG = PF + 1;             // extra systolic lag (iterations): the entry of column n + PF
                                        // must be final while column n is evaluated.  Measured on B200:
                                        // PF = 3 changes nothing (364 GCUPS on config 2; a warp that
                                        // runs alone does 0.35 GCUPS either way: it is bound by its own
                                        // dependent-issue latency, not by the band loads)
constexpr int TRACE_PAD = 36;           // trace steps per strip <= width + 31; slab stride width + TRACE_PAD
constexpr int MAXQ = 8;
constexpr int MTX_LD = 36;              // row stride (words) of the substitution table: with the four
                                        // common residues remapped to 0..3 their 16 pairs hit 16 distinct banks
constexpr int ZROW = 31;                // all-zero row / column: cells outside the matrix score 0
constexpr int WARPS_PER_CTA = 4;
constexpr int CTA_THREADS = 32 * WARPS_PER_CTA;
constexpr int RING = 32;                // doubled 16-entry ring of per-column inputs
constexpr int PEN_INVALID = -200000;    // added to an int16: always far below the int16 range

// TraceBackCode, src/rhomb_coord.h:36-61
enum : unsigned { TB_DIAG = 1, TB_HORI = 2, TB_HORL = 3, TB_VERT = 8, TB_VERL = 9, TB_ACCR = 14,
                  TB_NHOR = 16, TB_NVER = 32, TB_NHOL = 64, TB_NVEL = 128, TB_DONR = 128 };

struct DevParams {
    int gn, ge;                 // (short)(gep + gop), (short) gep
    int gn2, ge2;               // double affine (Noll == 3): (short)(lgep + lgop), (short) lgep
    int ipen, mil, nquant;
    int quant[MAXQ], mean[MAXQ];
    int avmch, local, spj, simdim, gappen1, gop, gep;
    int pen_cap;                // pen table has pen_cap + 1 entries
    int lgop, lgep, noll, llmt, codonk1;    // raw values for the scalar kernel (gspaln_ng.cuh)
    int pk_ok;                  // the packed int16x2 kernel may run problems of this parameter set
    int pk_pvmax;               // largest positive substitution score (high-side monitor)
    int pk_nidx;                // table index of the residue class N
    int mtxT[32 * MTX_LD];      // [genome index][query index] (remapped codes); row/col ZROW == 0
    unsigned char perm[32];     // residue code -> table index
};

struct ColInfo {                // one genome column (8 B)
    short sig5, sig3;           // Exinon::data_n[n]
    unsigned char code;         // table index of *b->at(n - 1) (DevParams::perm applied on the host)
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

// per-column inputs as the rows consume them
struct __align__(16) RingEntry {
    int prof;                   // byte offset of the substitution-table row of this column's residue
    int s3;                     // acceptor signal
    int s5;                     // donor signal + mean intron penalty
    int pad;
};

__host__ __device__ __forceinline__ int sat16(int x) { return max(min(x, 32767), -32768); }
__host__ __device__ __forceinline__ int satlo(int x) { return max(x, -32768); }     // addend <= 0
__host__ __device__ __forceinline__ int lo16(unsigned w) { return (int) (short) (w & 0xffffu); }
__host__ __device__ __forceinline__ int hi16(unsigned w) { return (int) (short) (w >> 16); }
__host__ __device__ __forceinline__ unsigned pack16(int lo, int hi)
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

struct WarpMax { int val, mr, nr, err; };

// One-shot submits stream the batch in while the persistent kernel runs: `ready` counts the
// problems (in ticket order) whose inputs have arrived in HBM.  Returns false on time-out
// (1 << 36 cycles, ~35 s: the host died or a copy failed -- far beyond what a loaded host needs
// to pack the next chunk) so that the kernel never spins forever; the problem then carries
// status GSPALN_ST_INTERNAL and the drivers skip its post-work.
__device__ __forceinline__ bool wait_inputs(const int* ready, int tk)
{
    if (!ready) return true;
    int ok = 1;
    if ((threadIdx.x & 31) == 0) {
        const volatile int* r = ready;
        const long long t0 = clock64();
        while (*r <= tk) {
            __nanosleep(256);
            if (clock64() - t0 > (1ll << 36)) { ok = 0; break; }
        }
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    __syncwarp();
    return ok != 0;
}

// shared-memory carve-up of one CTA
struct SmemLayout {
    RingEntry* ring;            // [RING][CTA_THREADS]
    const int2* pen;            // [pen_cap + 1] {penalty, lower clamp}
    const int* mtx;             // [32 * MTX_LD]
};

// ---------------------------------------------------------------------------
// one step of one thread: NR independent cell updates (its rows NR-1 .. 0)
//   HN: H of the previous step (read as "left" and, shifted by one row, as
//       "up"); HO: H of two steps ago (read shifted as "diag"); the new H
//       overwrites HO, so callers alternate the two arrays.
//   NJ: minus the step index at which the row's intron-length counter was
//       last reset (counter value == step + NJ; saturation is implied by the
//       clamp to the table size).
// ---------------------------------------------------------------------------
template <int NR, bool TRACE, bool LOCAL, bool SPJ, bool DAGP>
__host__ __device__ __forceinline__ void strip_step(
    int (&HO)[NR], const int (&HN)[NR], int (&F)[NR], int (&E)[NR],
    int (&F2)[NR], int (&E2)[NR], int up_f2, int gn2, int ge2,
    int (&V2)[NR], int (&NJ)[NR], const int (&arow)[NR],
    const char* __restrict__ ring_hi, const char* __restrict__ mtx_bytes,
    const int2* __restrict__ pen_tab, int pen_cap, int step,
    int up_h, int up_f, int up_d, int gn, int ge, int floorL, unsigned (&tw)[(NR + 3) / 4],
    int& best_v, int& best_k)
{
    if (TRACE) {
#pragma unroll
        for (int w = 0; w < (NR + 3) / 4; ++w) tw[w] = 0u;
    }
#pragma unroll
    for (int k = NR - 1; k >= 0; --k) {
        // column n - (row index in the strip): ring slot (n & 15) + 16 - row
        const RingEntry re = *reinterpret_cast<const RingEntry*>(
            ring_hi - k * (CTA_THREADS * (int) sizeof(RingEntry)));
        const int left = HN[k];
        const int uh = k ? HN[k ? k - 1 : 0] : up_h;
        const int uf = k ? F[k ? k - 1 : 0] : up_f;
        const int dg = k ? HO[k ? k - 1 : 0] : up_d;
        unsigned hb = 0;
        // horizontal: genome residue against a gap (gn, ge <= 0: only the low side can saturate)
        int x = satlo(left + gn);
        int e = satlo(E[k] + ge);
        if (!(e > x)) { e = x; hb = TB_NHOR; }
        E[k] = e;
        int e2 = NEV, f2 = NEV;
        if (DAGP) {
            // second (long) gap state, src/fwd2s1_wip_simd.h:99-109, 312-321
            const int x2 = satlo(left + gn2);
            e2 = satlo(E2[k] + ge2);
            if (!(e2 > x2)) { e2 = x2; hb |= TB_NHOL; }
            E2[k] = e2;
            if (!TRACE && e2 > e) { e = e2; E[k] = e; }     // score only: the two states are merged
        }
        // vertical: query residue against a gap
        int f = satlo(uf + ge);
        x = satlo(uh + gn);
        if (!(f > x)) { f = x; hb |= TB_NVER; }
        if (DAGP) {
            const int uf2 = k ? F2[k ? k - 1 : 0] : up_f2;
            f2 = satlo(uf2 + ge2);
            // forwardS1_wip re-uses a register that by then holds the NVER flag (0 | 32): its long
            // vertical gap opens from that value, not from H (wip.h:333,339); score only is sound
            x = TRACE ? satlo((int) (hb & TB_NVER) + gn2) : satlo(uh + gn2);
            if (!(f2 > x)) { f2 = x; hb |= TB_NVEL; }
            F2[k] = f2;
            if (!TRACE && f2 > f) f = f2;
        }
        F[k] = f;
        // diagonal
        const int pv = *reinterpret_cast<const int*>(mtx_bytes + re.prof + arow[k]);
        int h = sat16(pv + dg);
        unsigned pb = TB_DIAG;
        if (f > h) { h = f; pb = TB_VERT; }
        if (TRACE && DAGP && f2 > h) { h = f2; pb = TB_VERL; }
        if (e > h) { h = e; pb = TB_HORI; }
        if (TRACE && DAGP && e2 > h) { h = e2; pb = TB_HORL; }
        bool acc = false;
        if (SPJ) {
            // acceptor: best donor of this row + 3' signal + binned length penalty,
            // only if the intron is longer than the lower limit
            const int q0 = sat16(V2[k] + re.s3);
            const int2 pq = pen_tab[min(step + NJ[k], pen_cap)];
            const int q = min(max(q0 + pq.x, pq.y), 32767);
            if (q > h) { h = q; pb = TB_ACCR; acc = true; }
        }
        if (LOCAL) { if (h < floorL) { h = floorL; hb = 0; } }
        if (SPJ) {
            // donor
            int q = sat16(h + re.s5);
            if (TRACE && acc) q = NEV;          // no empty intron
            const bool don = q > V2[k];
            V2[k] = max(V2[k], q);
            NJ[k] = don ? -step : NJ[k];        // counter := 0, then += 1 at the end of this step
            if (TRACE && don) hb |= TB_DONR;
        }
        if (LOCAL) {
            if (h >= best_v) { best_v = h; best_k = k; }    // descending k: ties end at the lowest row
        }
        HO[k] = h;
        if (TRACE) tw[k >> 2] |= ((hb | pb) & 0xffu) << (8 * (k & 3));
    }
}

// ---------------------------------------------------------------------------
// one segment: strips ml0, ml0+16, ... (nstr of them; a segment ends at a
// re-basing check point or at the last query row).  The warp has SPP strip
// slots (TPS neighbouring threads each, NR rows per thread, lock step; the
// lower thread takes the (H, F) of the row above it from its neighbour by warp
// shuffle).  Slot s runs strips s, s + SPP, s + 2 SPP, ... back to back: a slot
// that finishes a strip starts its next one as soon as the strip above that one
// is 15 + LAG steps ahead, so the systolic chain never drains inside a segment.
// ---------------------------------------------------------------------------
template <int NR, bool TRACE, bool LOCAL, bool SPJ, bool DAGP>
__device__ void run_pass(const DevParams& P, const SmemLayout& sm,
                         const DevTask& t, const unsigned char* __restrict__ aseq,
                         const ColInfo* __restrict__ cols, unsigned* band, int* band2,
                         unsigned char* trace, int ml0, int nstr, bool localL_now,
                         bool localR, int accscr, WarpMax& wmax)
{
    constexpr int TPS = NELEM / NR;             // threads per strip (lock step)
    constexpr int SPP = 32 / TPS;               // strip slots of the warp
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int slot = lane / TPS;                // strip slot of this thread
    const int sub = lane % TPS;                 // which NR-row slice of the strip
    const int row0 = sub * NR;                  // first strip row owned by this thread
    const int pred_lane = ((slot + SPP - 1) % SPP) * TPS;
    const int width = t.up - t.lw + 3;

    // current strip of this slot.  Step j of the strip runs in iteration j + off.
    // Strip s needs, at its step n, the band entry of column n, which strip s-1
    // (a full strip) produces at ITS step n + 15; hence
    //   off_s >= off_{s-1} + (n_start_s - n_start_{s-1}) + 15 + LAG.
    int sidx = slot;                            // strip index inside the segment
    StripGeom g = strip_geom<TRACE>(t, ml0 + NELEM * min(sidx, nstr - 1));
    int nsteps = g.n_last - g.n_start + 1;
    int off = 0;
    // state: 0 = waiting for the strip above to be scheduled, 1 = scheduled / running,
    //        2 = no strip left, 3 = scheduled but empty (band outside the matrix)
    int state = sidx < nstr ? (sidx == 0 ? (nsteps > 0 ? 1 : 3) : 0) : 2;
    // schedule records of the two strips this slot scheduled last (index, off - n_start): the
    // slot below needs the record of strip sidx - 1, and this slot can be at most one strip
    // further by the time it asks
    int rec_si = sidx == 0 ? 0 : -1000, rec_d = -g.n_start, old_si = -1000, old_d = 0;

    int HA[NR], HB[NR], F[NR], E[NR], V2[NR], NJ[NR], arow[NR];
    int F2[NR], E2[NR];     // dead (and eliminated) unless DAGP
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        HA[k] = NEV; HB[k] = NEV; F[k] = NEV; E[k] = NEV; V2[k] = NEV; NJ[k] = 0; arow[k] = 4 * ZROW;
        if (DAGP) { F2[k] = NEV; E2[k] = NEV; }
    }
    const int gn = P.gn, ge = P.ge, gn2 = P.gn2, ge2 = P.ge2;
    int pf_f2[PF];                      // prefetched F2 band words (DAGP only)
#pragma unroll
    for (int d = 0; d < PF; ++d) pf_f2[d] = NEV;
    const int floorL = localL_now ? 0 : INT_MIN;
    int prev_uh = NEV;
    int bval = INT_MIN, bstep = 0, bk = 0, bsi = 0;     // best local-mode cell of this thread

    RingEntry* ring = sm.ring + threadIdx.x;    // slot s at ring[s * CTA_THREADS]
    const char* mtx_bytes = reinterpret_cast<const char*>(sm.mtx);
    const int ipen = P.ipen;

    // Column inputs are prefetched RAW one iteration ahead (no dependent ALU
    // work behind the load) and decoded when the column is entered.
    auto col_fetch = [&](int c) -> uint2 {
        // one 8-byte load: {sig5 | sig3 << 16, code}
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
            // column b_left carries signals but pairs no residue (ke == 0)
            if (c > t.b_left) re.prof = (int) (ci.y & 0xffu) * (MTX_LD * 4);
            if (SPJ && with_sig) {
                re.s3 = hi16(ci.x);
                re.s5 = (int) (short) (lo16(ci.x) + ipen);
            }
        }
        return re;
    };

    unsigned pf_band[PF];               // prefetched {H | F} band words of columns n .. n + PF - 1
#pragma unroll
    for (int d = 0; d < PF; ++d) pf_band[d] = 0;
    uint2 nxt_col = make_uint2(0u, 0xffffffffu);

    // every strip run one after the other would need fewer iterations than this
    const int max_iter = nstr * (width + 3 * NELEM + 8) + 64;
    for (int i = -1; ; ++i) {
        if (i > max_iter) { wmax.err = 1; break; }      // scheduling bug guard: never spin on the device
        // scheduling: has the slot above scheduled strip sidx - 1 yet, and for when
        {
            const int p_rec_si = __shfl_sync(FULL, rec_si, pred_lane);
            const int p_rec_d = __shfl_sync(FULL, rec_d, pred_lane);
            const int p_old_si = __shfl_sync(FULL, old_si, pred_lane);
            const int p_old_d = __shfl_sync(FULL, old_d, pred_lane);
            if (state == 0 && (p_rec_si == sidx - 1 || p_old_si == sidx - 1)) {
                const int pd = p_rec_si == sidx - 1 ? p_rec_d : p_old_d;
                // the init iteration (j == -1) is this one at the earliest
                off = max(i + 1, pd + g.n_start + (NELEM - 1 + LAG));
                old_si = rec_si; old_d = rec_d;
                rec_si = sidx; rec_d = off - g.n_start;
                state = nsteps > 0 ? 1 : 3;
            } else if (state == 3) {
                // empty strip: nothing to run, this slot's next strip
                sidx += SPP;
                if (sidx < nstr) {
                    g = strip_geom<TRACE>(t, ml0 + NELEM * sidx);
                    nsteps = g.n_last - g.n_start + 1;
                    state = 0;
                } else
                    state = 2;
            }
        }
        if (!__any_sync(FULL, state != 2)) break;
        const bool run = state == 1;
        const int j = i - off;
        const int j8 = g.j9 - 1;
        // neighbour exchange inside a strip: bottom row of the thread above, as
        // of the previous step (every thread of the warp takes part)
        int sh_h = NEV, sh_f = NEV, sh_f2 = NEV;
        if (TPS > 1) {
            sh_h = __shfl_up_sync(FULL, (i & 1) ? HA[NR - 1] : HB[NR - 1], 1);
            sh_f = __shfl_up_sync(FULL, F[NR - 1], 1);
            if (DAGP) sh_f2 = __shfl_up_sync(FULL, F2[NR - 1], 1);
        }
        const int band_bias = g.ml + t.lw - 1;      // band entry of column c (diagonal c - ml): c - band_bias
        if (run && j == -1) {
            // One iteration before the first step: the entries of columns
            // n_start - 1 .. n_start + PF - 1 are final by now (the strip above
            // is LAG = PF + 1 iterations ahead of the write of column n_start - 1).
#pragma unroll
            for (int k = 0; k < NR; ++k) {
                HA[k] = NEV; HB[k] = NEV; F[k] = NEV; E[k] = NEV; V2[k] = NEV; NJ[k] = 0;
                if (DAGP) { F2[k] = NEV; E2[k] = NEV; }
                // rows beyond the last query residue score 0 (reference: pv_a stays 0)
                arow[k] = (row0 + k < g.j9) ? 4 * (int) aseq[(g.ml - t.a_left) + row0 + k] : 4 * ZROW;
            }
            prev_uh = NEV;
            if (sub == 0) {
#pragma unroll
                for (int d = 0; d < PF; ++d) {
                    if (d < nsteps) {
                        pf_band[d] = __ldcg(band + (g.n_start + d - band_bias));
                        if (DAGP) pf_f2[d] = __ldcg(band2 + (g.n_start + d - band_bias));
                    }
                }
                prev_uh = lo16(__ldcg(band + (g.n_start - 1 - band_bias)));
            }
            nxt_col = col_fetch(g.n_start);
            // ring pre-fill: the 15 columns left of n_start pair residues (if
            // inside the sequence) but carry no splice signal (s3_a / s5_a
            // start as zeros, src/fwd2s1_wip_simd.h:291)
#pragma unroll 1
            for (int d = 1; d < NELEM; ++d) {
                const int c = g.n_start - d;
                const RingEntry re = col_decode(col_fetch(c), c, false);
                ring[(c & 15) * CTA_THREADS] = re;
                ring[((c & 15) + 16) * CTA_THREADS] = re;
            }
        } else if (run && j >= 0) {
            const int n = g.n_start + j;
            const unsigned cur_band = pf_band[0];
            const int cur_f2 = pf_f2[0];
#pragma unroll
            for (int d = 0; d + 1 < PF; ++d) { pf_band[d] = pf_band[d + 1]; if (DAGP) pf_f2[d] = pf_f2[d + 1]; }
            const RingEntry cur_col = col_decode(nxt_col, n, n <= t.b_right);
            if (sub == 0 && j + PF < nsteps) {
                pf_band[PF - 1] = __ldcg(band + (n + PF - band_bias));
                if (DAGP) pf_f2[PF - 1] = __ldcg(band2 + (n + PF - band_bias));
            }
            if (j + 1 < nsteps) nxt_col = col_fetch(n + 1);
            const int rslot = n & 15;
            ring[rslot * CTA_THREADS] = cur_col;
            ring[(rslot + 16) * CTA_THREADS] = cur_col;
            const char* ring_hi = reinterpret_cast<const char*>(ring + (rslot + 16 - row0) * CTA_THREADS);
            int up_h, up_f, up_d, up_f2;
            if (sub == 0) {
                up_h = lo16(cur_band); up_f = hi16(cur_band); up_f2 = cur_f2;
            } else {
                up_h = sh_h; up_f = sh_f; up_f2 = sh_f2;
            }
            up_d = prev_uh;
            prev_uh = up_h;
            unsigned tw[(NR + 3) / 4];
            int sv = INT_MIN, sk = 0;
            if (i & 1)      // warp-uniform ping-pong (every thread steps once per iteration)
                strip_step<NR, TRACE, LOCAL, SPJ, DAGP>(HB, HA, F, E, F2, E2, up_f2, gn2, ge2, V2, NJ, arow, ring_hi,
                                                    mtx_bytes, sm.pen, P.pen_cap, j, up_h, up_f, up_d, gn, ge,
                                                    floorL, tw, sv, sk);
            else
                strip_step<NR, TRACE, LOCAL, SPJ, DAGP>(HA, HB, F, E, F2, E2, up_f2, gn2, ge2, V2, NJ, arow, ring_hi,
                                                    mtx_bytes, sm.pen, P.pen_cap, j, up_h, up_f, up_d, gn, ge,
                                                    floorL, tw, sv, sk);
            if (LOCAL && localR) {
                // vmax over the real rows of this step; strictly greater wins (earlier steps keep ties)
                int v = INT_MIN, kk = 0;
                if (row0 + NR <= g.j9) { v = sv; kk = sk; }
                else {
#pragma unroll
                    for (int k = NR - 1; k >= 0; --k) {
                        const int hv = (i & 1) ? HB[k] : HA[k];
                        if (row0 + k < g.j9 && hv >= v) { v = hv; kk = k; }
                    }
                }
                if (v > bval) { bval = v; bstep = n; bk = row0 + kk; bsi = sidx; }
            }
            if (TRACE) {
                unsigned char* tr = trace + ((long long) (sidx + (ml0 - t.a_left) / NELEM) * (width + TRACE_PAD) + j) * NELEM + row0;
                constexpr int NW = (NR + 3) / 4;
                if (NR == 16)
                    *reinterpret_cast<uint4*>(tr) = make_uint4(tw[0], tw[NW > 1 ? 1 : 0], tw[NW > 2 ? 2 : 0], tw[NW - 1]);
                else if (NR == 8)
                    *reinterpret_cast<uint2*>(tr) = make_uint2(tw[0], tw[NW - 1]);
                else if (NR == 4)
                    *reinterpret_cast<unsigned*>(tr) = tw[0];
                else if (NR == 2)
                    *reinterpret_cast<unsigned short*>(tr) = (unsigned short) tw[0];
                else
                    *tr = (unsigned char) tw[0];
            }
            // bottom row of the strip -> band buffer (src/fwd2s1_wip_simd.h:438-442)
            if (j8 >= row0 && j8 < row0 + NR) {
                const int kbot = j8 - row0;
                int out_h = (i & 1) ? HB[NR - 1] : HA[NR - 1];
                int out_f = F[NR - 1];
                int out_f2 = DAGP ? F2[NR - 1] : 0;
                if (kbot != NR - 1) {
#pragma unroll
                    for (int k = 0; k < NR - 1; ++k)
                        if (k == kbot) { out_h = (i & 1) ? HB[k] : HA[k]; out_f = F[k]; if (DAGP) out_f2 = F2[k]; }
                }
                const int cb = n - j8;                  // column of the bottom row
                const int r0 = cb - (g.ml + g.j9);
                if (cb > t.b_left && r0 >= t.lw && r0 <= t.up) {
                    __stcg(band + (r0 - t.lw + 1), pack16(out_h, out_f));
                    if (DAGP) __stcg(band2 + (r0 - t.lw + 1), out_f2);
                }
            }
            if (j == nsteps - 1) {
                // strip finished: this slot's next strip
                sidx += SPP;
                if (sidx < nstr) {
                    g = strip_geom<TRACE>(t, ml0 + NELEM * sidx);
                    nsteps = g.n_last - g.n_start + 1;
                    state = 0;
                } else
                    state = 2;
            }
        } else if (TPS > 1) {
            prev_uh = NEV;      // (inactive) keep the exchange registers defined
        }
        __syncwarp();
    }

    if (LOCAL && localR) {
        // reference order: strips ascending, then step, then lane (first max)
        int bv = bval;
        int bs = bstep, bkk = bk, bst = bsi;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const int ov = __shfl_xor_sync(FULL, bv, o);
            const int os = __shfl_xor_sync(FULL, bs, o);
            const int ok = __shfl_xor_sync(FULL, bkk, o);
            const int ot = __shfl_xor_sync(FULL, bst, o);
            const bool take = ov > bv || (ov == bv && (ot < bst || (ot == bst && (os < bs || (os == bs && ok < bkk)))));
            if (take) { bv = ov; bs = os; bkk = ok; bst = ot; }
        }
        // reference: k1 = lane + 1; mr = ml + k1; nr = n - k1 + 1
        if (bv > INT_MIN && bv + accscr > wmax.val) {
            wmax.val = bv + accscr;
            wmax.mr = ml0 + NELEM * bst + bkk + 1;
            wmax.nr = bs - bkk;
        }
    }
}

// shared-memory carve-up of the packed kernel
struct SmemPk {
    PkRingA* ringA;             // [2 NP][CTA_THREADS]
    PkRingB* ringB;             // [2 NP][CTA_THREADS]
    const uint2* t4;            // [PK_T4] pair table
    const PkPen* pen;           // [pen_cap + 1]
};

struct PkMonitor { int hmax, s3max, s5max; };

// ---------------------------------------------------------------------------
// run_pass with the packed cell update (gspaln_packed.cuh): same systolic chain, same band-row and
// trace-slab traffic.  Differences: a thread holds its 8 strip rows as 4 registers of two rows
// (j, j + 4); it follows the column its FIRST row sits on (n - row0), so that its private ring
// needs only the last 4 column pairs; the two gap states travel minus gn (converted where they
// meet the band rows); no local-mode bookkeeping (local problems run on the 32-bit kernel).
// ---------------------------------------------------------------------------
template <int NP, bool TRACE, bool SPJ>
__device__ void run_pass_pk(const DevParams& P, const SmemPk& sm, const DevTask& t,
                            const unsigned char* __restrict__ aseq, const ColInfo* __restrict__ cols,
                            unsigned* band, unsigned char* trace, int ml0, int nstr, WarpMax& wmax,
                            PkMonitor& mon)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    constexpr int NR = 2 * NP;                  // rows per thread: 8 (two threads per strip) or 16 (one)
    constexpr int TPS = NELEM / NR;
    constexpr int SPP = 32 / TPS;               // strip slots of the warp
    static_assert(NP == 4 || NP == 8, "4 or 8 packed registers per thread");
    const int slot = lane / TPS;
    const int sub = lane % TPS;
    const int row0 = sub * NR;
    const int pred_lane = ((slot + SPP - 1) % SPP) * TPS;
    const int width = t.up - t.lw + 3;

    int sidx = slot;
    StripGeom g = strip_geom<TRACE>(t, ml0 + NELEM * min(sidx, nstr - 1));
    int nsteps = g.n_last - g.n_start + 1;
    int off = 0;
    int state = sidx < nstr ? (sidx == 0 ? (nsteps > 0 ? 1 : 3) : 0) : 2;
    int rec_si = sidx == 0 ? 0 : -1000, rec_d = -g.n_start, old_si = -1000, old_d = 0;

    PkConst K;
    K.gn = pk_dup(P.gn); K.ge = pk_dup(P.ge); K.cgn = pk_dup(-32768 - P.gn); K.nev = pk_dup(NEV);
    K.one = 0x00010001u; K.eight = 0x00080008u; K.cap8 = pk_dup(8 * P.pen_cap);
    const unsigned nev2 = K.nev, hg0 = pk_max(K.nev, K.cgn), gt0 = pk_dup(NEV - P.gn);
    unsigned HA[NP], HB[NP], HG[NP], Ft[NP], Et[NP], V2[NP], HL[NP], arow4[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        HA[j] = nev2; HB[j] = nev2; HG[j] = hg0; Ft[j] = gt0; Et[j] = gt0; V2[j] = nev2; HL[j] = 0u;
        arow4[j] = (unsigned) ((PK_ZC * PK_NC + PK_ZC) * 8);
    }
    // incoming (row above) words: H in the half `sel` picks, likewise the vertical state
    const unsigned sel = sub ? 0x5432u : 0x5410u;
    const unsigned fconv = sub ? 0u : ((unsigned) (-P.gn) & 0xffffu);   // band rows hold F, registers F - gn
    unsigned prev_in = nev2;
    unsigned hmax = pk_dup(-32768);
    int s3max = 0, s5max = 0;

    PkRingA* ringA = sm.ringA + threadIdx.x;
    PkRingB* ringB = sm.ringB + threadIdx.x;
    const char* t4_bytes = reinterpret_cast<const char*>(sm.t4);
    const char* pen_bytes = reinterpret_cast<const char*>(sm.pen);
    const int ipen = P.ipen;
    const int nidx = P.pk_nidx;
    auto cls_of = [&](int idx) -> int { return idx < 4 ? idx : (idx == nidx ? 4 : PK_ZC); };

    auto col_fetch = [&](int c) -> uint2 {
        if (c >= t.b_left && c <= t.b_right)
            return __ldg(reinterpret_cast<const uint2*>(cols + (c - t.b_left)));
        return make_uint2(0u, 0xffffffffu);
    };
    // column c into the ring; columns left of the strip's first one pair residues but carry no signal
    auto col_push = [&](uint2 ci, int c, bool with_sig) {
        int cls = PK_ZC, s3 = 0, s5 = 0;
        if (ci.y != 0xffffffffu) {
            if (c > t.b_left) cls = cls_of((int) (ci.y & 0xffu));
            if (SPJ && with_sig) {
                s3 = hi16(ci.x);
                s5 = (int) (short) (lo16(ci.x) + ipen);
                s3max = max(s3max, s3); s5max = max(s5max, s5);
            }
        }
        pk_ring_push<NP>(ringA, ringB, CTA_THREADS, c, cls, s3, s5);
    };

    unsigned pf_band = 0;
    uint2 nxt_col = make_uint2(0u, 0xffffffffu);
    unsigned char* tr = trace;          // trace cell of the current step (TRACE)
    bool writer = false;                // this thread holds the strip's bottom row
    int kbot = 0;
    const int max_iter = nstr * (width + 3 * NELEM + 8) + 64;
    for (int i = -1; ; ++i) {
        if (i > max_iter) { wmax.err = 1; break; }
        // the schedule records only matter while some slot waits for the strip above it (or skips an
        // empty strip): one vote per iteration instead of four shuffles
        if (__any_sync(FULL, state == 0 || state == 3)) {
            const int p_rec_si = __shfl_sync(FULL, rec_si, pred_lane);
            const int p_rec_d = __shfl_sync(FULL, rec_d, pred_lane);
            const int p_old_si = __shfl_sync(FULL, old_si, pred_lane);
            const int p_old_d = __shfl_sync(FULL, old_d, pred_lane);
            if (state == 0 && (p_rec_si == sidx - 1 || p_old_si == sidx - 1)) {
                const int pd = p_rec_si == sidx - 1 ? p_rec_d : p_old_d;
                off = max(i + 1, pd + g.n_start + (NELEM - 1 + LAG));
                old_si = rec_si; old_d = rec_d;
                rec_si = sidx; rec_d = off - g.n_start;
                state = nsteps > 0 ? 1 : 3;
            } else if (state == 3) {
                sidx += SPP;
                if (sidx < nstr) {
                    g = strip_geom<TRACE>(t, ml0 + NELEM * sidx);
                    nsteps = g.n_last - g.n_start + 1;
                    state = 0;
                } else
                    state = 2;
            }
        }
        if (!__any_sync(FULL, state != 2)) break;
        const bool run = state == 1;
        const int j = i - off;
        const int j8 = g.j9 - 1;
        // the upper thread's last row (H of the previous step, vertical state) for the lower thread
        unsigned sh_h = 0, sh_f = 0;
        if (TPS > 1) {
            sh_h = __shfl_up_sync(FULL, (i & 1) ? HA[NP - 1] : HB[NP - 1], 1);
            sh_f = __shfl_up_sync(FULL, Ft[NP - 1], 1);
        }
        const int band_bias = g.ml + t.lw - 1;
        if (run && j == -1) {
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                HA[q] = nev2; HB[q] = nev2; HG[q] = hg0; Ft[q] = gt0; Et[q] = gt0; V2[q] = nev2; HL[q] = 0u;
                const int r_lo = row0 + q, r_hi = row0 + q + NP;
                const int a_lo = r_lo < g.j9 ? cls_of((int) aseq[(g.ml - t.a_left) + r_lo]) : PK_ZC;
                const int a_hi = r_hi < g.j9 ? cls_of((int) aseq[(g.ml - t.a_left) + r_hi]) : PK_ZC;
                arow4[q] = (unsigned) ((a_lo * PK_NC + a_hi) * 8);
            }
            prev_in = nev2;
            if (sub == 0) {
                if (nsteps > 0) pf_band = __ldcg(band + (g.n_start - band_bias));
                prev_in = __ldcg(band + (g.n_start - 1 - band_bias));
            }
            if (TRACE)
                tr = trace + ((long long) (sidx + (ml0 - t.a_left) / NELEM) * (width + TRACE_PAD)) * NELEM + row0;
            writer = g.j9 - 1 >= row0 && g.j9 - 1 < row0 + NR;
            kbot = g.j9 - 1 - row0;
            const int cs = g.n_start - row0;            // first column of this thread's first row
            nxt_col = col_fetch(cs);
#pragma unroll 1
            for (int d = 2 * NP - 1; d >= 1; --d) col_push(col_fetch(cs - d), cs - d, false);
        } else if (run && j >= 0) {
            const int n = g.n_start + j;
            const int cn = n - row0;
            const unsigned cur_band = pf_band;
            col_push(nxt_col, cn, cn >= g.n_start);
            if (sub == 0 && j + 1 < nsteps) pf_band = __ldcg(band + (n + 1 - band_bias));
            if (j + 1 < nsteps) nxt_col = col_fetch(cn + 1);
            const int rslot = (cn & (NP - 1)) + NP;
            const char* ra_hi = reinterpret_cast<const char*>(ringA + rslot * CTA_THREADS);
            const char* rb_hi = reinterpret_cast<const char*>(ringB + rslot * CTA_THREADS);
            const unsigned in = sub ? sh_h : cur_band;
            const unsigned in_f = sub ? sh_f : cur_band;
            unsigned tw[NP / 2];
            if (i & 1) {
                const unsigned uh0 = pk_perm(in, HA[NP - 1], sel);
                const unsigned uft0 = pk_add(pk_perm(in_f, Ft[NP - 1], 0x5432u), fconv);
                const unsigned dg0 = pk_perm(prev_in, HB[NP - 1], sel);
                strip_step_pk<NP, TRACE, SPJ>(HB, HA, HG, Ft, Et, V2, HL, arow4, ra_hi, rb_hi,
                                          CTA_THREADS * (int) sizeof(PkRingA), CTA_THREADS * (int) sizeof(PkRingB),
                                          t4_bytes, pen_bytes, uh0, uft0, dg0, K, tw, hmax);
            } else {
                const unsigned uh0 = pk_perm(in, HB[NP - 1], sel);
                const unsigned uft0 = pk_add(pk_perm(in_f, Ft[NP - 1], 0x5432u), fconv);
                const unsigned dg0 = pk_perm(prev_in, HA[NP - 1], sel);
                strip_step_pk<NP, TRACE, SPJ>(HA, HB, HG, Ft, Et, V2, HL, arow4, ra_hi, rb_hi,
                                          CTA_THREADS * (int) sizeof(PkRingA), CTA_THREADS * (int) sizeof(PkRingB),
                                          t4_bytes, pen_bytes, uh0, uft0, dg0, K, tw, hmax);
            }
            prev_in = in;
            if (TRACE) {
                if (NP == 8)
                    *reinterpret_cast<uint4*>(tr) = make_uint4(tw[0], tw[1], tw[NP / 2 - 2], tw[NP / 2 - 1]);
                else
                    *reinterpret_cast<uint2*>(tr) = make_uint2(tw[0], tw[1]);
                tr += NELEM;
            }
            // bottom row of the strip -> band buffer (src/fwd2s1_wip_simd.h:438-442)
            if (writer) {
                const int kq = kbot & (NP - 1);
                // (selects on values: an indexed read would push the register arrays to local memory)
                unsigned hw = (i & 1) ? HB[0] : HA[0], fw = Ft[0];
#pragma unroll
                for (int q = 1; q < NP; ++q) {
                    const unsigned hq = (i & 1) ? HB[q] : HA[q];
                    hw = kq == q ? hq : hw;
                    fw = kq == q ? Ft[q] : fw;
                }
                const int out_h = kbot < NP ? lo16(hw) : hi16(hw);
                const int out_f = (int) (short) ((kbot < NP ? lo16(fw) : hi16(fw)) + P.gn);
                const int cb = n - j8;
                const int r0 = cb - (g.ml + g.j9);
                if (cb > t.b_left && r0 >= t.lw && r0 <= t.up)
                    __stcg(band + (r0 - t.lw + 1), pack16(out_h, out_f));
            }
            if (j == nsteps - 1) {
                sidx += SPP;
                if (sidx < nstr) {
                    g = strip_geom<TRACE>(t, ml0 + NELEM * sidx);
                    nsteps = g.n_last - g.n_start + 1;
                    state = 0;
                } else
                    state = 2;
            }
        }
        __syncwarp();
    }
    mon.hmax = max(mon.hmax, max(lo16(hmax), hi16(hmax)));
    mon.s3max = max(mon.s3max, s3max);
    mon.s5max = max(mon.s5max, s5max);
}

// ---------------------------------------------------------------------------
// trace-code lookup for the walk (cells never evaluated read as STOP = 0)
// ---------------------------------------------------------------------------
template <int PK>
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
        const int j = cur_n + T.b_left + k - g.n_start;     // step at which row k sat on this column
        if (j < 0 || j > g.n_last - g.n_start) return 0u;
        const unsigned char* cell = trace + ((long long) s * (width + TRACE_PAD) + j) * NELEM;
        if constexpr (PK != 0) {    // raw decision bits, rows permuted inside each thread's 2 PK bytes
            constexpr int NRP = 2 * PK;
            return pk_trace_code(cell[(k / NRP) * NRP + pk_trace_byte<PK>(k % NRP)]);
        }
        return cell[k];
    }
};

// Anti_rhomb_coord<CHAR>::traceback + go_back (src/rhomb_coord.h:142-235), step = 1
template <int PK>
static __device__ int walk_trace(const DevTask& t, const unsigned char* trace, int m_abs, int n_abs,
                          int2* skl, int cap, int* status)
{
    TraceView<PK> tv{&t, trace, t.up - t.lw + 3};
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
        } else if (dir == TB_HORL) {
            bool stop = false;
            while (!(code & TB_NHOL)) { code = to_left(1); if (!code) { stop = true; break; } }
            if (!stop) code = to_left(1);
        } else if (dir == TB_VERL) {
            bool stop = false;
            while (!(code & TB_NVEL)) { code = to_upper(0); if (!code) { stop = true; break; } }
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
// PK = 4 / 8: the packed int16x2 kernel with 4 / 8 registers (8 / 16 rows) per thread.  It takes the problems marked eligible by the host (residue
// classes A, C, G, T, N only; bounded signals -- the byte behind the query codes) and reports
// status 6 for a problem whose values came too close to +32767 (see gspaln_packed.cuh); the 32-bit
// kernel (PK = 0), launched behind it on the same stream, runs everything else plus those.
constexpr int ST_NEED_EXACT = 6;

// NRT = strip rows per thread of the 32-bit kernel: 8 (two threads per strip, 16 strip slots per warp)
// for problems that can fill such a chain, 4 / 2 / 1 (8 / 4 / 2 slots) for problems whose band or
// row count allows fewer strips in flight -- the host assigns the class (DevTask::pad0) so that a
// thin or narrow problem still keeps all 32 lanes of its warp busy.
// CTAs per SM the register budget of the packed kernels is cut for.  Measured on B200 (config 2):
// NP = 8 at 2 / 3 / 4 CTAs per SM (204 / 168 / 128 registers, the last with spills): 472 / 515 / 319 GCUPS
#ifndef GSPALN_PK8_MINB
#define GSPALN_PK8_MINB 3
#endif
#ifndef GSPALN_PK4_MINB
#define GSPALN_PK4_MINB 3
#endif
template <bool TRACE, bool LOCAL, bool SPJ, bool DAGP, int PK = 0, int NRT = 8>
__global__ void __launch_bounds__(CTA_THREADS, PK == 8 ? GSPALN_PK8_MINB : (PK == 4 ? GSPALN_PK4_MINB : 3))
dp_wip_kernel(const DevParams* __restrict__ gP, const int2* __restrict__ gpen,
              const DevTask* __restrict__ tasks,
              const int* __restrict__ order, int ntasks, int* ticket,
              const unsigned char* __restrict__ apool, const ColInfo* __restrict__ cpool,
              unsigned* bandpool, long long band_slab, unsigned char* tracepool,
              long long trace_slab, int2* sklpool, DevResult* results, const int* ready)
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
    SmemPk smk;
    if (PK) {
        smk.ringA = reinterpret_cast<PkRingA*>(smem_raw);
        constexpr int PK_RING = 2 * (PK ? PK : 4);
        smk.ringB = reinterpret_cast<PkRingB*>(smem_raw + sizeof(PkRingA) * PK_RING * CTA_THREADS);
        uint2* t4 = reinterpret_cast<uint2*>(smem_raw + (sizeof(PkRingA) + sizeof(PkRingB)) * PK_RING * CTA_THREADS);
        PkPen* ppen = reinterpret_cast<PkPen*>(t4 + PK_T4);
        struct Mfun {
            const int* t; int nidx;
            __device__ int operator()(int cc, int ac) const
            {
                const int g = cc < 4 ? cc : (cc == 4 ? nidx : ZROW), q = ac < 4 ? ac : (ac == 4 ? nidx : ZROW);
                return t[g * MTX_LD + q];
            }
        } mfun{sP.mtxT, P.pk_nidx};
        for (int i = threadIdx.x; i < PK_T4; i += blockDim.x) pk_t4_entry(t4[i], i, mfun);
        for (int i = threadIdx.x; i <= P.pen_cap; i += blockDim.x) {
            const int2 e = gpen[i];
            const bool valid = e.y == -32768;
            const int pv = valid ? e.x : 0;
            ppen[i].pc = ((unsigned) pv & 0xffffu) | ((unsigned) (-32768 - pv) << 16);
            ppen[i].valid = valid ? 0xffffu : 0u;
        }
        for (int sl = 0; sl < PK_RING; ++sl) {
            smk.ringA[sl * CTA_THREADS + threadIdx.x] = PkRingA{0u, 0u, 0x80008000u, 0u};
            smk.ringB[sl * CTA_THREADS + threadIdx.x] = PkRingB{0x80008000u, (unsigned) PK_ZC};
        }
        smk.t4 = t4;
        smk.pen = ppen;
    } else {
        sm.ring = reinterpret_cast<RingEntry*>(smem_raw);
        int2* spen = reinterpret_cast<int2*>(smem_raw + sizeof(RingEntry) * RING * CTA_THREADS);
        for (int i = threadIdx.x; i <= P.pen_cap; i += blockDim.x) spen[i] = gpen[i];
        sm.pen = spen;
        sm.mtx = sP.mtxT;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    // per-warp workspace: band rows and trace matrix are reused by every
    // problem this warp picks up (the walk runs before the next forward pass)
    const long long wslot = (long long) blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    // double affine: the slab holds the {H | F} words, then as many F2 words
    unsigned* band = bandpool + wslot * band_slab * (DAGP ? 2 : 1);
    int* band2 = DAGP ? reinterpret_cast<int*>(band + band_slab) : nullptr;
    unsigned char* trace = TRACE ? tracepool + wslot * trace_slab : nullptr;

    for (;;) {
        int tk = 0;
        if (lane == 0) tk = atomicAdd(ticket, 1);
        tk = __shfl_sync(0xffffffffu, tk, 0);
        if (tk >= ntasks) break;
        const int ti = order[tk];
        const DevTask t = tasks[ti];
        if (t.kind > 1 || (t.kind == 0) != TRACE) continue;    // handled by another kernel
        if ((t.pad0 ? t.pad0 : 8) != NRT) continue;             // another chain width runs this problem
        if (!wait_inputs(ready, tk)) {
            if (lane == 0) { DevResult r; r.score = 0; r.status = 4; r.n_skl = 0; r.pad = 0; results[ti] = r; }
            continue;
        }
        {
            // who runs this problem: the packed kernel if the host marked it eligible, else (or if
            // the packed kernel gave it back) the 32-bit kernel
            const bool fast = NRT == 8 && P.pk_ok && apool[t.a_off + (t.a_right - t.a_left)] == 1;
            if (PK != 0 ? !fast : (fast && results[ti].status != ST_NEED_EXACT)) continue;
        }
        PkMonitor mon{-32768, 0, 0};
        const unsigned char* aseq = apool + t.a_off;
        const ColInfo* cols = cpool + t.col_off;
        const int width = t.up - t.lw + 3;
        const int buf_size = width + 2 * NELEM;
        const bool a_exgl = t.flags & 1, a_exgr = t.flags & 2, b_exgl = t.flags & 4, b_exgr = t.flags & 8;
        const bool LocalL = LOCAL && a_exgl && b_exgl;
        const bool LocalR = LOCAL && a_exgr && b_exgr;

        // ---- fhinitS1 (src/fwd2s1_simd.cc:163-184); entry i <-> diagonal lw - 1 + i
        for (int i = lane; i < buf_size; i += 32) { band[i] = pack16(NEV, NEV); if (DAGP) band2[i] = NEV; }
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

        // ---- strips in segments cut at the re-basing check points
        int accscr = 0;
        const int md = checkpoint(P.avmch, 0);
        int mc = md + t.a_left;
        WarpMax wmax{NEV, t.a_right, t.b_right};
        int ml0 = t.a_left;
        while (ml0 < t.a_right) {
            int nstr = (t.a_right - ml0 + NELEM - 1) / NELEM;
            if (mc >= ml0 && mc < ml0 + nstr * NELEM && ((mc - ml0) % NELEM) == 0)
                nstr = (mc - ml0) / NELEM + 1;
            if constexpr (PK != 0)
                run_pass_pk<PK, TRACE, SPJ>(P, smk, t, aseq, cols, band, trace, ml0, nstr, wmax, mon);
            else
                run_pass<NRT, TRACE, LOCAL, SPJ, DAGP>(P, sm, t, aseq, cols, band, band2, trace, ml0, nstr,
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
                        if (DAGP) {
                            int f2 = __ldcg(band2 + i) - cm;
                            f2 = i < nn ? sat16(f2) : (int) (short) f2;
                            __stcg(band2 + i, f2);
                        }
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
        bool need_exact = false;
        if (PK) {
            // high-side monitor (gspaln_packed.cuh): no add of the whole problem can have wrapped
            // unless the largest H plus the largest positive addend passes 32767
            int hm = mon.hmax, s3m = mon.s3max, s5m = mon.s5max;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                hm = max(hm, __shfl_xor_sync(0xffffffffu, hm, o));
                s3m = max(s3m, __shfl_xor_sync(0xffffffffu, s3m, o));
                s5m = max(s5m, __shfl_xor_sync(0xffffffffu, s5m, o));
            }
            need_exact = hm + P.pk_pvmax > 32767 || hm + s5m + s3m > 32767;
            if (need_exact) status = ST_NEED_EXACT;
        }
        if (TRACE && !need_exact) {
            __threadfence_block();
            __syncwarp();
            if (lane == 0)
                n_skl = walk_trace<PK>(t, trace, wmax.mr, wmax.nr, sklpool + t.skl_off, t.skl_cap, &status);
            if (lane == 0 && n_skl > t.skl_cap && status == 0) status = 1;
        }
        if (lane == 0) {
            DevResult r;
            r.score = wmax.val; r.status = wmax.err ? 4 : status; r.n_skl = n_skl; r.pad = 0;
            results[ti] = r;
        }
        __syncwarp();
    }
}

}   // namespace gspaln
