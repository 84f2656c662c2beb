// gspaln_h1.cuh -- sm_100a device code of the protein x genome spliced-alignment DP.
//
// Semantics: bit-identical (score, start point, trace-back corners) to
// SimdAln2h1::forwardH1_wip of the reference (src/fwd2h1_wip_simd.h:50-336) with
// fhinitH1 / fhlastH1 (src/fwd2h1_simd.h:546-789, mode 1) and
// Anti_rhomb_coord<SHORT> with step 3 (src/rhomb_coord.h:65-235), at the AVX2
// lane count (strips of 16 query rows).
//
// Mapping (ours): one warp owns one DP problem; TPSH = 4 neighbouring threads own
// one 16-row strip (NRH = 4 rows each, lock step, two warp shuffles per step).
// At step n strip row k sits on genome column n - 3k (the reference's lane
// geometry: one amino acid = three nucleotides), so the rows of a step are
// independent.  Every row keeps six generations of H, three of F and three of E
// in registers, because moves arrive from 1, 2, 3 columns back in the own row
// (frame shifts / codon insertion) and from 3, 4, 5, 6 steps back in the row
// above (codon deletion, frame shifts, diagonal).  Three donor slots per row
// (splice phase -1 / 0 / +1: best value + step of the last reset) replace the
// reference's hiv_v / hil_v vectors.  The 8 strips of a pass form a systolic
// chain: strip s runs 45 + HLAG steps behind strip s - 1 and takes the (H, F) of
// the row above from the diagonal-indexed band buffer (same contents as the
// reference's hv[] / fv[], including entries nobody overwrote).  Per-column
// inputs (substitution-table row, coding potential, three acceptor and three
// donor signals) are decoded once per strip into a 64-slot shared-memory ring;
// row k reads the slot of column n - 3k.  Trace codes (2 B per cell) leave as one
// 8-byte store per thread and step.
#pragma once
#include "gspaln_kernels.cuh"

namespace gspaln {

#ifndef GSPALN_NRH
#define GSPALN_NRH 4
#endif
constexpr int NRH = GSPALN_NRH;         // strip rows owned by one thread
constexpr int TPSH = NELEM / NRH;       // threads per strip
constexpr int SPPH = 32 / TPSH;         // strips per pass of one warp
constexpr int HSKEW = 3 * (NELEM - 1);  // 45: column distance between the first and last row of a strip
constexpr int HLAG = 2;                 // extra systolic lag (iterations) hiding the band-load latency
constexpr int TRACE_PAD_H = 96;         // steps per strip <= width + 91
constexpr int RINGH = 64;               // ring slots per strip (>= HSKEW + 2, power of two)
constexpr int BAND_PAD_H = 6 * NELEM;   // band entries: width + 6 * nelem (src/fwd2h1_simd.h:212)
constexpr int SIG_NONE = -100000;       // "no signal of this phase in this column"
constexpr int COL_TAIL_H = 3 * NELEM + 2;   // columns past b_right the last strip still enters

// TraceBackCode, src/rhomb_coord.h:36-61
enum : unsigned { TH_DIAG = 1, TH_HORI = 2, TH_HOR1 = 4, TH_HOR2 = 5, TH_VERT = 8, TH_VER1 = 10,
                  TH_VER2 = 11, TH_ACCM = 13, TH_ACCZ = 14, TH_ACCP = 15, TH_NHOR = 16, TH_NVER = 32,
                  TH_DONM = 64, TH_DONZ = 128, TH_DONP = 256 };

struct DevParamsH {
    int g1, g2, g3, ge;         // (short) GapW1, GapW2, GapW3, BasicGEP (all <= 0)
    int gop, gep, lgep, codonk1, gw1, gw2, gw3;     // as ints, for the end rows
    int avmch, local, spj, lcl;
    int pen_cap;                // pen table has pen_cap + 1 entries, followed by a "never" copy
    int mtxT[32 * MTX_LD];      // [tron code][amino-acid code]; row / column ZROW == 0
};

struct __align__(16) ColH {     // one genome column as the DP rows consume it (16 B, host-derived)
    short s3[3];                // acceptor signal by splice phase -1 / 0 / +1 (valid if flag bit f)
    short s5[3];                // donor signal + mean intron penalty (valid if flag bit 3 + f)
    short cv;                   // coding potential sigE of the codon that ends here
    unsigned char prof;         // tron code of at(c - 2), ZROW outside the sequence
    unsigned char flags;
};

struct ColEnd {                 // raw SGPT6 fields the two end rows read (8 B)
    short sigS, sigT, sigE, sig5;
};

struct DevTaskH {
    int kind;
    int a_left, a_right, b_left, b_right;
    int lw, up;
    int flags;                  // a_exgl | a_exgr << 2 | b_exgl << 4 | b_exgr << 6 (INEX values 0..3)
    int skl_cap;
    int b_len;
    long long a_off;            // query-code pool; element 0 == a->at(a_left)
    long long col_off;          // ColH / ColEnd pools; element 0 == column b_left
    long long skl_off;
    long long pad1;
};

struct __align__(16) RingH {    // expanded column record (32 B = two 16-byte halves)
    int s3[3];
    int prof;                   // byte offset of the substitution-table row
    int s5[3];
    int cv;
};

struct StripGeomH { int ml, j9, n_start, n_last; };

__device__ __forceinline__ StripGeomH strip_geom_h(const DevTaskH& t, int ml)
{
    // src/fwd2h1_wip_simd.h:103-110
    StripGeomH g;
    g.ml = ml;
    g.j9 = min(NELEM, t.a_right - ml);
    g.n_start = max(t.b_left, t.lw + 3 * ml);
    g.n_last = min(t.b_right, t.up + 3 * (ml + g.j9) + 1) + 3 * g.j9;
    return g;
}

struct SmemH {
    RingH* ring;                // [strips per CTA][RINGH]
    const int2* pen;            // [2 * (pen_cap + 1)]: binned penalty, then the "never" table
    const int* mtx;
};

// ---------------------------------------------------------------------------
// one segment: strips ml0, ml0 + 16, ... (nstr of them, up to the next re-basing check
// point).  The warp has SPPH strip slots; slot s runs strips s, s + SPPH, ... back to back and
// starts its next strip as soon as the strip above that one is 45 + HLAG steps ahead.
// ---------------------------------------------------------------------------
template <bool TRACE, bool LOCAL, bool SPJ>
__device__ void run_pass_h(const DevParamsH& P, const SmemH& sm, const DevTaskH& t,
                           const unsigned char* __restrict__ aseq, const ColH* __restrict__ cols,
                           unsigned* band, unsigned short* trace, int ml0, int nstr,
                           bool localL_now, bool localR, int accscr, WarpMax& wmax)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int slot = lane / TPSH;
    const int sub = lane % TPSH;
    const int row0 = sub * NRH;
    const int pred_lane = ((slot + SPPH - 1) % SPPH) * TPSH;
    const int width = t.up - t.lw + 7;

    // Step j of the slot's current strip runs in iteration j + off.  Strip s needs, at its
    // step n, the band entry the full strip above writes at ITS step n + 45; hence
    //   off_s >= off_{s-1} + (n_start_s - n_start_{s-1}) + 45 + HLAG.
    int sidx = slot;                            // strip index inside the segment
    StripGeomH g = strip_geom_h(t, ml0 + NELEM * min(sidx, nstr - 1));
    int nsteps = g.n_last - g.n_start + 1;
    int off = 0;
    // state: 0 = waiting for the strip above to be scheduled, 1 = scheduled / running,
    //        2 = no strip left, 3 = scheduled but empty (band outside the matrix)
    int state = sidx < nstr ? (sidx == 0 ? (nsteps > 0 ? 1 : 3) : 0) : 2;
    // schedule records (strip index, off - n_start) of the two strips this slot scheduled last
    int rec_si = sidx == 0 ? 0 : -1000, rec_d = -g.n_start, old_si = -1000, old_d = 0;

    int H[6][NRH], F[3][NRH], E[3][NRH], V[3][NRH], NJ[3][NRH], arow[NRH];
#pragma unroll
    for (int k = 0; k < NRH; ++k) {
#pragma unroll
        for (int a = 0; a < 6; ++a) H[a][k] = NEV;
#pragma unroll
        for (int a = 0; a < 3; ++a) { F[a][k] = NEV; E[a][k] = NEV; V[a][k] = NEV; NJ[a][k] = 0; }
        arow[k] = 4 * ZROW;
    }
    int up4 = NEV, up5 = NEV, up6 = NEV;        // H of the row above, 4 / 5 / 6 steps ago
    const int g1 = P.g1, g2 = P.g2, g3 = P.g3, ge = P.ge;
    const bool clampL = LOCAL && localL_now;
    int bval = INT_MIN, bstep = 0, bk = 0, bsi = 0;

    RingH* ring = sm.ring + (size_t) ((threadIdx.x >> 5) * SPPH + slot) * RINGH;
    const char* mtx_bytes = reinterpret_cast<const char*>(sm.mtx);
    const int pen_cap = P.pen_cap;

    auto col_fetch = [&](int c) -> uint4 {
        if (c >= t.b_left && c <= t.b_right + COL_TAIL_H)
            return __ldg(reinterpret_cast<const uint4*>(cols + (c - t.b_left)));
        return make_uint4(0u, 0u, 0u, (unsigned) ZROW << 16);
    };
    // with_sig == false: columns left of the strip's first step pair residues but carry neither
    // coding potential nor splice signals (cp_a / s3_a / s5_a start as zeros, wip.h:112)
    auto col_decode = [&](uint4 ci, bool with_sig) -> RingH {
        RingH re;
        const unsigned fl = ci.w >> 24;
        re.prof = (int) ((ci.w >> 16) & 0xffu) * (MTX_LD * 4);
        re.cv = with_sig ? lo16(ci.w) : 0;
        const bool s = SPJ && with_sig;
        re.s3[0] = (s && (fl & 1u)) ? lo16(ci.x) : SIG_NONE;
        re.s3[1] = (s && (fl & 2u)) ? hi16(ci.x) : SIG_NONE;
        re.s3[2] = (s && (fl & 4u)) ? lo16(ci.y) : SIG_NONE;
        re.s5[0] = (s && (fl & 8u)) ? hi16(ci.y) : SIG_NONE;
        re.s5[1] = (s && (fl & 16u)) ? lo16(ci.z) : SIG_NONE;
        re.s5[2] = (s && (fl & 32u)) ? hi16(ci.z) : SIG_NONE;
        return re;
    };
    // Bank-conflict-free ring: the four threads of a strip read slots 12 apart (3 columns x 4
    // rows) = 384 B = the same banks; XOR-ing the low two slot bits with the next two spreads
    // them over all four 32-byte bank groups, and odd strips store the two 16-byte halves
    // swapped so that the 8 strips of the warp fill both halves of every group.
    const int hswap = (slot & 1) * 16;
    auto ring_addr = [&](int c) -> char* {
        const int sl = c & (RINGH - 1);
        return reinterpret_cast<char*>(ring) + ((sl ^ ((sl >> 2) & 3)) * (int) sizeof(RingH));
    };
    auto ring_store = [&](int c, const RingH& re) {
        char* dst = ring_addr(c);
        *reinterpret_cast<int4*>(dst + hswap) = make_int4(re.s3[0], re.s3[1], re.s3[2], re.prof);
        *reinterpret_cast<int4*>(dst + (16 - hswap)) = make_int4(re.s5[0], re.s5[1], re.s5[2], re.cv);
    };

    unsigned nxt_band = 0;

    // every strip run one after the other would need fewer iterations than this
    const int max_iter = nstr * (width + TRACE_PAD_H + 8) + 64;
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
                off = max(i + 1, pd + g.n_start + (HSKEW + HLAG));
                old_si = rec_si; old_d = rec_d;
                rec_si = sidx; rec_d = off - g.n_start;
                state = nsteps > 0 ? 1 : 3;
            } else if (state == 3) {
                sidx += SPPH;
                if (sidx < nstr) {
                    g = strip_geom_h(t, ml0 + NELEM * sidx);
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
        // band entry of diagonal d (hv[d]): index d - lw + 3; the top row at step n reads
        // hv[r + 3], r = n - 3 (ml + 1), i.e. index n - band_bias
        const int band_bias = 3 * g.ml + t.lw - 3;
        // neighbour exchange inside a strip: last row of the thread above, three steps ago
        const int sh_h = __shfl_up_sync(FULL, H[2][NRH - 1], 1);
        const int sh_f = __shfl_up_sync(FULL, F[2][NRH - 1], 1);
        if (run && j == -1) {
#pragma unroll
            for (int k = 0; k < NRH; ++k) {
#pragma unroll
                for (int a = 0; a < 6; ++a) H[a][k] = NEV;
#pragma unroll
                for (int a = 0; a < 3; ++a) { F[a][k] = NEV; E[a][k] = NEV; V[a][k] = NEV; NJ[a][k] = 0; }
                arow[k] = (row0 + k < g.j9) ? 4 * (int) aseq[(g.ml - t.a_left) + row0 + k] : 4 * ZROW;
            }
            up4 = up5 = up6 = NEV;
            if (sub == 0) {
                const int ix = g.n_start - band_bias;
                nxt_band = __ldcg(band + ix);
                up4 = lo16(__ldcg(band + ix - 1));
                up5 = lo16(__ldcg(band + ix - 2));
                up6 = lo16(__ldcg(band + ix - 3));
            }
            // ring pre-fill: the 45 columns left of n_start, then n_start itself
            for (int d = sub + 1; d <= HSKEW; d += TPSH) {
                const int c = g.n_start - d;
                ring_store(c, col_decode(col_fetch(c), false));
            }
            if (sub == 0) ring_store(g.n_start, col_decode(col_fetch(g.n_start), true));
        } else if (run && j >= 0) {
            const int n = g.n_start + j;
            const unsigned cur_band = nxt_band;
            // both global loads of the iteration are issued here, a whole step ahead of their
            // use: __syncwarp() at the bottom waits for outstanding loads
            const bool feed = sub == 0 && j + 1 < nsteps;
            uint4 nxt_col = make_uint4(0u, 0u, 0u, (unsigned) ZROW << 16);
            if (feed) {
                nxt_band = __ldcg(band + (n + 1 - band_bias));
                nxt_col = col_fetch(n + 1);
            }
            int u3, uf;
            if (sub == 0) { u3 = lo16(cur_band); uf = hi16(cur_band); }
            else { u3 = sh_h; uf = sh_f; }

            // `if (AllZero(ph_v)) continue;` (wip.h:223) never skips in the canonical AVX2 build
            // (all_zero == _mm256_testnzc_si256(v, v) == 0, src/simd_functions.h:1057): a phase
            // slot without an acceptor always contributes nevsel
            const int liftv = NEV;
            const int2* pen_base = sm.pen;
            unsigned tw[NRH];
            int sv = INT_MIN, sk = 0;
#pragma unroll
            for (int k = NRH - 1; k >= 0; --k) {
                const char* rp = ring_addr(n - 3 * (row0 + k));
                const int4 ra = *reinterpret_cast<const int4*>(rp + hswap);         // s3[0..2], prof
                const int4 rb = *reinterpret_cast<const int4*>(rp + (16 - hswap));  // s5[0..2], cv
                const int cv = rb.w;
                const int U3 = k ? H[2][k ? k - 1 : 0] : u3;
                const int U4 = k ? H[3][k ? k - 1 : 0] : up4;
                const int U5 = k ? H[4][k ? k - 1 : 0] : up5;
                const int DV = k ? H[5][k ? k - 1 : 0] : up6;
                const int UF = k ? F[2][k ? k - 1 : 0] : uf;
                // horizontal: 1- / 2-nt frame shifts, codon insertion, extension (wip.h:123-145)
                int h = satlo(H[0][k] + g1);
                int x = satlo(H[1][k] + g2);
                unsigned eb = TH_HOR1;
                if (!(h > x)) { h = x; eb = TH_HOR2; }
                x = sat16(satlo(H[2][k] + g3) + cv);
                if (!(h > x)) { h = x; eb = TH_HORI; }
                int e = sat16(satlo(E[0][k] + ge) + cv);
                unsigned hb;
                if (e > h) { hb = 0; eb = TH_HORI; } else { e = h; hb = TH_NHOR; }
                E[0][k] = E[1][k]; E[1][k] = E[2][k]; E[2][k] = e;
                // vertical: codon deletion, frame shifts, extension (wip.h:147-178)
                int f = satlo(UF + ge);
                h = satlo(U3 + g3);
                x = satlo(U4 + g2);
                unsigned pb = TH_VERT;
                if (!(h > x)) { h = x; pb = TH_VER1; }
                x = satlo(U5 + g1);
                if (!(h > x)) { h = x; pb = TH_VER2; }
                if (f > h) pb = TH_VERT; else { f = h; hb |= TH_NVER; }
                F[2][k] = F[1][k]; F[1][k] = F[0][k]; F[0][k] = f;
                // diagonal (wip.h:180-202)
                const int pv = *reinterpret_cast<const int*>(mtx_bytes + ra.w + arow[k]);
                h = sat16(sat16(pv + DV) + cv);
                if (f > h) h = f; else pb = TH_DIAG;
                if (e > h) { h = e; pb = eb; }
                bool ab = false;
                if (SPJ) {
                    // acceptors, one candidate per splice phase (wip.h:206-246); a slot without
                    // an acceptor in this column contributes nevsel
                    const int h0 = h;
                    if (liftv > h) { h = liftv; pb = TH_ACCM; }
                    const int s3v[3] = {ra.x, ra.y, ra.z};
#pragma unroll
                    for (int fz = 0; fz < 3; ++fz) {
                        const int q0 = sat16(V[fz][k] + s3v[fz]);
                        const int2 pq = pen_base[min(j + NJ[fz][k], pen_cap)];
                        const int q = min(max(q0 + pq.x, pq.y), 32767);
                        if (q > h) { h = q; pb = TH_ACCM + fz; }
                    }
                    ab = h > h0 && max(max(ra.x, ra.y), ra.z) > SIG_NONE / 2;
                }
                if (clampL) { if (h < 0) { h = 0; hb = 0; } }
                if (SPJ) {
                    // donors (wip.h:268-296): phase +1 leaves from the diagonal predecessor
                    const int s5v[3] = {rb.x, rb.y, rb.z};
#pragma unroll
                    for (int fz = 0; fz < 3; ++fz) {
                        const int q = sat16((fz == 2 ? DV : h) + s5v[fz]);
                        const bool don = !ab && q > V[fz][k];
                        V[fz][k] = don ? q : V[fz][k];
                        NJ[fz][k] = don ? -j : NJ[fz][k];
                        if (TRACE && don) hb |= TH_DONM << fz;
                    }
                }
                if (LOCAL) {
                    if (h >= sv) { sv = h; sk = k; }        // descending k: ties end at the lowest row
                }
#pragma unroll
                for (int a = 5; a > 0; --a) H[a][k] = H[a - 1][k];
                H[0][k] = h;
                tw[k] = hb | pb;
            }
            up6 = up5; up5 = up4; up4 = u3;

            if (LOCAL && localR) {
                int v = INT_MIN, kk = 0;
                if (row0 + NRH <= g.j9) { v = sv; kk = sk; }
                else {
#pragma unroll
                    for (int k = NRH - 1; k >= 0; --k)
                        if (row0 + k < g.j9 && H[0][k] >= v) { v = H[0][k]; kk = k; }
                }
                if (v > bval) { bval = v; bstep = n; bk = row0 + kk; bsi = sidx; }
            }
            if (TRACE) {
                unsigned short* tr_base = trace + ((long long) (sidx + (ml0 - t.a_left) / NELEM) * (width + TRACE_PAD_H)) * NELEM + row0;
                if (NRH == 4)
                    *reinterpret_cast<uint2*>(tr_base + (long long) j * NELEM) =
                        make_uint2(tw[0] | (tw[1] << 16), tw[NRH - 2] | (tw[NRH - 1] << 16));
                else {
#pragma unroll
                    for (int k = 0; k < NRH; ++k) tr_base[(long long) j * NELEM + k] = (unsigned short) tw[k];
                }
            }
            // bottom row of the strip -> band buffer (wip.h:299-303)
            if (j8 >= row0 && j8 < row0 + NRH) {
                const int kbot = j8 - row0;
                int out_h = H[0][NRH - 1], out_f = F[0][NRH - 1];
#pragma unroll
                for (int k = 0; k < NRH - 1; ++k)
                    if (k == kbot) { out_h = H[0][k]; out_f = F[0][k]; }
                const int r0 = n - 3 * (g.ml + 1) - 6 * j8;
                if ((n - t.b_left) / 3 >= g.j9 && r0 >= t.lw && r0 <= t.up)
                    __stcg(band + (r0 - t.lw + 3), pack16(out_h, out_f));
            }
            // next column -> ring (read from the next iteration on)
            if (feed) ring_store(n + 1, col_decode(nxt_col, true));
            if (j == nsteps - 1) {
                // strip finished: this slot's next strip
                sidx += SPPH;
                if (sidx < nstr) {
                    g = strip_geom_h(t, ml0 + NELEM * sidx);
                    nsteps = g.n_last - g.n_start + 1;
                    state = 0;
                } else
                    state = 2;
            }
        }
        __syncwarp();
    }

    if (LOCAL && localR) {
        // reference order: strips ascending, then step, then lane (first maximum)
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
        // wip.h:256-264: k = lane + 1; mr = ml + k; nr = n - 3k + 3
        if (bv > INT_MIN && bv + accscr > wmax.val) {
            wmax.val = bv + accscr;
            wmax.mr = ml0 + NELEM * bst + bkk + 1;
            wmax.nr = bs - 3 * bkk;
        }
    }
}

// ---------------------------------------------------------------------------
// trace-code lookup + walk: Anti_rhomb_coord<SHORT>::traceback / go_back, step 3
// ---------------------------------------------------------------------------
struct TraceViewH {
    const DevTaskH* t;
    const unsigned short* trace;
    const unsigned char* row0;      // codes of row a_left written by fhinitH1 (semi-global start)
    const unsigned char* lastp;     // fhlastH1 patches of row a_right: bit0 := HORI, bit1 |= NHOR
    int width, row0_len, rw, rr;
    __device__ __forceinline__ unsigned code(int cur_m, int cur_n) const
    {
        const DevTaskH& T = *t;
        if (cur_m == 0) {
            if (!(T.flags & 3)) return cur_n >= 1 ? (unsigned) TH_HOR1 : 0u;    // initialize_m0(4)
            return cur_n < row0_len ? (unsigned) row0[cur_n] : 0u;
        }
        const int s = (cur_m - 1) / NELEM, k = (cur_m - 1) % NELEM;
        const StripGeomH g = strip_geom_h(T, T.a_left + s * NELEM);
        const int j = cur_n + T.b_left + 3 * k - g.n_start;
        unsigned c = 0u;
        if (j >= 0 && j <= g.n_last - g.n_start)
            c = trace[((long long) s * (width + TRACE_PAD_H) + j) * NELEM + k];
        if (cur_m == T.a_right - T.a_left && lastp) {
            const int r = cur_n + T.b_left - 3 * T.a_right;
            if (r >= rw && r <= rr) {
                const unsigned p = lastp[r - rw];
                if (p & 1u) c = TH_HORI;
                if (p & 2u) c |= TH_NHOR;
            }
        }
        return c;
    }
};

static __device__ int walk_trace_h(const TraceViewH& tv, int m_abs, int n_abs, int2* skl, int cap, int* status)
{
    const DevTaskH& t = *tv.t;
    int m = m_abs - t.a_left, n = n_abs - t.b_left;     // cursor
    int dm = 0, dn = 0;                                 // one-shot offset of the next record (ACCP)
    const int m_width = t.a_right - t.a_left + 1;
    const int n_width = t.b_right - t.b_left + 1 + 3 * m_width;
    int cnt = 0;
    unsigned code = 0u;
    // a start point outside the rhomb reads foreign memory in the reference: stop instead
    if (m >= 0 && m < m_width && n >= 0 && 3 * m + n < n_width) code = tv.code(m, n);
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
    const int guard = 4 * (m_width + n_width);
    while (code) {
        if (cnt < cap) skl[cnt] = make_int2(m + dm + t.a_left, n + dn + t.b_left);
        dm = dn = 0;
        if (++cnt > guard) { *status = 2; break; }
        unsigned dir = code & 15u;
        if (dir == TH_DIAG) {
            do { code = to_upper(3); } while (code && (code & 15u) == TH_DIAG);
        } else if (dir == TH_HORI) {
            bool stop = false;
            while (!(code & TH_NHOR)) { code = to_left(3); if (!code) { stop = true; break; } }
            if (!stop) {
                dir = code & 15u;
                if (dir != TH_HOR1 && dir != TH_HOR2) code = to_left(3);
            }
        } else if (dir == TH_VERT) {
            bool stop = false;
            while (!(code & TH_NVER)) { code = to_upper(0); if (!code) { stop = true; break; } }
            if (!stop) {
                dir = code & 15u;
                if (dir != TH_VER1 && dir != TH_VER2) code = to_upper(0);
            }
        } else if (dir == TH_ACCZ) {
            do { code = to_left(1); } while (code && !(code & TH_DONZ));
        } else if (dir == TH_ACCM) {
            do { code = to_left(1); } while (code && !(code & TH_DONM));
        } else if (dir == TH_ACCP) {
            do { code = to_left(1); } while (code && !(code & TH_DONP));
            // the phase +1 donor sits one codon down-right of the recorded corner
            if (code) { code = to_upper(3); dm = 1; dn = 3; }
        } else if (dir == TH_HOR1) code = to_left(1);
        else if (dir == TH_HOR2) code = to_left(2);
        else if (dir == TH_VER1) code = to_upper(1);
        else if (dir == TH_VER2) code = to_upper(2);
        else { *status = 2; break; }
    }
    if (cnt < cap) skl[cnt] = make_int2(m + dm + t.a_left, n + dn + t.b_left);
    ++cnt;
    return cnt;
}

__device__ __forceinline__ int gap_ext_pen3(const DevParamsH& P, int i) { return i > P.codonk1 ? P.lgep : P.gep; }

// ---------------------------------------------------------------------------
// persistent kernel: each warp pulls problems from a global ticket counter
// ---------------------------------------------------------------------------
template <bool TRACE, bool LOCAL, bool SPJ>
#ifndef GSPALN_H1_MINB
#define GSPALN_H1_MINB 3
#endif
__global__ void __launch_bounds__(CTA_THREADS, GSPALN_H1_MINB)
dp_h1_kernel(const DevParamsH* __restrict__ gP, const int2* __restrict__ gpen,
             const DevTaskH* __restrict__ tasks, const int* __restrict__ order, int ntasks, int* ticket,
             const unsigned char* __restrict__ apool, const ColH* __restrict__ cpool,
             const ColEnd* __restrict__ epool, unsigned* bandpool, long long band_slab,
             unsigned short* tracepool, long long trace_slab, unsigned char* rowpool, long long row_slab,
             int2* sklpool, DevResult* results, const int* ready)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ DevParamsH sP;
    {
        const int* src = reinterpret_cast<const int*>(gP);
        int* dst = reinterpret_cast<int*>(&sP);
        for (int i = threadIdx.x; i < (int) (sizeof(DevParamsH) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const DevParamsH& P = sP;
    SmemH sm;
    sm.ring = reinterpret_cast<RingH*>(smem_raw);
    int2* spen = reinterpret_cast<int2*>(smem_raw + sizeof(RingH) * RINGH * SPPH * WARPS_PER_CTA);
    for (int i = threadIdx.x; i < 2 * (P.pen_cap + 1); i += blockDim.x) spen[i] = gpen[i];
    sm.pen = spen;
    sm.mtx = sP.mtxT;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const long long wslot = (long long) blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    unsigned* band = bandpool + wslot * band_slab;
    unsigned short* trace = TRACE ? tracepool + wslot * trace_slab : nullptr;
    unsigned char* row0c = rowpool + wslot * row_slab;
    unsigned char* lastp = row0c + row_slab / 2;

    for (;;) {
        int tk = 0;
        if (lane == 0) tk = atomicAdd(ticket, 1);
        tk = __shfl_sync(0xffffffffu, tk, 0);
        if (tk >= ntasks) break;
        const int ti = order[tk];
        const DevTaskH t = tasks[ti];
        if (t.kind > 1 || (t.kind == 0) != TRACE) continue;
        if (!wait_inputs(ready, tk)) {
            if (lane == 0) { DevResult r; r.score = 0; r.status = 4; r.n_skl = 0; r.pad = 0; results[ti] = r; }
            continue;
        }
        const unsigned char* aseq = apool + t.a_off;
        const ColH* cols = cpool + t.col_off;
        const ColEnd* ce = epool + t.col_off;       // ce[c - b_left]
        const int width = t.up - t.lw + 7;
        const int buf_size = width + BAND_PAD_H;
        const int a_exgl = t.flags & 3, a_exgr = (t.flags >> 2) & 3, b_exgl = (t.flags >> 4) & 3,
                  b_exgr = (t.flags >> 6) & 3;
        const bool LocalL = LOCAL && a_exgl && b_exgl;
        const bool LocalR = LOCAL && a_exgr && b_exgr;
        auto hvp = [&](int r) -> unsigned* { return band + (r - t.lw + 3); };
        // end-row reads / writes outside the buffer (band not reaching the corner; undefined in
        // the reference) see nevsel and are dropped
        auto ldh = [&](int r) -> unsigned {
            const int ix = r - t.lw + 3;
            return (ix >= 0 && ix < buf_size) ? __ldcg(band + ix) : pack16(NEV, NEV);
        };
        auto sth = [&](int r, unsigned w) {
            const int ix = r - t.lw + 3;
            if (ix >= 0 && ix < buf_size) __stcg(band + ix, w);
        };

        // ---- fhinitH1 (src/fwd2h1_simd.h:546-689, mode 1)
        for (int i = lane; i < buf_size; i += 32) band[i] = pack16(NEV, NEV);
        const int rl = t.b_left - 3 * t.a_left;
        int rr_i = t.b_right - 3 * t.a_left;
        if (t.up < rr_i) rr_i = t.up;
        const int row0_len = max(0, rr_i - rl);
        if (TRACE && a_exgl) for (int i = lane; i < row0_len; i += 32) row0c[i] = 0;
        __syncwarp();
        if (b_exgl == 1)
            for (int r = t.lw + lane; r < rl; r += 32) *hvp(r) = pack16(0, NEV);
        __syncwarp();
        if (lane == 0) {
            int fv_rl = NEV;
            if (b_exgl == 2) fv_rl = 0;
            int r = rl;
            int rr = rr_i;
            if (!a_exgl) {
                if (b_exgl) fv_rl = 0;
                *hvp(r++) = pack16(0, fv_rl);
                *hvp(r++) = pack16((short) P.gw1, NEV);
                *hvp(r++) = pack16((short) P.gw2, NEV);
                *hvp(r++) = pack16((short) P.gw3, NEV);
                if (P.gep) {
                    const int x = (NEV - P.gw3) / P.gep + r;
                    if (x < rr) rr = x;
                    int h1 = (short) P.gw1, h2 = (short) P.gw2, h3 = (short) P.gw3;     // hv[r-3..r-1]
                    for ( ; r < rr; ++r) {
                        const int v = (short) (h1 + P.gep);
                        *hvp(r) = pack16(v, NEV);
                        h1 = h2; h2 = h3; h3 = v;
                    }
                } else {
                    const int v = (short) P.gw3;
                    for ( ; r < rr; ++r) *hvp(r) = pack16(v, NEV);
                }
            } else {
                int lend[3] = {r, r + 1, r + 2};
                int hm3, hm2, hm1;              // hv[r-3], hv[r-2], hv[r-1]
                int hs[3];
#pragma unroll
                for (int ph = 0; ph < 3; ++ph) {
                    // the first three entries are written without a range check (reference)
                    const int bn = t.b_left + 1 + ph;
                    const int sS = ce[bn - t.b_left].sigS;
                    hs[ph] = sS > 0 ? sS : 0;
                    const int fvv = (ph == 0) ? fv_rl : hi16(*hvp(r));
                    *hvp(r) = pack16(hs[ph], fvv);
                    ++r;
                }
                hm3 = hs[0]; hm2 = hs[1]; hm1 = hs[2];
                int ph = 0;
                for ( ; r < rr; ++r, ph = ph == 2 ? 0 : ph + 1) {
                    const int bn = t.b_left + 1 + (r - rl);
                    int h = hm3;
                    const int gl = r - lend[ph];
                    if (!(a_exgl & 1) && gl == 3) h = (short) (h + P.gop);
                    if (!(a_exgl & 2)) h = (short) (h + gap_ext_pen3(P, gl));
                    h = (short) (h + ce[bn - 3 - t.b_left].sigE);
                    unsigned code = 0u;
                    bool brk = false;
                    if (h < NEV) brk = true;
                    else {
                        int x = (short) (hm1 + P.gw1);
                        if (x > h) { h = x; code = TH_HOR1; }
                        x = (short) (hm2 + P.gw2);
                        if (x > h) { h = x; code = TH_HOR2; }
                        const int sS = ce[bn - t.b_left].sigS;
                        x = sS > 0 ? sS : 0;
                        if (x > h) { h = x; lend[ph] = r; }
                        else code = TH_HORI;
                    }
                    *hvp(r) = pack16(h, hi16(*hvp(r)));
                    if (brk) break;
                    if (TRACE) row0c[r - rl] = (unsigned char) code;
                    hm3 = hm2; hm2 = hm1; hm1 = h;
                }
                if (b_exgl == 2) {
                    // fv[rl] = 0 was set before hv[rl] (same entry): re-apply the F half
                    *hvp(rl) = pack16(lo16(*hvp(rl)), 0);
                }
            }
        }
        __threadfence_block();
        __syncwarp();

        // ---- strips in passes, cut at the re-basing check points (wip.h:317-328)
        int accscr = 0;
        const int md = checkpoint(P.avmch, 0);
        int mc = md + t.a_left;
        WarpMax wmax{NEV, t.a_right, t.b_right};
        int ml0 = t.a_left;
        while (ml0 < t.a_right) {
            int nstr = (t.a_right - ml0 + NELEM - 1) / NELEM;
            if (mc >= ml0 && mc < ml0 + nstr * NELEM && ((mc - ml0) % NELEM) == 0)
                nstr = (mc - ml0) / NELEM + 1;
            run_pass_h<TRACE, LOCAL, SPJ>(P, sm, t, aseq, cols, band, trace, ml0, nstr,
                                          LocalL && !accscr, LocalR, accscr, wmax);
            const int last_ml = ml0 + (nstr - 1) * NELEM;
            if (last_ml == mc) {
                int cm = lo16(__ldcg(band));
                for (int i = lane; i < width; i += 32) cm = max(cm, lo16(__ldcg(band + i)));
#pragma unroll
                for (int o = 16; o; o >>= 1) cm = max(cm, __shfl_xor_sync(0xffffffffu, cm, o));
                const int d = checkpoint(P.avmch, cm);
                if (d < md / 2) {
                    const int nn = width / NELEM * NELEM;
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
        __threadfence_block();
        __syncwarp();

        // ---- fhlastH1 (src/fwd2h1_simd.h:691-789, mode 1, no Vmf)
        int rw_l = t.lw, rr_l = t.b_right - 3 * t.a_right;
        bool have_patch = false;
        if (!LocalR || wmax.mr == t.a_right) {
            const int m3 = 3 * t.a_right;
            int rw = t.lw;
            int rf = t.b_left - m3;
            if (rf > rw) rw = rf;
            const int rr = rr_l;
            rw_l = rw;
            int maxr = rr, maxt = rr;
            int mxv = lo16(ldh((rr)));
            bool done = false;
            if (a_exgr) {
                have_patch = TRACE;
                __syncwarp();
                // Three lanes, one per nt phase (the recurrence only looks 3 entries back).  The
                // reference scans r ascending and moves mx (which starts at hv[rr]) to the first
                // entry strictly greater than *mx; entry rr itself only counts once mx has moved.
                int bv = INT_MIN, br = INT_MAX, bmaxr = rr;     // best over r < rr
                int fv_rr = mxv, k_rr = 0;                      // final value / choice at r == rr
                if (lane < 3) {
                    int glen = 0, prev = 0;
                    bool tcdn = false;
                    for (int r = rw + lane; r <= rr; r += 3) {
                        const int bn = r + m3;
                        glen += 3;
                        const unsigned wcur = ldh((r));
                        int hcur = lo16(wcur);
                        int c1 = NEV, c2 = NEV;
                        const bool far = r - rw >= 3;
                        if (far) {
                            const ColEnd e2 = ce[bn - 2 - t.b_left];
                            if (!tcdn) {
                                c1 = prev + e2.sigE;
                                if (!(a_exgr & 2)) c1 += gap_ext_pen3(P, glen);
                                if (!(a_exgr & 1) && glen == 3) c1 += P.gop;
                                if (P.lcl & 2) c2 = prev + e2.sigT;
                            }
                            tcdn = tcdn || e2.sigT > 0;
                        }
                        const int s5r = ce[bn - t.b_left].sig5;
                        const int sig5 = (LOCAL && s5r > 0) ? s5r : 0;
                        const int c0 = hcur + sig5;
                        c1 += sig5;
                        int k = 0, cb = c0;
                        if (c1 > cb) { k = 1; cb = c1; }
                        if (c2 > cb) k = 2;
                        unsigned patch = 0u;
                        if (k == 0) { glen = 0; tcdn = false; }
                        else {
                            hcur = (short) (k == 1 ? c1 - sig5 : c2);
                            patch = 1u;
                            sth(r, pack16(hcur, hi16(wcur)));
                            if (glen == 3) patch |= 2u;
                        }
                        if (TRACE) lastp[r - rw] = (unsigned char) patch;
                        if (r < rr) {
                            if (hcur > bv) { bv = hcur; br = r; bmaxr = r - (k == 2 ? 3 : 0); }
                        } else { fv_rr = hcur; k_rr = k; }
                        prev = hcur;
                    }
                }
                __syncwarp();
                {
                    const int src = rr >= rw ? (rr - rw) % 3 : 0;
                    fv_rr = __shfl_sync(0xffffffffu, fv_rr, src);
                    k_rr = __shfl_sync(0xffffffffu, k_rr, src);
                    int v0 = bv, r0 = br, m0 = bmaxr;
#pragma unroll
                    for (int q = 1; q < 3; ++q) {
                        const int ov = __shfl_sync(0xffffffffu, bv, q);
                        const int orr = __shfl_sync(0xffffffffu, br, q);
                        const int om = __shfl_sync(0xffffffffu, bmaxr, q);
                        if (ov > v0 || (ov == v0 && orr < r0)) { v0 = ov; r0 = orr; m0 = om; }
                    }
                    // (lane 0 holds the combined result)
                    if (v0 > mxv) {
                        mxv = v0; maxt = r0; maxr = m0;
                        if (rr >= rw && fv_rr > mxv) { mxv = fv_rr; maxt = rr; maxr = rr - (k_rr == 2 ? 3 : 0); }
                    } else if (rr >= rw)
                        mxv = fv_rr;        // mx still points at hv[rr]: its live value
                }
            } else if (lane == 0) {
                const int bn = rw + m3 + (rr - rw);
                const int h93 = lo16(ldh((rr - 3)));
                const int y = (short) (h93 + ce[bn - t.b_left].sigT);
                if (y > mxv) {
                    sth(rr, pack16(y, hi16(ldh((rr)))));
                    mxv = y;
                    maxr = rr - 3;
                }
            }
            maxr = __shfl_sync(0xffffffffu, maxr, 0);
            maxt = __shfl_sync(0xffffffffu, maxt, 0);
            mxv = __shfl_sync(0xffffffffu, mxv, 0);
            __syncwarp();
            if (b_exgr) {
                if (lane == 0) {
                    int rw2 = min(t.up - 1, t.b_right - 3 * t.a_left);
                    int gq[3] = {NEV, NEV, NEV};
                    int ph = 0;
                    // h runs from rw2 - 3 down to rr + 1; h[3] of the first three is read from memory
                    int h3v[3];
#pragma unroll
                    for (int q = 0; q < 3; ++q) h3v[q] = lo16(ldh((rw2 - q)));
                    for (int r = rw2 - 3; r > rr; --r, ph = ph == 2 ? 0 : ph + 1) {
                        int x = h3v[ph];
                        if (!(b_exgr & 1)) x = (short) (x + P.gop);
                        if (x > gq[ph]) gq[ph] = x;
                        if (!(b_exgr & 2)) gq[ph] = (short) (gq[ph] + P.gep);
                        int hc = lo16(ldh((r)));
                        if (hc > gq[ph]) gq[ph] = NEV;
                        else if (gq[ph] > mxv) {
                            mxv = gq[ph]; maxt = r; hc = gq[ph];
                            sth(r, pack16(hc, hi16(ldh((r)))));
                        }
                        h3v[ph] = hc;
                    }
                }
                maxt = __shfl_sync(0xffffffffu, maxt, 0);
            } else if (b_exgr == 2)
                done = true;
            if (!done) {
                if (maxr - rr > 0) wmax.mr = (t.b_right - maxr) / 3;
                else wmax.nr = maxt + m3;
            }
            wmax.val += accscr;
        }

        int status = 0, n_skl = 0;
        if (TRACE) {
            __threadfence_block();
            __syncwarp();
            if (lane == 0) {
                TraceViewH tv{&t, trace, row0c, have_patch ? lastp : nullptr, width, row0_len, rw_l, rr_l};
                n_skl = walk_trace_h(tv, wmax.mr, wmax.nr, sklpool + t.skl_off, t.skl_cap, &status);
                if (n_skl > t.skl_cap && status == 0) status = 1;
            }
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
