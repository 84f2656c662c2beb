// gspaln_h1_udh.cuh -- protein x genome unidirectional-Hirschberg forward pass on sm_100a.
//
// Semantics: SimdAln2h1::hirschbergH1_wip of the reference (src/fwd2h1_wip_simd.h:338-773) with
// fhinitH1 / fhlastH1 in mode 2 (src/fwd2h1_simd.h:546-789) and the intermediates of
// src/udh_intermediate.h, at the AVX2 lane count.  No trace matrix: every cell carries a link
// (the diagonal on which its path crossed the previous intermediate row, or started) and, in
// local mode, the row where it started.  At the n_imd intermediate rows the links are recorded
// and reset; a back-walk over those records yields the crossing records (`Dim10 cpos[]`) from
// which the host driver cuts the problem into blocks (mimd_postwork / rcsv_postwork).
//
// Mapping: this pass is rare for proteins (Aln2h1::lspH_ng takes it only when the rhombic
// volume 2 m (n + 3 m) reaches MaxVmfSpace), so it is written for exactness, not speed: one
// warp per problem, lanes 0..15 ARE the 16 vector lanes of the reference (row ml + 1 + lane on
// column n - 3 lane at step n), strips run one after the other, the shift of every lane
// buffer is a warp shuffle.  The quirks of the reference are kept: the phase +1 intron-length
// counter advances twice per step (wip.h:669 sits inside the phase loop and AllZero never
// skips in the AVX2 build), the substitution-score lane buffer is overwritten by flag vectors
// on intermediate-row strips (wip.h:596, 675, 685), LocalR reports mr one row low (wip.h:634).
#pragma once
#include "gspaln_h1.cuh"

namespace gspaln {

constexpr int END_OF_ULK_H = INT_MAX - 2;       // src/aln.h:49
constexpr int NEVSEL32_H = INT_MIN / 16 * 7;    // NEVSEL, src/cmn.h:79
constexpr int MIN_SSV_H = -1000;                // src/fwd2h1_wip_simd.h:48

struct DevUdhOutH {
    int score, status;
    int a_left, a_right, b_left, b_right;
    int pad0, pad1;
};

// DevTaskH.pad1 of a Hirschberg task: cpos offset (low 40 bits) | n_imd << 40

template <bool SPJ, bool LOCAL>
__global__ void __launch_bounds__(CTA_THREADS, 2)
dp_h1_udh_kernel(const DevParamsH* __restrict__ gP, const int2* __restrict__ gpen,
                 const DevTaskH* __restrict__ tasks, const int* __restrict__ order, int ntasks, int* ticket,
                 const unsigned char* __restrict__ apool, const ColH* __restrict__ cpool,
                 const ColEnd* __restrict__ epool, int* wspool, long long ws_slab,
                 int* cpospool, DevUdhOutH* results, const int* ready)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ DevParamsH sP;
    __shared__ int s_rlst[WARPS_PER_CTA][4];
    {
        const int* src = reinterpret_cast<const int*>(gP);
        int* dst = reinterpret_cast<int*>(&sP);
        for (int i = threadIdx.x; i < (int) (sizeof(DevParamsH) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const DevParamsH& P = sP;
    int2* spen = reinterpret_cast<int2*>(smem_raw);
    for (int i = threadIdx.x; i <= P.pen_cap; i += blockDim.x) spen[i] = gpen[i];
    __syncthreads();
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int wq = threadIdx.x >> 5;
    const int k = lane & 15;                    // vector lane (lanes 16..31 shadow 0..15 and are ignored)
    const bool vl = lane < NELEM;
    const long long wslot = (long long) blockIdx.x * WARPS_PER_CTA + wq;
    int* ws = wspool + wslot * ws_slab;
    const char* mtx_bytes = reinterpret_cast<const char*>(sP.mtxT);
    const int g1 = P.g1, g2 = P.g2, g3 = P.g3, ge = P.ge;

    for (;;) {
        int tk = 0;
        if (lane == 0) tk = atomicAdd(ticket, 1);
        tk = __shfl_sync(FULL, tk, 0);
        if (tk >= ntasks) break;
        const int ti = order[tk];
        const DevTaskH t = tasks[ti];
        if (t.kind != 2) continue;
        if (!wait_inputs(ready, tk)) {
            if (lane == 0) {
                // inputs never arrived: no crossing records (the driver skips the post-work)
                DevUdhOutH r; memset(&r, 0, sizeof(r)); r.status = 4; r.score = INT_MIN / 16 * 7; results[ti] = r;
                int* cp = cpospool + (t.pad1 & ((1ll << 40) - 1));
                const int nim = (int) (t.pad1 >> 40);
                for (int i = 0; i <= nim; ++i) cp[10 * i] = cp[10 * i + 2] = INT_MAX - 2;
            }
            continue;
        }
        const int n_im = (int) (t.pad1 >> 40);
        int* cpos = cpospool + (t.pad1 & ((1ll << 40) - 1));
        const unsigned char* aseq = apool + t.a_off;
        const ColH* cols = cpool + t.col_off;
        const ColEnd* ce = epool + t.col_off;
        const int lw = t.lw, up = t.up;
        const int width = up - lw + 7;
        const int buf_size = width + BAND_PAD_H;
        int a_left = t.a_left, a_right = t.a_right, b_left = t.b_left, b_right = t.b_right;
        const int a_exgl = t.flags & 3, a_exgr = (t.flags >> 2) & 3, b_exgl = (t.flags >> 4) & 3,
                  b_exgr = (t.flags >> 6) & 3;
        const bool LocalL = LOCAL && a_exgl && b_exgl;
        const bool LocalR = LOCAL && a_exgr && b_exgr;
        // workspace: hv|fv (packed), hc, fc, hb|fb (packed), then n_im x 4 x width link records
        unsigned* bandv = reinterpret_cast<unsigned*>(ws);
        int* bandhc = ws + buf_size;
        int* bandfc = ws + 2 * buf_size;
        unsigned* bandb = reinterpret_cast<unsigned*>(ws + 3 * buf_size);
        int* imdbase = ws + 4 * buf_size;
        auto IX = [&](int r) { return r - lw + 3; };
        auto hlnk = [&](int i, int d, int r) -> int* { return imdbase + ((long long) i * 4 + d) * width + (r - lw + 1); };
        auto vlnk = [&](int i, int d, int r) -> int* { return imdbase + ((long long) i * 4 + 2 + d) * width + (r - lw + 1); };

        // ---- fhinitH1, mode 2
        for (int i = lane; i < buf_size; i += 32) {
            bandv[i] = pack16(NEV, NEV);
            bandhc[i] = 0; bandfc[i] = 0;
            bandb[i] = pack16((short) a_left, (short) a_left);
        }
        for (long long i = lane; i < (long long) n_im * 4 * width; i += 32) imdbase[i] = END_OF_ULK_H;
        for (int i = lane; i < 10 * (n_im + 1); i += 32) cpos[i] = END_OF_ULK_H;
        __syncwarp();
        const int rl0 = b_left - 3 * a_left;
        {
            const int rre = a_exgl ? rl0 : up;
            for (int r = lw + lane; r < rre; r += 32) bandhc[IX(r)] = r;
            for (int r = rl0 - lane; r >= lw; r -= 32)
                bandb[IX(r)] = pack16((short) (a_left + (rl0 - r) / 3), (short) a_left);
            if (b_exgl == 1) for (int r = lw + lane; r < rl0; r += 32) bandv[IX(r)] = pack16(0, NEV);
        }
        __syncwarp();
        if (lane == 0) {
            int rr = b_right - 3 * a_left;
            if (up < rr) rr = up;
            int r = rl0;
            int fv_rl = NEV;
            if (b_exgl == 2) { fv_rl = 0; bandfc[IX(rl0)] = rl0; }
            if (!a_exgl) {
                if (b_exgl) { fv_rl = 0; bandfc[IX(r)] = bandhc[IX(r)]; }
                bandv[IX(r++)] = pack16(0, fv_rl);
                bandv[IX(r++)] = pack16((short) P.gw1, NEV);
                bandv[IX(r++)] = pack16((short) P.gw2, NEV);
                bandv[IX(r++)] = pack16((short) P.gw3, NEV);
                if (P.gep) {
                    const int x = (NEV - P.gw3) / P.gep + r;
                    if (x < rr) rr = x;
                    int h1 = (short) P.gw1, h2 = (short) P.gw2, h3 = (short) P.gw3;
                    for ( ; r < rr; ++r) {
                        const int v = (short) (h1 + P.gep);
                        bandv[IX(r)] = pack16(v, NEV);
                        h1 = h2; h2 = h3; h3 = v;
                    }
                } else {
                    for ( ; r < rr; ++r) bandv[IX(r)] = pack16((short) P.gw3, NEV);
                }
            } else {
                int lend[3] = {r, r + 1, r + 2};
                int hs[3], cs[3];
#pragma unroll
                for (int ph = 0; ph < 3; ++ph) {
                    const int sS = ce[1 + ph].sigS;                     // column b_left + 1 + ph
                    hs[ph] = sS > 0 ? sS : 0;
                    cs[ph] = r;
                    bandv[IX(r)] = pack16(hs[ph], ph == 0 ? fv_rl : hi16(bandv[IX(r)]));
                    bandhc[IX(r)] = r;
                    ++r;
                }
                int hm3 = hs[0], hm2 = hs[1], hm1 = hs[2], cm3 = cs[0], cm2 = cs[1], cm1 = cs[2];
                int ph = 0;
                for ( ; r < rr; ++r, ph = ph == 2 ? 0 : ph + 1) {
                    const int bnl = 1 + (r - rl0);                      // column bn - b_left
                    int h = hm3, c = cm3;
                    const int gl = r - lend[ph];
                    if (!(a_exgl & 1) && gl == 3) h = (short) (h + P.gop);
                    if (!(a_exgl & 2)) h = (short) (h + gap_ext_pen3(P, gl));
                    h = (short) (h + ce[bnl - 3].sigE);
                    bool brk = false;
                    if (h < NEV) brk = true;
                    else {
                        int x = (short) (hm1 + P.gw1);
                        if (x > h) { h = x; c = cm1; }
                        x = (short) (hm2 + P.gw2);
                        if (x > h) { h = x; c = cm2; }
                        const int sS = ce[bnl].sigS;
                        x = sS > 0 ? sS : 0;
                        if (x > h) { h = x; lend[ph] = r; c = r; }
                    }
                    bandv[IX(r)] = pack16(h, hi16(bandv[IX(r)]));
                    bandhc[IX(r)] = c;
                    if (brk) break;
                    hm3 = hm2; hm2 = hm1; hm1 = h; cm3 = cm2; cm2 = cm1; cm1 = c;
                }
            }
            s_rlst[wq][0] = s_rlst[wq][1] = s_rlst[wq][2] = INT_MAX;
        }
        __threadfence_block();
        __syncwarp();

        // ---- intermediates (wip.h:366-371)
        const int mmstep = (a_right - a_left + n_im) / (n_im + 1);
        int ii = 0;
        int imd_mi = a_left + mmstep;
        int mm = a_left + (imd_mi - a_left - 1) / NELEM * NELEM;
        int k9 = imd_mi - mm, k8 = k9 - 1;

        int accscr = 0;
        const int md = checkpoint(P.avmch, 0);
        int mc = md + a_left;
        int mh_val = NEV, mh_ulk = END_OF_ULK_H, mh_ml = (short) a_left, mh_mr = (short) a_right, mh_nr = b_right;

        // link lanes persist across strips (the reference never resets them)
        int HC[6] = {0, 0, 0, 0, 0, 0}, FC[3] = {0, 0, 0}, EC[3] = {0, 0, 0};

        for (int ml = a_left; ml < a_right; ml += NELEM) {
            const int j9 = min(NELEM, a_right - ml);
            const int j8 = j9 - 1;
            int n = max(b_left, lw + 3 * ml);
            const int n9 = min(b_right, up + 3 * (ml + j9) + 1) + 3 * j9;
            const int n_start = n;
            const int mp1 = ml + 1;
            int donor_r[3];
            donor_r[0] = donor_r[1] = donor_r[2] = n - 3 * mp1;
            const bool is_imd_ = ml == mm;
            // lane state (ages: index 0 = previous step)
            int H[6], F[3], E[3], HB[6], FB[3], EB[3];
#pragma unroll
            for (int a = 0; a < 6; ++a) { H[a] = NEV; HB[a] = 0; }
#pragma unroll
            for (int a = 0; a < 3; ++a) { F[a] = NEV; E[a] = NEV; FB[a] = 0; EB[a] = 0; }
            int hiv[3] = {NEV, NEV, NEV}, hib[3] = {0, 0, 0}, hic[3] = {0, 0, 0}, hil[3] = {0, 0, 0};
            int sm = 0;
            const int arow = (vl && k < j9) ? 4 * (int) aseq[(ml - a_left) + k] : 4 * ZROW;

            for ( ; n < n9; ++n) {
                const int r = n - 3 * mp1;
                const int ph = ((n + 3 * mp1) % 6) % 3;
                const int rj = r - 6 * k8;
                const int nb = max(0, n - b_right + 1);
                const int kb = (nb - 1) / 3;
                const int ke = min(j9, (n - b_left) / 3);
                const bool is_imd = is_imd_ && rj >= lw && rj <= up;
                // this lane's column (entered the reference's shift registers at step c)
                const int c = n - 3 * k;
                uint4 ci = make_uint4(0u, 0u, 0u, (unsigned) ZROW << 16);
                if (c >= b_left && c <= b_right + COL_TAIL_H)
                    ci = __ldg(reinterpret_cast<const uint4*>(cols + (c - b_left)));
                const bool entered = c >= n_start;
                const unsigned fl = entered ? (ci.w >> 24) : 0u;
                const int cv = entered ? lo16(ci.w) : 0;
                // row above: shuffles (lane 0: band)
                int U3 = __shfl_up_sync(FULL, H[2], 1), U4 = __shfl_up_sync(FULL, H[3], 1);
                int U5 = __shfl_up_sync(FULL, H[4], 1), DV = __shfl_up_sync(FULL, H[5], 1);
                int UF = __shfl_up_sync(FULL, F[2], 1);
                int U3c = __shfl_up_sync(FULL, HC[2], 1), U4c = __shfl_up_sync(FULL, HC[3], 1);
                int U5c = __shfl_up_sync(FULL, HC[4], 1), DVc = __shfl_up_sync(FULL, HC[5], 1);
                int UFc = __shfl_up_sync(FULL, FC[2], 1);
                int U3b = 0, U4b = 0, U5b = 0, DVb = 0, UFb = 0;
                if (LOCAL) {
                    U3b = __shfl_up_sync(FULL, HB[2], 1); U4b = __shfl_up_sync(FULL, HB[3], 1);
                    U5b = __shfl_up_sync(FULL, HB[4], 1); DVb = __shfl_up_sync(FULL, HB[5], 1);
                    UFb = __shfl_up_sync(FULL, FB[2], 1);
                }
                if (k == 0) {
                    const unsigned w3 = bandv[IX(r + 3)];
                    U3 = lo16(w3); UF = hi16(w3);
                    U4 = lo16(bandv[IX(r + 2)]); U5 = lo16(bandv[IX(r + 1)]); DV = lo16(bandv[IX(r)]);
                    U3c = bandhc[IX(r + 3)]; U4c = bandhc[IX(r + 2)]; U5c = bandhc[IX(r + 1)]; DVc = bandhc[IX(r)];
                    UFc = bandfc[IX(r + 3)];
                    if (LocalL) {
                        const unsigned b3 = bandb[IX(r + 3)];
                        U3b = lo16(b3); UFb = hi16(b3);
                        U4b = lo16(bandb[IX(r + 2)]); U5b = lo16(bandb[IX(r + 1)]); DVb = lo16(bandb[IX(r)]);
                    }
                }
                // ---- horizontal (wip.h:412-461)
                int h = satlo(H[0] + g1), hb_ = HB[0], hc_ = HC[0];
                int x = satlo(H[1] + g2);
                if (!(h > x)) { h = x; hb_ = HB[1]; hc_ = HC[1]; }
                x = sat16(satlo(H[2] + g3) + cv);
                if (!(h > x)) { h = x; hb_ = HB[2]; hc_ = HC[2]; }
                int e = sat16(satlo(E[0] + ge) + cv), eb = EB[0], ec = EC[0];
                if (!(e > h)) { e = h; eb = hb_; ec = hc_; }
                // ---- vertical (wip.h:463-530)
                int f = satlo(UF + ge), fb = UFb, fc = UFc;
                x = satlo(U3 + g3);
                if (!(f > x)) { f = x; fb = U3b; fc = U3c; }
                x = satlo(U4 + g2);
                if (!(f > x)) { f = x; fb = U4b; fc = U4c; }
                x = satlo(U5 + g1);
                if (!(f > x)) { f = x; fb = U5b; fc = U5c; }
                // ---- diagonal (wip.h:532-558)
                if (nb) sm = 0;
                if (k >= kb && k < ke)
                    sm = *reinterpret_cast<const int*>(mtx_bytes + (int) ((ci.w >> 16) & 0xffu) * (MTX_LD * 4) + arow);
                h = sat16(sat16(sm + DV) + cv); hb_ = DVb; hc_ = DVc;
                int pb = 0;
                if (f > h) { h = f; hb_ = fb; hc_ = fc; pb = 2; }
                if (e > h) { h = e; hb_ = eb; hc_ = ec; pb = 1; }
                // ---- acceptors (wip.h:560-606): both phase slots, all three phases each
                int ab = 0;
                if (SPJ) {
                    const int s3v[3] = {lo16(ci.x), hi16(ci.x), lo16(ci.y)};
                    const bool two = (fl & 1u) && (fl & 4u);            // phs3 == 2: -1 in slot 0, +1 in slot 1
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
                        for (int fz = 0; fz < 3; ++fz) {
                            const bool has = (fl >> fz) & 1u;
                            const bool mine = has && (fz != 2 ? kk == 0 : (two ? kk == 1 : kk == 0));
                            int qv = NEV;
                            if (mine) {
                                // length filter and binned penalty come from the table: {penalty | invalid, clamp}
                                const int2 pq = spen[min(hil[fz], P.pen_cap)];
                                const int q0 = sat16(hiv[fz] + s3v[fz]);
                                qv = min(max(q0 + pq.x, pq.y), 32767);
                            }
                            const bool upd = qv > h;
                            if (upd) { h = qv; if (LocalL) hb_ = hib[fz]; hc_ = hic[fz]; }
                            ab |= upd ? 1 : 0;
                            if (is_imd) {
                                sm = upd ? 1 : 0;                       // Store(sm_a, qv_v)
                                if (vl && k == k8 && upd) {
                                    *hlnk(ii, 0, rj) = donor_r[fz];
                                    *hlnk(ii, 1, rj) = donor_r[fz] + width;
                                    s_rlst[wq][ph] = rj;
                                }
                            }
                        }
                    }
                }
                // ---- store H; left / right ends (wip.h:608-639)
                if (LocalL && !accscr && h < 0) h = 0;
                int hnb = hb_, hnc = hc_;       // lane-buffer copies (patched below); donors use hb_ / hc_
                if (LocalL && !accscr && k >= kb && k < ke && h == 0) { hnb = (short) (ml + k); hnc = r - 6 * k; }
                if (LocalR) {
                    int bv = (vl && k < j9) ? h : INT_MIN, bl = k;
#pragma unroll
                    for (int o = 8; o; o >>= 1) {
                        const int ov = __shfl_xor_sync(FULL, bv, o);
                        const int ol = __shfl_xor_sync(FULL, bl, o);
                        if (ov > bv || (ov == bv && ol < bl)) { bv = ov; bl = ol; }
                    }
                    bv = __shfl_sync(FULL, bv, 0); bl = __shfl_sync(FULL, bl, 0);
                    const int w_ml = __shfl_sync(FULL, hnb, bl), w_ulk = __shfl_sync(FULL, hnc, bl);
                    if (bv + accscr > mh_val) {
                        mh_val = bv + accscr; mh_ml = w_ml; mh_ulk = w_ulk;
                        mh_mr = ml + (bl + 1) + 1;              // literal: ml + k + 1 with k the 1-based lane
                        mh_nr = n - 3 * (bl + 1);
                    }
                }
                // ---- donors (wip.h:643-681)
                if (SPJ) {
                    const int s5v[3] = {hi16(ci.y), lo16(ci.z), hi16(ci.z)};
                    const bool two = (fl & 8u) && (fl & 32u);
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
                        for (int fz = 0; fz < 3; ++fz) {
                            if (kk == 1 && fz < 2) continue;
                            const bool has = (fl >> (3 + fz)) & 1u;
                            const bool mine = has && (fz != 2 ? kk == 0 : (two ? kk == 1 : kk == 0));
                            int pvv = NEV;
                            if (mine && !ab) pvv = sat16((fz == 2 ? DV : h) + s5v[fz]);
                            const bool upd = pvv > hiv[fz];
                            if (upd) { hiv[fz] = pvv; hil[fz] = 0; if (LocalL) hib[fz] = hb_; hic[fz] = hc_; }
                            hil[fz] = min(hil[fz] + 1, 32767);
                            if (is_imd) {
                                sm = upd ? 1 : 0;
                                if (vl && k == k8 && upd) donor_r[fz] = rj;
                            }
                        }
                    }
                }
                // ---- intermediate row (wip.h:684-694)
                int fnc = fc;
                if (is_imd) {
                    sm = ab;
                    if (vl && k == k8) {
                        if (pb == 0) s_rlst[wq][ph] = rj;
                        if (!ab && pb == 1) *hlnk(ii, 0, rj) = s_rlst[wq][ph];
                        *vlnk(ii, 0, rj) = hnc;
                        hnc = rj;
                        *vlnk(ii, 1, rj) = fnc;
                        fnc = rj + width;
                    }
                }
                // ---- shift the lane buffers
#pragma unroll
                for (int a = 5; a > 0; --a) { H[a] = H[a - 1]; HC[a] = HC[a - 1]; if (LOCAL) HB[a] = HB[a - 1]; }
                H[0] = h; HC[0] = hnc; if (LOCAL) HB[0] = LocalL ? hnb : HB[0];
                F[2] = F[1]; F[1] = F[0]; F[0] = f;
                FC[2] = FC[1]; FC[1] = FC[0]; FC[0] = fnc;
                if (LOCAL) { FB[2] = FB[1]; FB[1] = FB[0]; if (LocalL) FB[0] = fb; }
                E[0] = E[1]; E[1] = E[2]; E[2] = e;
                EC[0] = EC[1]; EC[1] = EC[2]; EC[2] = ec;
                if (LOCAL) { EB[0] = EB[1]; EB[1] = EB[2]; if (LocalL) EB[2] = eb; }
                // ---- bottom row -> band (wip.h:697-707)
                const int r0 = r - 6 * j8;
                if (vl && k == j8 && j9 == ke && lw <= r0 && r0 <= up) {
                    bandv[IX(r0)] = pack16(h, f);
                    bandhc[IX(r0)] = hnc; bandfc[IX(r0)] = fnc;
                    if (LocalL) bandb[IX(r0)] = pack16(hnb, fb);
                }
                __syncwarp();
            }
            __threadfence_block();
            __syncwarp();
            if (ml == mc) {
                int cm = lo16(bandv[0]);
                for (int i = lane; i < width; i += 32) cm = max(cm, lo16(bandv[i]));
#pragma unroll
                for (int o = 16; o; o >>= 1) cm = max(cm, __shfl_xor_sync(FULL, cm, o));
                const int d = checkpoint(P.avmch, cm);
                if (d < md / 2) {
                    const int nn = width / NELEM * NELEM;
                    for (int i = lane; i < width; i += 32) {
                        const unsigned w = bandv[i];
                        int hh = lo16(w) - cm, ff = hi16(w) - cm;
                        if (i < nn) { hh = sat16(hh); ff = sat16(ff); }
                        else { hh = (short) hh; ff = (short) ff; }
                        bandv[i] = pack16(hh, ff);
                    }
                    accscr += cm;
                    mc += md;
                } else
                    mc += d;
                __syncwarp();
            }
            if (is_imd_ && ++ii < n_im) {
                imd_mi += mmstep;
                mm = a_left + (imd_mi - a_left - 1) / NELEM * NELEM;
                k9 = imd_mi - mm;
                k8 = k9 - 1;
            }
        }
        __threadfence_block();
        __syncwarp();

        // ---- ends + back-walk: lane 0
        if (lane == 0) {
            auto mi_of = [&](int i) { return t.a_left + mmstep * (i + 1); };      // Udh_Imds ctor
            int status = 0;
            if (LocalR && mh_mr < a_right) {
                a_right = mh_mr;
                b_right = mh_nr;
            } else {
                // fhlastH1, mode 2 (src/fwd2h1_simd.h:691-789)
                int glen[3] = {0, 0, 0};
                bool tcdn[3] = {false, false, false};
                const int m3 = 3 * a_right;
                int rw = lw;
                int rf = b_left - m3;
                if (rf > rw) rw = rf; else rf = rw;
                const int rr = b_right - m3;
                int maxr = rr, maxt = rr;
                auto HV = [&](int r) -> int { const int ix = IX(r); return (ix >= 0 && ix < buf_size) ? lo16(bandv[ix]) : NEV; };
                auto SETHV = [&](int r, int v) { const int ix = IX(r); if (ix >= 0 && ix < buf_size) bandv[ix] = pack16(v, hi16(bandv[ix])); };
                bool early = false;
                if (a_exgr) {
                    int ph = 0;
                    for (int r = rw; r <= rr; ++r, ++rf, ph = ph == 2 ? 0 : ph + 1) {
                        const int bn = r + m3;
                        glen[ph] += 3;
                        int hcur = HV(r);
                        int c0 = hcur, c1 = NEV, c2 = NEV;
                        if (rf - rw >= 3 && !tcdn[ph]) {
                            c1 = HV(r - 3) + ce[bn - 2 - b_left].sigE;
                            if (!(a_exgr & 2)) c1 += gap_ext_pen3(P, glen[ph]);
                            if (!(a_exgr & 1) && glen[ph] == 3) c1 += P.gop;
                            if (P.lcl & 2) c2 = HV(r - 3) + ce[bn - 2 - b_left].sigT;
                        }
                        if (rf - rw >= 3) tcdn[ph] = tcdn[ph] || ce[bn - 2 - b_left].sigT > 0;
                        const int s5r = ce[bn - b_left].sig5;
                        const int sig5 = (LOCAL && s5r > 0) ? s5r : 0;
                        c0 += sig5; c1 += sig5;
                        int kq = 0, cb = c0;
                        if (c1 > cb) { kq = 1; cb = c1; }
                        if (c2 > cb) kq = 2;
                        if (kq == 0) { glen[ph] = 0; tcdn[ph] = false; }
                        else {
                            hcur = (short) (kq == 1 ? c1 - sig5 : c2);
                            SETHV(r, hcur);
                        }
                        if (hcur > HV(maxt)) { maxt = r; maxr = rf - (kq == 2 ? 3 : 0); }   // *h > *mx
                    }
                } else {
                    const int bn = rw + m3 + (rr - rw);
                    const int y = (short) (HV(rr - 3) + ce[bn - b_left].sigT);
                    if (y > HV(rr)) { SETHV(rr, y); maxr = rr - 3; }
                }
                if (b_exgr) {
                    int rw2 = min(up - 1, b_right - 3 * t.a_left);
                    int gq[3] = {NEV, NEV, NEV};
                    int ph = 0;
                    for (int r = rw2 - 3; r > rr; --r, ph = ph == 2 ? 0 : ph + 1) {
                        int x = HV(r + 3);
                        if (!(b_exgr & 1)) x = (short) (x + P.gop);
                        if (x > gq[ph]) gq[ph] = x;
                        if (!(b_exgr & 2)) gq[ph] = (short) (gq[ph] + P.gep);
                        if (HV(r) > gq[ph]) gq[ph] = NEV;
                        else if (gq[ph] > HV(maxt)) { maxt = r; SETHV(r, gq[ph]); }
                    }
                } else if (b_exgr == 2)
                    early = true;
                int ret = rr;
                if (!early) {
                    const int ixm = IX(maxr), ixt = IX(maxt);
                    if (ixm >= 0 && ixm < buf_size && ixt >= 0 && ixt < buf_size) {
                        bandb[ixt] = pack16(lo16(bandb[ixm]), hi16(bandb[ixt]));
                        mh_ulk = bandhc[ixm];
                    }
                    if (maxr - rr > 0) mh_mr = (b_right - maxr) / 3;
                    else mh_nr = maxt + m3;
                    ret = maxt;
                }
                mh_val += accscr;
                const int ixr = IX(ret);
                mh_ml = LocalL ? ((ixr >= 0 && ixr < buf_size) ? lo16(bandb[ixr]) : t.a_left) : t.a_left;
                a_right = mh_mr;
                b_right = mh_nr;
            }
            // ---- back-walk over the intermediates (wip.h:736-771)
            int i = n_im;
            while (--i >= 0 && mi_of(i) > a_right) ;
            if (i < 0 && mi_of(0) > a_right) cpos[2] = b_right;
            int r = mh_ulk;
            for ( ; i >= 0 && mi_of(i) > mh_ml; --i) {
                int cc = 0, d = 0;
                for ( ; r > up; r -= width) ++d;
                if (d > 1 || r < lw - 1 || r >= lw - 1 + width) { status = 2; break; }   // foreign memory in the reference
                if (*vlnk(i, d, r) < END_OF_ULK_H) {
                    cpos[10 * i + cc++] = mi_of(i);
                    cpos[10 * i + cc++] = d > 0 ? 1 : 0;
                    const int mm3 = 3 * mi_of(i);
                    for (int rp = *hlnk(i, d, r); lw <= rp && rp < up && r != rp && cc < 8; rp = *hlnk(i, d, r = rp))
                        cpos[10 * i + cc++] = r + mm3;
                    cpos[10 * i + cc++] = r + mm3;
                    cpos[10 * i + cc] = END_OF_ULK_H;
                    r = *vlnk(i, d, r);
                    if (r == END_OF_ULK_H) break;
                } else
                    cpos[10 * i + 0] = END_OF_ULK_H;
            }
            for ( ; r > up; r -= width) ;
            a_left = t.a_left; b_left = t.b_left;
            if (LocalL) {
                a_left = mh_ml;
                b_left = r + 3 * a_left;
            } else {
                const int rl = b_left - 3 * a_left;
                if (b_exgl && rl > r) {
                    a_left = (b_left - r) / 3;
                    for (int j = 0; j < n_im && mi_of(j) < a_left; ++j) cpos[10 * j + 0] = END_OF_ULK_H;
                }
                if (a_exgl && rl < r) b_left = 3 * a_left + r;
            }
            ++i;
            bool bad = false;
            if (i >= 0 && i < n_im && mi_of(i) < a_left) bad = true;
            if (!bad && cpos[10 * i + 2] < b_left) bad = true;
            DevUdhOutH o;
            o.score = bad ? NEVSEL32_H : mh_val; o.status = status;
            o.a_left = a_left; o.a_right = a_right; o.b_left = b_left; o.b_right = b_right;
            o.pad0 = o.pad1 = 0;
            results[ti] = o;
        }
        __syncwarp();
    }
}

}   // namespace gspaln
