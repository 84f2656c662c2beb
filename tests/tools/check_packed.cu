// Host-side checker (test infrastructure): the packed int16x2 cell update strip_step_pk
// (spaln_b200/csrc/gspaln_packed.cuh) against the 32-bit strip_step (gspaln_kernels.cuh) it must
// reproduce bit for bit -- H, both gap states, donor value, intron-length counter and trace code of
// all 8 rows of a thread, step by step, over random and adversarial inputs (values at both ends of
// the int16 range, extreme signals).  Also checks the monitor argument: whenever the two disagree,
// max H + largest positive addend must have passed 32767.
//
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a [-DGSPALN_NR=16] -o check_packed \
//        tests/tools/check_packed.cu && ./check_packed [runs] [seed]
// (default: 8 rows per thread = 4 packed registers; -DGSPALN_NR=16: 16 rows = 8 registers)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "../../spaln_b200/csrc/gspaln_kernels.cuh"
#include "../../spaln_b200/csrc/gspaln_packed.cuh"

using namespace gspaln;

struct Params { int gn, ge, ipen, nquant, mil, quant[8], mean[8]; int mtx[6][6]; };

constexpr int NP = NR / 2;      // registers per thread of the packed form (compile with -DGSPALN_NR=16 for NP = 8)

template <bool TRACE, bool SPJ>
static long run_case(std::mt19937& rng, const Params& PR, int regime, int steps, long& trips, long& cmp)
{
    auto U = [&](int lo, int hi) { return (int) (lo + (long long) (rng() % (unsigned) (hi - lo + 1))); };
    // ---- tables
    static int mtxT[32 * MTX_LD];
    memset(mtxT, 0, sizeof(mtxT));
    // classes 0..3 = table index 0..3, N = index 16, zero row = ZROW
    const int idx_of[6] = {0, 1, 2, 3, 16, ZROW};
    for (int g = 0; g < 5; ++g)
        for (int q = 0; q < 5; ++q) mtxT[idx_of[g] * MTX_LD + idx_of[q]] = PR.mtx[g][q];
    int cap = PR.mil > 0 ? PR.mil : 0;
    for (int j = 0; j + 1 < PR.nquant; ++j) cap = cap > PR.quant[j] ? cap : PR.quant[j];
    cap += 1;
    std::vector<int2> pen(cap + 1);
    std::vector<PkPen> ppen(cap + 1);
    for (int h = 0; h <= cap; ++h) {
        int pv = PR.mean[0];
        for (int j = 1; j < PR.nquant; ++j) if (h > PR.quant[j - 1]) pv = PR.mean[j];
        const bool valid = h > PR.mil;
        pen[h] = make_int2(valid ? pv : PEN_INVALID, valid ? -32768 : NEV);
        ppen[h].pc = pk_mk(valid ? pv : 0, -32768 - (valid ? pv : 0));
        ppen[h].valid = valid ? 0xffffu : 0u;
    }
    std::vector<uint2> t4(PK_T4);
    struct Mfun {
        const int* t; const int* idx;
        __host__ __device__ int operator()(int cc, int ac) const { return t[idx[cc] * MTX_LD + idx[ac]]; }
    } mfun{mtxT, idx_of};
    for (int i = 0; i < PK_T4; ++i) pk_t4_entry(t4[i], i, mfun);
    PkConst K;
    K.gn = pk_dup(PR.gn); K.ge = pk_dup(PR.ge); K.cgn = pk_dup(-32768 - PR.gn); K.nev = pk_dup(NEV);
    K.one = 0x00010001u; K.eight = 0x00080008u; K.cap8 = pk_dup(8 * cap);

    // ---- scalar state (one thread, strip rows 0..7)
    int HA[NR], HB[NR], F[NR], E[NR], V2[NR], NJ[NR], arow[NR], F2[NR], E2[NR];
    int acls[NR];
    for (int k = 0; k < NR; ++k) {
        HA[k] = HB[k] = F[k] = E[k] = V2[k] = NEV; NJ[k] = 0; F2[k] = E2[k] = NEV;
        acls[k] = U(0, 9) == 0 ? (U(0, 1) ? 4 : 5) : U(0, 3);
        arow[k] = 4 * idx_of[acls[k]];
    }
    std::vector<RingEntry> ring(RING * CTA_THREADS);
    // ---- packed state
    unsigned pHA[NP], pHB[NP], pHG[NP], pFt[NP], pEt[NP], pV2[NP], pHL[NP], parow[NP];
    for (int j = 0; j < NP; ++j) {
        pHA[j] = pHB[j] = pV2[j] = pk_dup(NEV);
        pHG[j] = pk_max(pk_dup(NEV), K.cgn);
        pFt[j] = pEt[j] = pk_dup(NEV - PR.gn);
        pHL[j] = 0;
        parow[j] = (unsigned) ((acls[j] * PK_NC + acls[j + NP]) * 8);
    }
    PkRingA ringA[2 * NP];
    PkRingB ringB[2 * NP];
    for (int s = 0; s < 2 * NP; ++s) { ringA[s] = PkRingA{0, 0, pk_dup(-32768), 0}; ringB[s] = PkRingB{pk_dup(-32768), PK_ZC}; }

    // level around which the incoming row lives
    const int base = regime == 0 ? U(-3000, 20000) : regime == 1 ? U(-32768, -31000) : U(24000, 32000);
    const int n0 = U(0, 1000);
    // pre-fill: columns left of the first one pair residues but carry no signal
    std::vector<int> ccls(steps + 16), cs3(steps + 16), cs5(steps + 16);
    for (int d = 15; d >= 1; --d) {
        const int c = n0 - d;
        const int cls = U(0, 11) == 0 ? (U(0, 1) ? 4 : 5) : U(0, 3);
        RingEntry re; re.pad = 0; re.prof = idx_of[cls] * (MTX_LD * 4); re.s3 = 0; re.s5 = 0;
        ring[(c & 15) * CTA_THREADS] = re; ring[((c & 15) + 16) * CTA_THREADS] = re;
        if (d <= 2 * NP - 1) pk_ring_push<NP>(ringA, ringB, 1, c, cls, 0, 0);
    }
    int prev_uh = NEV;
    unsigned prev_in = pk_dup(NEV);
    int hmax_true = -32768, s3max = 0, s5max = 0, pvmax = 0;
    for (int g = 0; g < 6; ++g) for (int q = 0; q < 6; ++q) if (g < 5 && q < 5 && PR.mtx[g][q] > pvmax) pvmax = PR.mtx[g][q];
    unsigned hmax_pk = pk_dup(-32768);
    long bad = 0;
    for (int j = 0; j < steps; ++j) {
        const int n = n0 + j;
        // inputs of this step
        const int cls = U(0, 11) == 0 ? (U(0, 1) ? 4 : 5) : U(0, 3);
        int s3 = U(0, 5) == 0 ? U(-PK_SIGMAX, 120) : U(-400, 60);
        int s5 = U(0, 5) == 0 ? U(-PK_SIGMAX, 120) : U(-400, 60);
        if (U(0, 40) == 0) { s3 = 100; s5 = 100 - PR.ipen; }
        const int s5i = (int) (short) (s5 + PR.ipen);
        if (s3 > s3max) s3max = s3;
        if (s5i > s5max) s5max = s5i;
        int up_h = base + U(-600, 600), up_f = up_h + PR.gn - U(0, 300);
        if (U(0, 30) == 0) up_f = NEV;
        if (regime == 1 && U(0, 3) == 0) { up_h = -32768; up_f = -32768; }
        up_h = sat16(up_h); up_f = sat16(up_f);
        if (up_f > up_h + PR.gn) up_f = sat16(up_h + PR.gn);        // F <= H + gn always holds
        // ---- scalar
        RingEntry re; re.pad = 0; re.prof = idx_of[cls] * (MTX_LD * 4); re.s3 = SPJ ? s3 : 0; re.s5 = SPJ ? s5i : 0;
        const int rslot = n & 15;
        ring[rslot * CTA_THREADS] = re; ring[(rslot + 16) * CTA_THREADS] = re;
        const char* ring_hi = reinterpret_cast<const char*>(ring.data() + (rslot + 16) * CTA_THREADS);
        unsigned tw[NR / 4];
        int sv = INT_MIN, sk = 0;
        int (&HOs)[NR] = (j & 1) ? HB : HA;
        int (&HNs)[NR] = (j & 1) ? HA : HB;
        strip_step<NR, TRACE, false, SPJ, false>(HOs, HNs, F, E, F2, E2, NEV, 0, 0, V2, NJ, arow, ring_hi,
                                             reinterpret_cast<const char*>(mtxT), pen.data(), cap, j, up_h, up_f,
                                             prev_uh, PR.gn, PR.ge, INT_MIN, tw, sv, sk);
        prev_uh = up_h;
        for (int k = 0; k < NR; ++k) if (HOs[k] > hmax_true) hmax_true = HOs[k];
        // ---- packed
        pk_ring_push<NP>(ringA, ringB, 1, n, cls, SPJ ? s3 : 0, SPJ ? s5i : 0);
        unsigned (&HOp)[NP] = (j & 1) ? pHB : pHA;
        unsigned (&HNp)[NP] = (j & 1) ? pHA : pHB;
        const unsigned in_h = pk_mk(up_h, 0), in_f = pk_mk(up_f - PR.gn, 0);
        const unsigned uh0 = pk_perm(in_h, HNp[NP - 1], 0x5410);
        const unsigned uft0 = pk_perm(in_f, pFt[NP - 1], 0x5410);
        const unsigned dg0 = pk_perm(prev_in, HOp[NP - 1], 0x5410);
        prev_in = in_h;
        unsigned ptw[NP / 2] = {0};
        const int slot = (n & (NP - 1)) + NP;
        strip_step_pk<NP, TRACE, SPJ>(HOp, HNp, pHG, pFt, pEt, pV2, pHL, parow,
                                  reinterpret_cast<const char*>(ringA + slot), reinterpret_cast<const char*>(ringB + slot),
                                  (int) sizeof(PkRingA), (int) sizeof(PkRingB), reinterpret_cast<const char*>(t4.data()),
                                  reinterpret_cast<const char*>(ppen.data()), uh0, uft0, dg0, K, ptw, hmax_pk);
        // ---- compare
        bool diff = false;
        for (int k = 0; k < NR; ++k) {
            const int jj = k & (NP - 1);
            auto half = [&](unsigned w) { return k < NP ? pk_lo(w) : pk_hi(w); };
            const int ph = half(HOp[jj]), pf = (short) (half(pFt[jj]) + PR.gn), pe = (short) (half(pEt[jj]) + PR.gn);
            if (ph != HOs[k] || pf != F[k] || pe != E[k]) diff = true;
            if (SPJ) {
                const int hil_s = (j + 1 + NJ[k]) < cap ? (j + 1 + NJ[k]) : cap;    // counter as the NEXT step sees it
                const int hil_p = (int) ((k < NP ? pHL[jj] & 0xffffu : pHL[jj] >> 16) / 8);
                if (half(pV2[jj]) != V2[k] || hil_s != hil_p) diff = true;
            }
            if (TRACE) {
                const unsigned sc = (tw[k >> 2] >> (8 * (k & 3))) & 0xffu;
                const unsigned word = ptw[pk_trace_byte<NP>(k) >> 2];
                const unsigned pc = pk_trace_code((word >> (8 * (pk_trace_byte<NP>(k) & 3))) & 0xffu);
                if (sc != pc) diff = true;
            }
        }
        ++cmp;
        const int hm = pk_lo(hmax_pk) > pk_hi(hmax_pk) ? pk_lo(hmax_pk) : pk_hi(hmax_pk);
        const int hmon = hm > hmax_true ? hm : hmax_true;       // (a wrapped packed value can only raise it)
        const bool tripped = hmon + pvmax > 32767 || hmon + s5max > 32767 || hmon + s5max + s3max > 32767;
        if (tripped) { ++trips; return bad; }   // the kernel would hand this problem to the exact path
        if (diff) {
            if (bad < 5) fprintf(stderr, "MISMATCH regime %d step %d (hmax %d)\n", regime, j, hmon);
            return bad + 1;
        }
    }
    return bad;
}

int main(int argc, char** argv)
{
    const int runs = argc > 1 ? atoi(argv[1]) : 3000;
    const unsigned seed = argc > 2 ? (unsigned) atoi(argv[2]) : 1u;
    std::mt19937 rng(seed);
    long bad = 0, trips = 0, cmp = 0;
    for (int r = 0; r < runs; ++r) {
        Params P;
        auto U = [&](int lo, int hi) { return (int) (lo + (long long) (rng() % (unsigned) (hi - lo + 1))); };
        const int gep = -U(1, 60), gop = -U(0, 200);
        P.ge = gep; P.gn = gep + gop;
        P.ipen = -U(0, 600);
        P.nquant = U(1, 6);
        P.mil = U(0, 40);
        int q = P.mil + U(1, 30);
        for (int j = 0; j < 8; ++j) { P.quant[j] = q; q += U(1, 60); P.mean[j] = -U(0, 500); }
        const int match = U(5, 60), mis = -U(5, 120), nn = -U(0, 60);
        for (int g = 0; g < 6; ++g)
            for (int a = 0; a < 6; ++a)
                P.mtx[g][a] = (g == 5 || a == 5) ? 0 : (g == 4 || a == 4) ? nn : (g == a ? match : mis);
        const int regime = r % 3;
        const int steps = U(20, 400);
        switch (r % 4) {
        case 0: bad += run_case<true, true>(rng, P, regime, steps, trips, cmp); break;
        case 1: bad += run_case<false, true>(rng, P, regime, steps, trips, cmp); break;
        case 2: bad += run_case<true, false>(rng, P, regime, steps, trips, cmp); break;
        default: bad += run_case<false, false>(rng, P, regime, steps, trips, cmp); break;
        }
    }
    printf("check_packed (NP = %d): %d runs, %ld steps compared, %ld runs ended by the high-side monitor, %ld mismatches\n",
           NP, runs, cmp, trips, bad);
    return bad ? 1 : 0;
}
