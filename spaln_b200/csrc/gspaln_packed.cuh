// gspaln_packed.cuh -- the int16x2 form of the DNA `_wip` cell update (two query rows per register).
//
// Same semantics as strip_step() in gspaln_kernels.cuh (SimdAln2s1::forwardS1_wip /
// scoreonlyS1_wip, src/fwd2s1_wip_simd.h:42-474, single affine, global / semi-global), evaluated
// with the packed DPX instructions of sm_100a (VIADDMNMX.S16x2, VIMNMX.S16x2, VIMNMX3.S16x2,
// VIADD.16x2, VIMNMX.U16x2): a thread owns 8 rows of a strip as 4 registers, register j holding
// rows j (low half) and j + 4 (high half), so that the "row above" of a register is simply the
// register before it.
//
// The hardware adds wrap, the reference's saturate.  Exactness is kept as follows:
//  * low side, where saturation is routine (cells no path reaches drift down to -32768): every
//    add is preceded by a max with the constant that makes it exact,
//        satlo(a + g) == max(a, -32768 - g) + g           (g <= 0),
//    and the two gap states are carried minus the open+extend constant (Et = E - gn, Ft = F - gn),
//    which turns "max(satlo(E + ge), satlo(H + gn))" into ONE add-max against HG = max(H, -32768 - gn)
//    and the reference's "extend only if strictly better" flag into Et' != HG;
//  * high side, where the reference's re-basing keeps values away from +32767: not clamped, but
//    MONITORED -- every H is folded into a running maximum and the problem is handed to the exact
//    32-bit kernel if  max H + (largest positive addend)  could have passed 32767.  By induction the
//    first wrapping add would need an input that the monitor has already seen.
//  * direction flags come out of the packed maxima as (new XOR old) != 0 and are stored as raw bits
//    (the walk decodes them into the reference's TraceBackCode).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gspaln {

#if defined(__CUDACC__)
#define GSPALN_HD __host__ __device__ __forceinline__
#else
#define GSPALN_HD inline
#endif
// portable per-half definitions (host side of the checker tools); the device uses the DPX forms
GSPALN_HD int pk_lo(unsigned w) { return (int) (short) (w & 0xffffu); }
GSPALN_HD int pk_hi(unsigned w) { return (int) (short) (w >> 16); }
GSPALN_HD unsigned pk_mk(int lo, int hi) { return ((unsigned) lo & 0xffffu) | ((unsigned) hi << 16); }
GSPALN_HD unsigned pk_max(unsigned a, unsigned b)
{
#if defined(__CUDA_ARCH__)
    return __vmaxs2(a, b);
#else
    return pk_mk(pk_lo(a) > pk_lo(b) ? pk_lo(a) : pk_lo(b), pk_hi(a) > pk_hi(b) ? pk_hi(a) : pk_hi(b));
#endif
}
GSPALN_HD unsigned pk_max3(unsigned a, unsigned b, unsigned c)
{
#if defined(__CUDA_ARCH__)
    return __vimax3_s16x2(a, b, c);
#else
    return pk_max(pk_max(a, b), c);
#endif
}
GSPALN_HD unsigned pk_add(unsigned a, unsigned b)
{
#if defined(__CUDA_ARCH__)
    return __vadd2(a, b);
#else
    return pk_mk(pk_lo(a) + pk_lo(b), pk_hi(a) + pk_hi(b));
#endif
}
GSPALN_HD unsigned pk_addmax(unsigned a, unsigned b, unsigned c)
{
#if defined(__CUDA_ARCH__)
    return __viaddmax_s16x2(a, b, c);
#else
    return pk_max(pk_add(a, b), c);
#endif
}
GSPALN_HD unsigned pk_addmin(unsigned a, unsigned b, unsigned c)
{
#if defined(__CUDA_ARCH__)
    return __viaddmin_s16x2(a, b, c);
#else
    const unsigned s = pk_add(a, b);
    return pk_mk(pk_lo(s) < pk_lo(c) ? pk_lo(s) : pk_lo(c), pk_hi(s) < pk_hi(c) ? pk_hi(s) : pk_hi(c));
#endif
}
GSPALN_HD unsigned pk_minu(unsigned a, unsigned b)
{
#if defined(__CUDA_ARCH__)
    return __vminu2(a, b);
#else
    const unsigned al = a & 0xffffu, bl = b & 0xffffu, ah = a >> 16, bh = b >> 16;
    return (al < bl ? al : bl) | ((ah < bh ? ah : bh) << 16);
#endif
}
GSPALN_HD unsigned pk_perm(unsigned a, unsigned b, unsigned s)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, s);
#else
    const unsigned long long v = ((unsigned long long) b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned) ((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
#endif
}

GSPALN_HD unsigned pk_dup(int x) { return ((unsigned) x & 0xffffu) * 0x10001u; }

constexpr int PK_NC = 6;                // residue classes of the pair table: A, C, G, T, N, zero row
constexpr int PK_T4 = PK_NC * PK_NC * PK_NC * PK_NC;    // entries of the pair table (8 B each)
constexpr int PK_ZC = 5;                // the zero class
constexpr int PK_SIGMAX = 8192;         // |signal| bound of the packed path (checked while packing)

// one column pair as the rows consume it: column c in the low halves, column c - NP in the high ones
// (NP = registers per thread: 4 -> 8 rows and two threads per strip, 8 -> 16 rows, one thread)
struct __align__(16) PkRingA {
    unsigned t4;                        // byte offset of the pair-table block of (code[c], code[c - 4])
    unsigned s3;                        // acceptor signals
    unsigned c3;                        // -32768 - min(s3, 0): makes the add of s3 exact on the low side
    unsigned s5;                        // donor signals + mean intron penalty
};
struct __align__(8) PkRingB {
    unsigned c5;                        // -32768 - min(s5, 0)
    unsigned code;                      // class of column c alone (the next pair's high half)
};

struct PkConst {
    unsigned gn, ge;                    // open + extend, extend (both halves)
    unsigned cgn;                       // -32768 - gn
    unsigned nev;                       // nevsel
    unsigned one, eight, cap8;          // 1 | 1 << 16, 8 | 8 << 16, 8 * (pen table size - 1)
};

// intron-length penalty table entry, indexed by 8 * length counter:
//   .x = penalty | (-32768 - penalty) << 16,  .y = 0xffff if the length is allowed else 0
struct PkPen { unsigned pc, valid; };

// ---------------------------------------------------------------------------
// one step of one thread: 8 independent cell updates as 4 packed registers.
//   HN: H of the previous step, HO: H of two steps ago (overwritten with the new H)
//   HG: max(HN, cgn) per register (overwritten with the same of the new H)
//   Ft, Et: vertical / horizontal gap state minus gn;  V2: best donor value;  HL: 8 x intron length
//   uh0 / uft0 / dg0: the "row above" inputs of register 0 -- low half from the neighbour or the
//   band row, high half = row 3 of this thread (prepared by the caller BEFORE the step)
//   ra_hi / rb_hi: ring slot of column n (register j reads the slot j entries below)
//   tw: 8 trace bytes (TRACE); hmax: running maximum of every H (monitor)
// ---------------------------------------------------------------------------
template <int NP, bool TRACE, bool SPJ>
GSPALN_HD void strip_step_pk(unsigned (&HO)[NP], const unsigned (&HN)[NP], unsigned (&HG)[NP], unsigned (&Ft)[NP],
                             unsigned (&Et)[NP], unsigned (&V2)[NP], unsigned (&HL)[NP], const unsigned (&arow4)[NP],
                             const char* ra_hi, const char* rb_hi, int ra_stride, int rb_stride,
                             const char* t4_bytes, const char* pen_bytes, unsigned uh0, unsigned uft0,
                             unsigned dg0, const PkConst& K, unsigned (&tw)[NP / 2], unsigned& hmax)
{
    const unsigned hgup0 = pk_max(uh0, K.cgn);
    unsigned code[NP];
#pragma unroll
    for (int j = NP - 1; j >= 0; --j) {
        const PkRingA ra = *reinterpret_cast<const PkRingA*>(ra_hi - j * ra_stride);
        const unsigned left_hg = HG[j];
        const unsigned up_hg = j ? HG[j ? j - 1 : 0] : hgup0;
        const unsigned up_ft = j ? Ft[j ? j - 1 : 0] : uft0;
        const unsigned dg = j ? HO[j ? j - 1 : 0] : dg0;
        // horizontal gap: extend or open, whichever is better (ties open)
        const unsigned et = pk_addmax(Et[j], K.ge, left_hg);
        Et[j] = et;
        // vertical gap
        const unsigned ft = pk_addmax(up_ft, K.ge, up_hg);
        Ft[j] = ft;
        // diagonal: substitution scores of both rows in one look-up, low side exact
        const uint2 t4 = *reinterpret_cast<const uint2*>(t4_bytes + ra.t4 + arow4[j]);
        const unsigned h0 = pk_add(pk_max(dg, t4.y), t4.x);
        const unsigned h1 = pk_addmax(ft, K.gn, h0);
        const unsigned h2 = pk_addmax(et, K.gn, h1);
        unsigned h = h2;
        unsigned xa = 0, xd = 0;
        if (SPJ) {
            // acceptor: best donor of the row + 3' signal + binned length penalty
            const unsigned q0 = pk_add(pk_max(V2[j], ra.c3), ra.s3);
            const unsigned hl = HL[j];
            const PkPen p_lo = *reinterpret_cast<const PkPen*>(pen_bytes + (hl & 0xffffu));
            const PkPen p_hi = *reinterpret_cast<const PkPen*>(pen_bytes + (hl >> 16));
            const unsigned pen = pk_perm(p_lo.pc, p_hi.pc, 0x5410);
            const unsigned cpen = pk_perm(p_lo.pc, p_hi.pc, 0x7632);
            const unsigned vm = pk_perm(p_lo.valid, p_hi.valid, 0x5410);
            unsigned q = pk_add(pk_max(q0, cpen), pen);
            q = (q & vm) | (K.nev & ~vm);
            h = pk_max(h2, q);
            xa = h ^ h2;
            // donor: leaves from the cell's final H; never from a cell an intron just entered
            const PkRingB rb = *reinterpret_cast<const PkRingB*>(rb_hi - j * rb_stride);
            unsigned qd = pk_add(pk_max(h, rb.c5), ra.s5);
            if (TRACE) {
                const unsigned ma = pk_minu(xa, K.one) * 0xffffu;
                qd = (qd & ~ma) | (K.nev & ma);
            }
            const unsigned v2n = pk_max(V2[j], qd);
            xd = v2n ^ V2[j];
            V2[j] = v2n;
            const unsigned md = pk_minu(xd, K.one) * 0xffffu;
            HL[j] = pk_addmin(hl & ~md, K.eight, K.cap8);
        }
        HO[j] = h;
        HG[j] = pk_max(h, K.cgn);
        if (TRACE) {
            // raw decision bits: 1 vertical beat diagonal, 2 horizontal beat that, 4 acceptor beat
            // that, 8 horizontal gap extended, 16 vertical gap extended, 32 new best donor
            unsigned c = pk_minu(h1 ^ h0, K.one);
            c += 2u * pk_minu(h2 ^ h1, K.one);
            c += 8u * pk_minu(et ^ left_hg, K.one);
            c += 16u * pk_minu(ft ^ up_hg, K.one);
            if (SPJ) {
                c += 4u * pk_minu(xa, K.one);
                c += 32u * pk_minu(xd, K.one);
            }
            code[j] = c;
        }
    }
#pragma unroll
    for (int j = 0; j < NP; j += 2) hmax = pk_max3(hmax, HO[j], HO[j + 1]);
    if (TRACE) {
        // word w: bytes (row 2w, row 2w + 1, row 2w + NP, row 2w + 1 + NP)
#pragma unroll
        for (int w = 0; w < NP / 2; ++w) tw[w] = code[2 * w] | (code[2 * w + 1] << 8);
    }
}

// Column c enters a thread's ring: the new pair entry takes column c in its low halves and column
// c - NP -- the low halves of the entry it replaces (slot c mod NP) -- in its high halves.  The ring
// is doubled (slots s and s + NP hold the same entry) so that register j reads column c - j at a
// constant offset below slot (c mod NP) + NP.
template <int NP>
GSPALN_HD void pk_ring_push(PkRingA* ringA, PkRingB* ringB, int stride, int c, int cls, int s3, int s5)
{
    const int slot = c & (NP - 1);
    const PkRingA oa = ringA[slot * stride];
    const PkRingB ob = ringB[slot * stride];
    PkRingA na;
    PkRingB nb;
    const unsigned c3 = (unsigned) (-32768 - (s3 < 0 ? s3 : 0));
    const unsigned c5 = (unsigned) (-32768 - (s5 < 0 ? s5 : 0));
    na.t4 = (unsigned) ((cls * PK_NC + (int) ob.code) * (PK_NC * PK_NC * 8));
    na.s3 = pk_perm((unsigned) s3, oa.s3, 0x5410);
    na.c3 = pk_perm(c3, oa.c3, 0x5410);
    na.s5 = pk_perm((unsigned) s5, oa.s5, 0x5410);
    nb.c5 = pk_perm(c5, ob.c5, 0x5410);
    nb.code = (unsigned) cls;
    ringA[slot * stride] = na; ringA[(slot + NP) * stride] = na;
    ringB[slot * stride] = nb; ringB[(slot + NP) * stride] = nb;
}

// pair table: entry ((cc * NC + cc4) * NC + ac) * NC + ac4 = {scores of (cc, ac) | (cc4, ac4) << 16,
// their low-side clamps}.  mtx(cc, ac): substitution score of genome class cc against query class ac
template <class MTX>
GSPALN_HD void pk_t4_entry(uint2& e, int idx, const MTX& mtx)
{
    const int ac4 = idx % PK_NC, ac = idx / PK_NC % PK_NC, cc4 = idx / (PK_NC * PK_NC) % PK_NC,
              cc = idx / (PK_NC * PK_NC * PK_NC);
    const int lo = mtx(cc, ac), hi = mtx(cc4, ac4);
    e.x = ((unsigned) lo & 0xffffu) | ((unsigned) hi << 16);
    e.y = ((unsigned) (-32768 - (lo < 0 ? lo : 0)) & 0xffffu) | ((unsigned) (-32768 - (hi < 0 ? hi : 0)) << 16);
}

// raw decision bits -> TraceBackCode (src/rhomb_coord.h:36-61)
GSPALN_HD unsigned pk_trace_code(unsigned c)
{
    unsigned pb = 1u;                               // DIAG
    if (c & 1u) pb = 8u;                            // VERT
    if (c & 2u) pb = 2u;                            // HORI
    if (c & 4u) pb = 14u;                           // ACCR
    return pb | ((c & 8u) ? 0u : 16u) | ((c & 16u) ? 0u : 32u) | ((c & 32u) ? 128u : 0u);
}

// byte of row k (0 .. 2 NP - 1 within the thread) inside the thread's 2 NP trace bytes
template <int NP>
GSPALN_HD int pk_trace_byte(int k)
{
    const int j = k & (NP - 1), half = k / NP;
    return 4 * (j >> 1) + (j & 1) + 2 * half;
}

}   // namespace gspaln
