// gspaln.cu -- host side of the C-ABI declared in include/gspaln.h.
//
// Owns the device pools (query codes, genome columns, band buffers, trace
// matrix, corner records), packs a batch of problems into them, launches the
// persistent DP kernels on the engine's own stream and times every phase with
// CUDA events on that stream.  There is no CPU implementation behind this
// API: without a CUDA device gspaln_create() fails with GSPALN_ENODEV.
#include "../../include/gspaln.h"
#include "gspaln_kernels.cuh"
#include "gspaln_udh.cuh"
#include "gspaln_ng.cuh"
#include "gspaln_xudh.cuh"
#include "gspaln_host.hpp"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

using namespace gspaln;

struct gspaln_ctx {
    int device = 0;
    int sm_count = 0;
    gspaln_params prm;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;    // H2D of the chunks that follow the first one
    cudaEvent_t ev_sync[2] = {nullptr, nullptr};
    PinBuf<int> h_marks;                    // watermark values, one per chunk
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    DevBuf<DevParams> d_prm;
    DevBuf<int2> d_pen;
    int pen_cap = 0;
    size_t smem_bytes = 0;
    unsigned char perm[32];
    DevBuf<DevTask> d_tasks;
    DevBuf<int> d_order;
    DevBuf<int> d_ticket;
    DevBuf<unsigned char> d_apool;
    DevBuf<ColInfo> d_cpool;
    DevBuf<unsigned> d_band;
    DevBuf<unsigned char> d_trace;
    DevBuf<int2> d_skl;
    DevBuf<DevResult> d_res;
    DevBuf<int> d_udh;              // per-warp UDH workspace (link band + intermediates)
    DevBuf<int> d_cpos;
    DevBuf<DevUdhOut> d_ures;
    // scalar exact-ILD kernel (GSPALN_FORWARD_NG)
    DevParams hP;                   // host copy of the device parameter block
    DevBuf<short> d_ngtab;          // sig53tab[544] | Penalty(0 .. n_pen - 1)
    DevBuf<unsigned char> d_ngwork; // per-thread workspaces
    int n_pen = 0;
    bool ng_ready = false;
    int n_ng = 0, grid_run_ng = 0, ng_rec_cap = 0;
    size_t ng_slab = 0, ng_width = 0;
    DevBuf<int> d_ngs;              // score-only scalar kernel: three int rows per thread
    int n_ngs = 0, grid_run_ngs = 0;
    size_t ngs_width = 0;
    bool ng_full_records = false;       // second run of the trace-backs whose record store overflowed
    int ng_rec_eighths = 4;             // records per cell of the first run, in eighths (GSPALN_NG_REC_EIGHTHS)
    int n_xudh = 0, grid_run_xudh = 0;  // scalar Hirschberg pass (GSPALN_HIRSCHBERG_NG): one warp per problem
    int n_xudh_w = 0, grid_run_xudh_w = 0;  // ... queries of XUDH_WIDE_ROWS rows and more: XUDH_WIDE warps per problem
    int n_ng_w = 0, grid_run_ng_w = 0, n_ngs_w = 0, grid_run_ngs_w = 0;    // exact-ILD kernels, wide class (NG_WIDE warps per problem)
    size_t xudh_width = 0;
    PinBuf<DevTask> h_tasks;
    PinBuf<int> h_order;
    PinBuf<unsigned char> h_apool;
    PinBuf<ColInfo> h_cpool;
    PinBuf<int2> h_skl;
    PinBuf<DevResult> h_res;
    PinBuf<int> h_cpos;
    PinBuf<DevUdhOut> h_ures;
    // resident batch
    int n = 0;
    int n_trace = 0, n_score = 0, n_udh = 0;
    size_t a_bytes = 0, c_elems = 0, band_slab = 0, trace_slab = 0, skl_elems = 0;
    size_t udh_slab = 0, cpos_elems = 0;
    int grid_run_trace = 0, grid_run_score = 0, grid_run_udh = 0, grid_udh = 0;
    int n_udh_t = 0, grid_run_udh_t = 0, grid_udh_t = 0;   // Hirschberg passes of the team class (a CTA per problem)
    std::vector<int64_t> cells;
    std::vector<int> skl_cap;
    gspaln_timing tim;
    std::string err;
    int grid_trace = 0, grid_score = 0;
    // narrower systolic chains for problems that cannot fill 16 strip slots: class c runs with
    // NRT = 4 >> c strip rows per thread, i.e. 8 / 4 / 2 strip slots per warp (gspaln_kernels.cuh)
    int n_trace_c[3] = {0, 0, 0}, n_score_c[3] = {0, 0, 0};
    size_t band_slab_c[3] = {0, 0, 0}, trace_slab_c[3] = {0, 0, 0};
    int grid_trace_c[3] = {0, 0, 0}, grid_score_c[3] = {0, 0, 0};
    int grid_run_trace_c[3] = {0, 0, 0}, grid_run_score_c[3] = {0, 0, 0};
    // packed int16x2 kernels (gspaln_packed.cuh)
    bool pk_ok = false;
    int pk_np = 4;                  // packed registers per thread of the current batch (4 | 8)
    int pk_np_forced = 0;           // GSPALN_PK_NP
    size_t smem_pk[2] = {0, 0};     // [np == 8]
    int grid_trace_pk[2] = {0, 0}, grid_score_pk[2] = {0, 0};
};

namespace {

int fail(gspaln_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess)
{
    if (c) {
        c->err = what;
        if (e != cudaSuccess) { c->err += ": "; c->err += cudaGetErrorString(e); }
    }
    return code;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, GSPALN_ECUDA, #call, e_); } while (0)

using KernelFn = void (*)(const DevParams*, const int2*, const DevTask*, const int*, int, int*,
                          const unsigned char*, const ColInfo*, unsigned*, long long,
                          unsigned char*, long long, int2*, DevResult*, const int*);

KernelFn kernel_fn(bool trace, bool local, bool spj, bool dagp = false)
{
    static const KernelFn tab[16] = {
        dp_wip_kernel<false, false, false, false>, dp_wip_kernel<false, false, true, false>,
        dp_wip_kernel<false, true, false, false>, dp_wip_kernel<false, true, true, false>,
        dp_wip_kernel<true, false, false, false>, dp_wip_kernel<true, false, true, false>,
        dp_wip_kernel<true, true, false, false>, dp_wip_kernel<true, true, true, false>,
        dp_wip_kernel<false, false, false, true>, dp_wip_kernel<false, false, true, true>,
        dp_wip_kernel<false, true, false, true>, dp_wip_kernel<false, true, true, true>,
        dp_wip_kernel<true, false, false, true>, dp_wip_kernel<true, false, true, true>,
        dp_wip_kernel<true, true, false, true>, dp_wip_kernel<true, true, true, true>,
    };
    return tab[(dagp ? 8 : 0) | (trace ? 4 : 0) | (local ? 2 : 0) | (spj ? 1 : 0)];
}

const void* kernel_ptr(bool trace, bool local, bool spj, bool dagp = false)
{
    return reinterpret_cast<const void*>(kernel_fn(trace, local, spj, dagp));
}

// 32-bit kernels with 4 / 2 / 1 strip rows per thread (8 / 4 / 2 strip slots per warp), single affine
KernelFn thin_kernel_fn(bool trace, bool local, bool spj, int nrt)
{
    static const KernelFn tab[24] = {
#define GSPALN_THIN(N) \
        dp_wip_kernel<false, false, false, false, 0, N>, dp_wip_kernel<false, false, true, false, 0, N>, \
        dp_wip_kernel<false, true, false, false, 0, N>, dp_wip_kernel<false, true, true, false, 0, N>,   \
        dp_wip_kernel<true, false, false, false, 0, N>, dp_wip_kernel<true, false, true, false, 0, N>,   \
        dp_wip_kernel<true, true, false, false, 0, N>, dp_wip_kernel<true, true, true, false, 0, N>
        GSPALN_THIN(4), GSPALN_THIN(2), GSPALN_THIN(1)
#undef GSPALN_THIN
    };
    const int c = nrt == 4 ? 0 : (nrt == 2 ? 1 : 2);
    return tab[8 * c + ((trace ? 4 : 0) | (local ? 2 : 0) | (spj ? 1 : 0))];
}

// Chain width of a forward / score-only problem: a strip can start ~33 steps after the strip above
// it and lives for width + 31 steps, so at most (width + 31) / 33 + 1 strips of a problem are in
// flight, and never more than it has.  The class is the widest chain (2 x rows-per-thread strip
// slots) the problem can keep full: all 32 lanes of its warp stay busy.
int wip_class(int rows, int width, bool dagp)
{
    if (dagp || getenv("GSPALN_NO_THIN")) return 8;
    const int nstr = (rows + NELEM - 1) / NELEM;
    const int need = std::min(nstr, (width + 31) / 33 + 1);
    return need >= 16 ? 8 : (need >= 8 ? 4 : (need >= 4 ? 2 : 1));
}

// the packed int16x2 kernels: single affine, global / semi-global; np = packed registers per
// thread (4: two threads per strip, 16 strip slots per warp; 8: one thread per strip, 32 slots)
KernelFn pk_kernel_fn(bool trace, bool spj, int np)
{
    static const KernelFn tab[8] = {
        dp_wip_kernel<false, false, false, false, 4>, dp_wip_kernel<false, false, true, false, 4>,
        dp_wip_kernel<true, false, false, false, 4>, dp_wip_kernel<true, false, true, false, 4>,
        dp_wip_kernel<false, false, false, false, 8>, dp_wip_kernel<false, false, true, false, 8>,
        dp_wip_kernel<true, false, false, false, 8>, dp_wip_kernel<true, false, true, false, 8>,
    };
    return tab[(np == 8 ? 4 : 0) | (trace ? 2 : 0) | (spj ? 1 : 0)];
}

using UdhKernelFn = void (*)(const DevParams*, const int2*, const DevTask*, const int*, int, int*,
                             const unsigned char*, const ColInfo*, unsigned*, long long, int*,
                             long long, int*, DevUdhOut*, const int*);

UdhKernelFn udh_kernel_fn(bool spj, bool local)
{
    static const UdhKernelFn tab[4] = {dp_udh_kernel<false, false>, dp_udh_kernel<true, false>,
                                       dp_udh_kernel<false, true>, dp_udh_kernel<true, true>};
    return tab[(local ? 2 : 0) | (spj ? 1 : 0)];
}

// the team class: one CTA per problem (queries of >= UDH_TEAM_ROWS rows)
UdhKernelFn udh_team_kernel_fn(bool spj, bool local)
{
    static const UdhKernelFn tab[4] = {dp_udh_kernel<false, false, WARPS_PER_CTA>, dp_udh_kernel<true, false, WARPS_PER_CTA>,
                                       dp_udh_kernel<false, true, WARPS_PER_CTA>, dp_udh_kernel<true, true, WARPS_PER_CTA>};
    return tab[(local ? 2 : 0) | (spj ? 1 : 0)];
}

const void* udh_kernel_ptr(bool spj, bool local)
{
    return reinterpret_cast<const void*>(udh_kernel_fn(spj, local));
}

int64_t task_cells(const gspaln_task& t)
{
    // rows m in (a_left, a_right], columns max(m + lw, b_left) < n <= min(m + up + 1, b_right)
    return band_cells(t.a_left, t.a_right, 1, t.lw, t.b_left, t.up + 1, t.b_right);
}

}   // namespace

extern "C" {

const char* gspaln_version(void) { return "gspaln 0.1 (sm_100a)"; }

int gspaln_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int64_t gspaln_task_cells(const gspaln_task* t) { return t ? task_cells(*t) : 0; }

const char* gspaln_last_error(const gspaln_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int gspaln_create(gspaln_ctx** out, const gspaln_params* prm, int device)
{
    if (!out || !prm) return GSPALN_EINVAL;
    *out = nullptr;
    if ((prm->noll != 2 && prm->noll != 3) || prm->simdim <= 0 || prm->simdim >= ZROW ||
        (prm->noll == 3 && ((short) prm->lgep > 0 || (short) (prm->lgep + prm->lgop) > 0)) ||
        prm->nquant < 1 || prm->nquant > GSPALN_MAXQUANT || prm->avmch <= 0 ||
        (short) prm->gep > 0 || (short) (prm->gep + prm->gop) > 0)     // kernels clamp gap terms on the low side only
        return GSPALN_EINVAL;
    int ndev = gspaln_device_count();
    if (ndev <= 0 || device < 0 || device >= ndev) return GSPALN_ENODEV;
    gspaln_ctx* ctx = new gspaln_ctx;
    ctx->device = device;
    ctx->prm = *prm;
    memset(&ctx->tim, 0, sizeof(ctx->tim));
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ctx->ev_sync[i], cudaEventDisableTiming);
    for (int i = 0; i < 6 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { gspaln_destroy(ctx); return GSPALN_ECUDA; }
    ctx->sm_count = prop.multiProcessorCount;

    DevParams P;
    memset(&P, 0, sizeof(P));
    P.ge = (short) prm->gep;                        // Splat() narrows to short
    P.gn = (short) (prm->gep + prm->gop);
    P.ge2 = (short) prm->lgep;                      // double affine (Noll == 3)
    P.gn2 = (short) (prm->lgep + prm->lgop);
    P.ipen = (short) (prm->spj ? prm->ipen : NEV);
    P.mil = (short) prm->llmt;
    P.nquant = prm->nquant;
    for (int j = 0; j < prm->nquant; ++j) {
        P.quant[j] = (short) prm->quant_len[j];
        P.mean[j] = (short) prm->quant_pen[j];
    }
    P.avmch = prm->avmch; P.local = prm->local ? 1 : 0; P.spj = prm->spj ? 1 : 0;
    P.lgop = prm->lgop; P.lgep = prm->lgep; P.noll = prm->noll; P.llmt = prm->llmt;
    P.codonk1 = INT_MAX;            // set by gspaln_set_ng_tables
    P.simdim = prm->simdim; P.gappen1 = prm->gappen1; P.gop = prm->gop; P.gep = prm->gep;
    // residue code -> table index.  The four unambiguous nucleotides of the
    // reference's DNA alphabet (A=2, C=3, G=5, T=9) go first so that their 16
    // pairs fall into 16 distinct shared-memory banks (row stride MTX_LD).
    {
        int nxt = 0;
        bool used[32] = {false};
        if (prm->simdim == 17)
            for (int c : {2, 3, 5, 9}) { P.perm[c] = (unsigned char) nxt++; used[c] = true; }
        for (int c = 0; c < 32; ++c)
            if (!used[c]) P.perm[c] = (unsigned char) (c < prm->simdim ? nxt++ : ZROW);
    }
    for (int q = 0; q < prm->simdim; ++q)
        for (int g = 0; g < prm->simdim; ++g)
            P.mtxT[P.perm[g] * MTX_LD + P.perm[q]] = (short) prm->simmtx[q * prm->simdim + g];
    memcpy(ctx->perm, P.perm, 32);
    // binned intron-length penalty as a table over the (saturating) length
    // counter: src/fwd2s1_wip_simd.h:389-396.  Entry h: {penalty, lower clamp};
    // lengths <= llmt give exactly nevsel.
    int cap = std::max(0, P.mil);
    for (int j = 0; j + 1 < P.nquant; ++j) cap = std::max(cap, P.quant[j]);
    cap += 1;
    std::vector<int2> pen(cap + 1);
    for (int h = 0; h <= cap; ++h) {
        int pv = P.mean[0];
        for (int j = 1; j < P.nquant; ++j) if (h > P.quant[j - 1]) pv = P.mean[j];
        const bool valid = h > P.mil;
        pen[h] = make_int2(valid ? pv : PEN_INVALID, valid ? -32768 : NEV);
    }
    P.pen_cap = cap;
    ctx->pen_cap = cap;
    // May the packed int16x2 kernel run problems of this parameter set?  It clamps the low side
    // only (gap and intron penalties must be <= 0), keeps 8 x the intron-length counter in 16 bits,
    // and knows the residue classes A, C, G, T, N of the DNA alphabet.
    {
        bool ok = prm->simdim == 17 && prm->noll == 2 && !prm->local && prm->gop <= 0 && cap <= 4000 &&
                  getenv("GSPALN_NO_PACKED") == nullptr;
        for (int j = 0; j < P.nquant; ++j) ok = ok && P.mean[j] <= 0 && P.mean[j] >= -PK_SIGMAX;
        if (prm->spj) ok = ok && P.ipen <= PK_SIGMAX && P.ipen >= -PK_SIGMAX;
        int pvmax = 0;
        for (int q = 0; q < prm->simdim; ++q)
            for (int g = 0; g < prm->simdim; ++g) {
                const int v = (short) prm->simmtx[q * prm->simdim + g];
                ok = ok && v >= -1024 && v <= 1024;
                pvmax = std::max(pvmax, v);
            }
        P.pk_ok = ok ? 1 : 0;
        P.pk_pvmax = pvmax;
        P.pk_nidx = P.perm[16];
        ctx->pk_ok = ok;
    }
    ctx->hP = P;
    ctx->smem_bytes = sizeof(RingEntry) * RING * CTA_THREADS + sizeof(int2) * (size_t) (cap + 1);
    if (ctx->smem_bytes > 108 * 1024) {     // two CTAs per SM must fit
        gspaln_destroy(ctx);
        return GSPALN_EINVAL;
    }
    if (ctx->d_prm.reserve(1) != cudaSuccess || ctx->d_ticket.reserve(64) != cudaSuccess ||
        ctx->d_pen.reserve(cap + 1) != cudaSuccess ||
        cudaMemcpy(ctx->d_pen.p, pen.data(), sizeof(int2) * (cap + 1), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(ctx->d_prm.p, &P, sizeof(P), cudaMemcpyHostToDevice) != cudaSuccess) {
        gspaln_destroy(ctx);
        return GSPALN_ENOMEM;
    }
    const bool dagp = prm->noll == 3;
    const void* kt = kernel_ptr(true, P.local, P.spj, dagp);
    const void* ks = kernel_ptr(false, P.local, P.spj, dagp);
    cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ctx->smem_bytes);
    cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ctx->smem_bytes);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kt, CTA_THREADS, ctx->smem_bytes);
    ctx->grid_trace = std::max(1, occ) * ctx->sm_count;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ks, CTA_THREADS, ctx->smem_bytes);
    ctx->grid_score = std::max(1, occ) * ctx->sm_count;
    {
        const void* ku = udh_kernel_ptr(P.spj, P.local);
        cudaFuncSetAttribute(ku, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ctx->smem_bytes);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ku, CTA_THREADS, ctx->smem_bytes);
        ctx->grid_udh = std::max(1, occ) * ctx->sm_count;
        const void* kt = reinterpret_cast<const void*>(udh_team_kernel_fn(P.spj, P.local));
        cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ctx->smem_bytes);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kt, CTA_THREADS, ctx->smem_bytes);
        ctx->grid_udh_t = std::max(1, occ) * ctx->sm_count;
    }
    if (!dagp)
        for (int c = 0; c < 3; ++c) {
            const void* kt2 = reinterpret_cast<const void*>(thin_kernel_fn(true, P.local, P.spj, 4 >> c));
            const void* ks2 = reinterpret_cast<const void*>(thin_kernel_fn(false, P.local, P.spj, 4 >> c));
            cudaFuncSetAttribute(kt2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ctx->smem_bytes);
            cudaFuncSetAttribute(ks2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ctx->smem_bytes);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kt2, CTA_THREADS, ctx->smem_bytes);
            ctx->grid_trace_c[c] = std::max(1, occ) * ctx->sm_count;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ks2, CTA_THREADS, ctx->smem_bytes);
            ctx->grid_score_c[c] = std::max(1, occ) * ctx->sm_count;
        }
    if (ctx->pk_ok) {
        if (const char* e = getenv("GSPALN_PK_NP")) ctx->pk_np_forced = atoi(e) == 8 ? 8 : 4;
        for (int v = 0; v < 2; ++v) {
            const int np = v ? 8 : 4;
            ctx->smem_pk[v] = (sizeof(PkRingA) + sizeof(PkRingB)) * 2 * np * CTA_THREADS + sizeof(uint2) * PK_T4 +
                              sizeof(PkPen) * (size_t) (cap + 1);
            const void* pt = reinterpret_cast<const void*>(pk_kernel_fn(true, P.spj, np));
            const void* ps = reinterpret_cast<const void*>(pk_kernel_fn(false, P.spj, np));
            cudaFuncSetAttribute(pt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ctx->smem_pk[v]);
            cudaFuncSetAttribute(ps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ctx->smem_pk[v]);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pt, CTA_THREADS, ctx->smem_pk[v]);
            ctx->grid_trace_pk[v] = std::max(1, occ) * ctx->sm_count;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ps, CTA_THREADS, ctx->smem_pk[v]);
            ctx->grid_score_pk[v] = std::max(1, occ) * ctx->sm_count;
        }
    }
    if (cudaGetLastError() != cudaSuccess) { gspaln_destroy(ctx); return GSPALN_ECUDA; }
    *out = ctx;
    return GSPALN_OK;
}

void gspaln_destroy(gspaln_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx->d_prm.release(); ctx->d_pen.release(); ctx->d_tasks.release(); ctx->d_order.release(); ctx->d_ticket.release();
    ctx->d_apool.release(); ctx->d_cpool.release(); ctx->d_band.release(); ctx->d_trace.release();
    ctx->d_skl.release(); ctx->d_res.release(); ctx->d_udh.release(); ctx->d_cpos.release(); ctx->d_ures.release();
    ctx->h_tasks.release(); ctx->h_order.release(); ctx->h_apool.release(); ctx->h_cpool.release();
    ctx->h_skl.release(); ctx->h_res.release(); ctx->h_cpos.release(); ctx->h_ures.release();
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    for (auto& e : ctx->ev_sync) if (e) cudaEventDestroy(e);
    ctx->h_marks.release();
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int gspaln_set_ng_tables(gspaln_ctx* ctx, const int16_t* sig53tab, const int16_t* penalty,
                         int32_t n_penalty, int32_t codonk1)
{
    if (!ctx || !sig53tab || !penalty || n_penalty < 1) return GSPALN_EINVAL;
    CK(cudaSetDevice(ctx->device));
    if (ctx->d_ngtab.reserve(544 + (size_t) n_penalty) != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, GSPALN_ENOMEM, "device allocation");
    }
    CK(cudaMemcpy(ctx->d_ngtab.p, sig53tab, 544 * sizeof(short), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_ngtab.p + 544, penalty, (size_t) n_penalty * sizeof(short), cudaMemcpyHostToDevice));
    ctx->hP.codonk1 = codonk1;
    CK(cudaMemcpy(ctx->d_prm.p, &ctx->hP, sizeof(DevParams), cudaMemcpyHostToDevice));
    ctx->n_pen = n_penalty;
    ctx->ng_ready = true;
    if (const char* e = getenv("GSPALN_NG_REC_EIGHTHS")) ctx->ng_rec_eighths = std::max(1, std::min(24, atoi(e)));
    return GSPALN_OK;
}

// Query bytes of one problem in the a pool: mw codes, the packed-kernel eligibility byte and, for
// the exact-ILD kinds with a Cip_score table, the int32 bonuses of rows a_left .. a_right behind them
static inline size_t cip_offset(const gspaln_task& t)
{
    const bool with = t.cip && (t.kind == GSPALN_FORWARD_NG || t.kind == GSPALN_SCOREALONE_NG ||
                                t.kind == GSPALN_HIRSCHBERG_NG);
    return with ? align_up((size_t) (t.a_right - t.a_left) + 1, 4) : 0;
}
static inline size_t a_span(const gspaln_task& t)
{
    const size_t mw1 = (size_t) (t.a_right - t.a_left) + 1, at = cip_offset(t);
    return align_up(at ? at + 4 * mw1 : mw1, 128);
}

// ---- planning: validation, longest-first order, pool offsets (assigned along that order so
// that any contiguous range of the order is contiguous in every pool), workspaces, grids
static int plan_batch(gspaln_ctx* ctx, const gspaln_task* tasks, int n)
{
    CK(cudaSetDevice(ctx->device));
    ctx->n = 0;
    ctx->cells.assign(n, 0);
    ctx->skl_cap.assign(n, 0);
    for (int i = 0; i < n; ++i) {
        const gspaln_task& t = tasks[i];
        if (t.a_right < t.a_left || t.b_right < t.b_left || t.up - t.lw + 3 < 0 ||
            (t.kind != GSPALN_FORWARD_WIP && t.kind != GSPALN_SCOREONLY_WIP && t.kind != GSPALN_HIRSCHBERG_WIP &&
             t.kind != GSPALN_FORWARD_NG && t.kind != GSPALN_SCOREALONE_NG && t.kind != GSPALN_HIRSCHBERG_NG) ||
            (t.kind == GSPALN_HIRSCHBERG_WIP && (t.n_imd < 1 || t.a_right - t.a_left < 2 || ctx->prm.noll != 2)) ||
            (t.kind == GSPALN_HIRSCHBERG_NG && (t.n_imd < 1 || t.a_right - t.a_left < 2 || !ctx->prm.spj)) ||
            ((t.kind == GSPALN_FORWARD_NG || t.kind == GSPALN_SCOREALONE_NG || t.kind == GSPALN_HIRSCHBERG_NG) &&
             ctx->prm.spj && (!ctx->ng_ready || !t.int53 || t.b_right - t.b_left >= ctx->n_pen)) ||
            !t.a || !t.b || (ctx->prm.spj && (!t.sig5 || !t.sig3))) {
            char msg[256];
            snprintf(msg, sizeof(msg), "bad task %d: kind %d a (%d, %d] b (%d, %d] band [%d, %d] n_imd %d",
                     i, t.kind, t.a_left, t.a_right, t.b_left, t.b_right, t.lw, t.up, t.n_imd);
            return fail(ctx, GSPALN_EINVAL, msg);
        }
        ctx->cells[i] = task_cells(t);
    }
    if (ctx->h_tasks.reserve(n + 1) != cudaSuccess || ctx->h_order.reserve(n + 1) != cudaSuccess)
        return fail(ctx, GSPALN_ENOMEM, "pinned host allocation");
    // largest problems first (longest-processing-time order for the ticket queue)
    std::iota(ctx->h_order.p, ctx->h_order.p + n, 0);
    std::stable_sort(ctx->h_order.p, ctx->h_order.p + n,
                     [&](int x, int y) { return ctx->cells[x] > ctx->cells[y]; });
    size_t a_bytes = 0, c_elems = 0, band_slab = 0, trace_slab = 0, skl_elems = 0;
    size_t udh_slab = 0, cpos_elems = 0;
    int n_trace = 0, n_score = 0, n_udh = 0, n_ng = 0, n_ngs = 0, n_xudh = 0, n_xudh_w = 0, n_ng_w = 0, n_ngs_w = 0;
    int n_udh_t = 0;
    // Hirschberg passes: the long queries (>= UDH_TEAM_ROWS rows) get a CTA (team of warps) each
    // instead of one warp when the batch could not keep every warp of the device busy for as long as
    // its largest problem would take on one warp -- with more than a warp's fair share of the batch's
    // cells that problem would be the tail of the launch.  In a saturated batch one warp per problem
    // is the faster form (no CTA barrier per step), and the two classes run one after the other, so
    // it is all of the long queries or none.
    // GSPALN_UDH_TEAM_ROWS=r puts every pass of >= r rows in the team class (tests, tuning).
    static const int team_rows_env = getenv("GSPALN_UDH_TEAM_ROWS") ? atoi(getenv("GSPALN_UDH_TEAM_ROWS")) : 0;
    const int udh_team_rows = team_rows_env > 0 ? team_rows_env : UDH_TEAM_ROWS;
    int64_t udh_team_cells = 0;
    if (team_rows_env <= 0) {
        int64_t total = 0, largest = 0;
        for (int i = 0; i < n; ++i)
            if (tasks[i].kind == GSPALN_HIRSCHBERG_WIP) {
                total += ctx->cells[i];
                if (tasks[i].a_right - tasks[i].a_left >= udh_team_rows) largest = std::max(largest, ctx->cells[i]);
            }
        if (largest <= total / (int64_t) std::max(1, ctx->grid_udh * WARPS_PER_CTA)) udh_team_cells = INT64_MAX;
    }
    size_t ng_width = 0, ng_rec = 0, ngs_width = 0, xudh_width = 0, xudh_links = 0;
    int n_trace_c[3] = {0, 0, 0}, n_score_c[3] = {0, 0, 0};
    size_t band_slab_c[3] = {0, 0, 0}, trace_slab_c[3] = {0, 0, 0};
    const bool dagp_prm = ctx->prm.noll == 3;
    for (int k = 0; k < n; ++k) {
        const int i = ctx->h_order.p[k];
        const gspaln_task& t = tasks[i];
        DevTask& d = ctx->h_tasks.p[i];
        d.kind = t.kind;
        d.a_left = t.a_left; d.a_right = t.a_right; d.b_left = t.b_left; d.b_right = t.b_right;
        d.lw = t.lw; d.up = t.up;
        d.flags = (t.a_exgl ? 1 : 0) | (t.a_exgr ? 2 : 0) | (t.b_exgl ? 4 : 0) | (t.b_exgr ? 8 : 0);
        d.skl_cap = (t.kind == GSPALN_FORWARD_WIP || t.kind == GSPALN_FORWARD_NG) ? std::max(0, t.skl_cap) : 0;
        d.pad0 = 0;
        const int mw = t.a_right - t.a_left, nw = t.b_right - t.b_left;
        const int width = t.up - t.lw + 3;
        // 128-byte granules: a problem whose inputs arrive later never shares a cache line with
        // one that is already being read (one-shot submits stream the batch in)
        d.a_off = (long long) a_bytes;      a_bytes += a_span(t);
        const size_t cip_at = cip_offset(t);
        d.col_off = (long long) c_elems;    c_elems += align_up((size_t) nw + 2, 16);
        const size_t bslab = align_up((size_t) width + 2 * NELEM, 32);
        d.skl_off = (long long) skl_elems;
        d.pad1 = (long long) cip_at;
        if (cip_at) d.flags |= 16;          // a Cip_score table rides behind the query codes
        int cls = 8;
        if (t.kind == GSPALN_FORWARD_WIP || t.kind == GSPALN_SCOREONLY_WIP) {
            cls = wip_class(mw, width, dagp_prm);
            d.pad0 = cls;
        }
        const int ci = cls == 4 ? 0 : (cls == 2 ? 1 : 2);
        if (cls == 8) band_slab = std::max(band_slab, bslab);
        else band_slab_c[ci] = std::max(band_slab_c[ci], bslab);
        if (t.kind == GSPALN_FORWARD_WIP) {
            const size_t nstrips = (mw + NELEM - 1) / NELEM;
            const size_t tslab = align_up(nstrips * (size_t) (width + TRACE_PAD) * NELEM + 64, 256);
            skl_elems += (size_t) d.skl_cap;
            if (cls == 8) { trace_slab = std::max(trace_slab, tslab); ++n_trace; }
            else { trace_slab_c[ci] = std::max(trace_slab_c[ci], tslab); ++n_trace_c[ci]; }
        } else if (t.kind == GSPALN_FORWARD_NG) {
            // path records: at most one per cell plus two per gap state at acceptor columns
            ng_width = std::max(ng_width, (size_t) width);
            const int64_t full = 3 * ctx->cells[i] + 2 * width + 64;
            const int64_t typical = ctx->cells[i] * ctx->ng_rec_eighths / 8 + 2 * width + 4096;
            ng_rec = std::max(ng_rec, (size_t) std::min<int64_t>(ctx->ng_full_records ? full : std::min(full, typical), INT_MAX / 4));
            skl_elems += (size_t) d.skl_cap;
            if (mw >= NG_WIDE_ROWS) { d.flags |= 32; ++n_ng_w; }           // a CTA of warps per problem
            else ++n_ng;
        } else if (t.kind == GSPALN_SCOREALONE_NG) {
            ngs_width = std::max(ngs_width, (size_t) width);
            if (mw >= NG_WIDE_ROWS) { d.flags |= 32; ++n_ngs_w; }
            else ++n_ngs;
        } else if (t.kind == GSPALN_HIRSCHBERG_NG) {
            // band rows of 32-byte cells + hlnk | vlnk | lwrb | uprb per intermediate row
            d.pad0 = t.n_imd;
            d.pad1 = (long long) cpos_elems;
            cpos_elems += (size_t) 10 * (t.n_imd + 1);
            xudh_width = std::max(xudh_width, (size_t) width);
            xudh_links = std::max(xudh_links, (size_t) t.n_imd * 4 * ctx->prm.noll * (size_t) width);
            if (mw >= XUDH_WIDE_ROWS) { d.flags |= 32; ++n_xudh_w; }       // a CTA of warps per problem
            else ++n_xudh;
        } else if (t.kind == GSPALN_HIRSCHBERG_WIP) {
            d.pad0 = t.n_imd;
            d.pad1 = (long long) cpos_elems;
            cpos_elems += (size_t) 10 * (t.n_imd + 1);
            udh_slab = std::max(udh_slab, align_up(4 * ((size_t) width + 2 * NELEM + 2) +
                                                   (size_t) t.n_imd * 4 * width + 8, 64));
            if (mw >= udh_team_rows && ctx->cells[i] > udh_team_cells) { d.flags |= 32; ++n_udh_t; }   // a CTA per problem
            else ++n_udh;
        } else if (cls == 8)
            ++n_score;
        else
            ++n_score_c[ci];
        ctx->skl_cap[i] = d.skl_cap;
    }
    if (ctx->h_apool.reserve(a_bytes + 16) != cudaSuccess || ctx->h_cpool.reserve(c_elems + 4) != cudaSuccess ||
        ctx->h_res.reserve(n + 1) != cudaSuccess || ctx->h_skl.reserve(skl_elems + 1) != cudaSuccess ||
        ctx->h_cpos.reserve(cpos_elems + 1) != cudaSuccess || ctx->h_ures.reserve(n + 1) != cudaSuccess)
        return fail(ctx, GSPALN_ENOMEM, "pinned host allocation");
    if (ctx->d_tasks.reserve(n + 1) != cudaSuccess || ctx->d_order.reserve(n + 1) != cudaSuccess ||
        ctx->d_apool.reserve(a_bytes + 16) != cudaSuccess || ctx->d_cpool.reserve(c_elems + 4) != cudaSuccess ||
        ctx->d_skl.reserve(skl_elems + 1) != cudaSuccess || ctx->d_res.reserve(n + 1) != cudaSuccess ||
        ctx->d_cpos.reserve(cpos_elems + 1) != cudaSuccess || ctx->d_ures.reserve(n + 1) != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, GSPALN_ENOMEM, "device allocation");
    }
    // per-warp workspaces; shrink the grid if the trace slabs would not fit
    {
        auto ctas = [&](int full, int count) {
            return std::max(1, std::min(full, (count + WARPS_PER_CTA - 1) / WARPS_PER_CTA));
        };
        // packed kernels: 8 registers per thread (one thread per strip, 32 strip slots per warp) pay
        // off when the queries are long enough to keep 32 strips in flight, 4 otherwise
        {
            double rows_w = 0, w = 0;
            for (int i = 0; i < n; ++i) {
                rows_w += (double) ctx->cells[i] * (tasks[i].a_right - tasks[i].a_left);
                w += (double) ctx->cells[i];
            }
            ctx->pk_np = ctx->pk_np_forced ? ctx->pk_np_forced : (w > 0 && rows_w / w >= 768 ? 8 : 4);
        }
        const int v8 = ctx->pk_np == 8;
        int gt = n_trace ? ctas(std::max(ctx->grid_trace, ctx->grid_trace_pk[v8]), n_trace) : 0;
        int gs = n_score ? ctas(std::max(ctx->grid_score, ctx->grid_score_pk[v8]), n_score) : 0;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        free_b += ctx->d_trace.cap + ctx->d_band.cap * sizeof(unsigned);
        const size_t budget = (size_t) (0.85 * (double) free_b);
        while (gt > 1 && (size_t) gt * WARPS_PER_CTA * (trace_slab + band_slab * 8) > budget) gt = gt * 3 / 4;
        int gu = n_udh ? ctas(ctx->grid_udh, n_udh) : 0;
        const int gut = std::min(ctx->grid_udh_t, n_udh_t);
        gu = std::max(gu, gut);                             // the two classes share the per-warp slabs
        const size_t warps = (size_t) std::max(std::max(gt, gs), gu) * WARPS_PER_CTA;
        // the narrower chain classes share the pools (the kernels run one after the other)
        size_t band_words = warps * band_slab * (ctx->prm.noll == 3 ? 2 : 1);
        size_t trace_bytes = (size_t) gt * WARPS_PER_CTA * trace_slab;
        for (int c = 0; c < 3; ++c) {
            int g1 = n_trace_c[c] ? ctas(ctx->grid_trace_c[c], n_trace_c[c]) : 0;
            const int g2 = n_score_c[c] ? ctas(ctx->grid_score_c[c], n_score_c[c]) : 0;
            while (g1 > 1 && (size_t) g1 * WARPS_PER_CTA * (trace_slab_c[c] + band_slab_c[c] * 8) > budget) g1 = g1 * 3 / 4;
            ctx->grid_run_trace_c[c] = g1; ctx->grid_run_score_c[c] = g2;
            ctx->n_trace_c[c] = n_trace_c[c]; ctx->n_score_c[c] = n_score_c[c];
            ctx->band_slab_c[c] = band_slab_c[c]; ctx->trace_slab_c[c] = trace_slab_c[c];
            band_words = std::max(band_words, (size_t) std::max(g1, g2) * WARPS_PER_CTA * band_slab_c[c]);
            trace_bytes = std::max(trace_bytes, (size_t) g1 * WARPS_PER_CTA * trace_slab_c[c]);
        }
        if (ctx->d_udh.reserve((size_t) gu * WARPS_PER_CTA * udh_slab + 64) != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, GSPALN_ENOMEM, "device UDH workspace allocation");
        }
        ctx->grid_run_udh = n_udh ? ctas(ctx->grid_udh, n_udh) : 0;
        ctx->grid_run_udh_t = gut;
        if (ctx->d_band.reserve(band_words + 32) != cudaSuccess ||
            ctx->d_trace.reserve(trace_bytes + 256) != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, GSPALN_ENOMEM, "device workspace allocation");
        }
        ctx->grid_run_trace = gt;
        ctx->grid_run_score = gs;
        // exact-ILD kernels: one warp per problem, per-warp workspace = band rows (H, F, F2 as
        // {value, record}) + direction bytes + path records; the trace-back and the score-only
        // kernel never run at the same time, so they share the pool
        ctx->grid_run_ng = ctx->grid_run_ngs = ctx->grid_run_xudh = ctx->grid_run_xudh_w = 0;
        ctx->grid_run_ng_w = ctx->grid_run_ngs_w = 0;
        if (n_ng || n_ngs || n_xudh || n_xudh_w || n_ng_w || n_ngs_w) {
            const size_t w = std::max(ng_width, ngs_width);
            const size_t rec = (n_ng || n_ng_w) ? ng_rec + 32 * NG_WIDE * NG_CHUNK + 64 : 64;
            size_t slab = align_up(3 * w * sizeof(NgRvp) + align_up(w, 16) + 12 * rec, 16);
            // the scalar Hirschberg pass shares the pool (the kernels run one after the other)
            const size_t xslab = (n_xudh || n_xudh_w) ? align_up(3 * xudh_width * sizeof(UxSlot) + 4 * xudh_links + 64, 32) : 0;
            slab = align_up(std::max(slab, xslab), 32);
            // (the wide class runs one problem per CTA: a quarter of the slabs of a 4-warp CTA)
            const int wide = std::max(std::max(n_xudh_w, n_ng_w), n_ngs_w);
            const int need = std::max(std::max(std::max(n_ng, n_ngs), n_xudh), std::min(wide, 2 * ctx->sm_count));
            int g = std::min((need + NG_WARPS - 1) / NG_WARPS, 4 * ctx->sm_count);
            cudaMemGetInfo(&free_b, &total_b);
            const size_t room = (size_t) (0.5 * (double) (free_b + ctx->d_ngwork.cap));
            while (g > 1 && (size_t) g * NG_WARPS * slab > room) g = g * 3 / 4;
            if (ctx->d_ngwork.reserve((size_t) g * NG_WARPS * slab + 64) != cudaSuccess) {
                cudaGetLastError();
                return fail(ctx, GSPALN_ENOMEM, "device exact-ILD workspace allocation");
            }
            ctx->grid_run_ng = n_ng ? std::min(g, (n_ng + NG_WARPS - 1) / NG_WARPS) : 0;
            ctx->grid_run_ngs = n_ngs ? std::min(g, (n_ngs + NG_WARPS - 1) / NG_WARPS) : 0;
            ctx->grid_run_xudh = n_xudh ? std::min(g, (n_xudh + NG_WARPS - 1) / NG_WARPS) : 0;
            ctx->grid_run_xudh_w = n_xudh_w ? std::min(std::min(n_xudh_w, g * NG_WARPS), 2 * ctx->sm_count) : 0;
            ctx->grid_run_ng_w = n_ng_w ? std::min(std::min(n_ng_w, g * NG_WARPS), 2 * ctx->sm_count) : 0;
            ctx->grid_run_ngs_w = n_ngs_w ? std::min(std::min(n_ngs_w, g * NG_WARPS), 2 * ctx->sm_count) : 0;
            ctx->ng_slab = slab; ctx->ng_width = w; ctx->ng_rec_cap = (int) rec;
            ctx->xudh_width = xudh_width;
        }
    }
    ctx->n = n; ctx->n_trace = n_trace; ctx->n_score = n_score; ctx->n_udh = n_udh; ctx->n_udh_t = n_udh_t; ctx->n_ng = n_ng;
    ctx->n_ngs = n_ngs; ctx->n_xudh = n_xudh; ctx->n_xudh_w = n_xudh_w; ctx->n_ng_w = n_ng_w; ctx->n_ngs_w = n_ngs_w;
    ctx->udh_slab = udh_slab; ctx->cpos_elems = cpos_elems;
    ctx->a_bytes = a_bytes; ctx->c_elems = c_elems; ctx->band_slab = band_slab;
    ctx->trace_slab = trace_slab; ctx->skl_elems = skl_elems;
    int64_t cells = 0, tb = 0;
    for (int i = 0; i < n; ++i) {
        cells += ctx->cells[i];
        if (tasks[i].kind == GSPALN_FORWARD_WIP) tb += ctx->cells[i];
    }
    ctx->tim.cells = cells;
    ctx->tim.trace_bytes = tb;
    return GSPALN_OK;
}

// ---- packing of the problems order[lo .. hi) into the pinned pools (host work is part of the
// end-to-end path): dealt to a few host threads, every problem writes a disjoint slice
static void pack_range(gspaln_ctx* ctx, const gspaln_task* tasks, int lo, int hi)
{
    auto pack_some = [&](int klo, int khi) {
        for (int k = klo; k < khi; ++k) {
            const int i = ctx->h_order.p[k];
            const gspaln_task& t = tasks[i];
            const DevTask& d = ctx->h_tasks.p[i];
            const int mw = t.a_right - t.a_left, nw = t.b_right - t.b_left;
            unsigned char* ap = ctx->h_apool.p + d.a_off;
            // residues the packed kernel knows: A, C, G, T, N (codes 2, 3, 5, 9, 16 of the DNA alphabet)
            constexpr unsigned PK_CODES = (1u << 2) | (1u << 3) | (1u << 5) | (1u << 9) | (1u << 16);
            unsigned seen = 0;
            int sigmax = 0;
            for (int j = 0; j < mw; ++j) {
                const unsigned c = t.a[t.a_left + j] & 31;
                seen |= 1u << c;
                ap[j] = ctx->perm[c];
            }
            ColInfo* col = ctx->h_cpool.p + d.col_off;
            const bool spj = ctx->prm.spj != 0;
            for (int j = 0; j <= nw; ++j) {
                const int c = t.b_left + j;     // column c pairs genome residue at(c - 1)
                ColInfo ci;
                ci.sig5 = spj ? t.sig5[c] : 0;
                ci.sig3 = spj ? t.sig3[c] : 0;
                sigmax = std::max(sigmax, std::max(std::abs((int) ci.sig5), std::abs((int) ci.sig3)));
                if (j > 0) seen |= 1u << (t.b[c - 1] & 31);
                ci.code = j > 0 ? ctx->perm[t.b[c - 1] & 31] : 0;
                // INT53 nibbles ride in the padding: [0] dinc5 | dinc3 << 4, [1] cano5 | cano3 << 4
                const unsigned i53 = t.int53 ? t.int53[c] : 0u;
                ci.pad[0] = (unsigned char) (i53 & 0xffu);
                ci.pad[1] = (unsigned char) (i53 >> 8);
                ci.pad[2] = 0;
                col[j] = ci;
            }
            // the byte behind the query codes: this problem may run on the packed int16x2 kernel
            ap[mw] = (ctx->pk_ok && !(seen & ~PK_CODES) && sigmax <= PK_SIGMAX &&
                      (t.kind == GSPALN_FORWARD_WIP || t.kind == GSPALN_SCOREONLY_WIP)) ? 1 : 0;
            if (const size_t cip_at = cip_offset(t))
                memcpy(ap + cip_at, t.cip + t.a_left, sizeof(int32_t) * ((size_t) mw + 1));     // rows a_left .. a_right
        }
    };
    size_t work = 0;
    for (int k = lo; k < hi; ++k) {
        const gspaln_task& t = tasks[ctx->h_order.p[k]];
        work += (size_t) (t.b_right - t.b_left) + (t.a_right - t.a_left);
    }
    int nthr = (int) std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 16);
    if (work < (1u << 20) || hi - lo < 2 * nthr) nthr = 1;
    if (nthr == 1) { pack_some(lo, hi); return; }
    std::vector<std::thread> pool;
    size_t acc = 0, per = (work + nthr - 1) / nthr;
    int from = lo;
    for (int k = lo; k < hi; ++k) {
        const gspaln_task& t = tasks[ctx->h_order.p[k]];
        acc += (size_t) (t.b_right - t.b_left) + (t.a_right - t.a_left);
        if (acc >= per || k == hi - 1) {
            pool.emplace_back(pack_some, from, k + 1);
            from = k + 1; acc = 0;
        }
    }
    for (auto& th : pool) th.join();
}

// pool ranges [begin, end) covered by order[lo .. hi)
static void pool_span(const gspaln_ctx* ctx, const gspaln_task* tasks, int lo, int hi, size_t& a0, size_t& a1,
                      size_t& c0, size_t& c1)
{
    const DevTask& f = ctx->h_tasks.p[ctx->h_order.p[lo]];
    const int li = ctx->h_order.p[hi - 1];
    const DevTask& l = ctx->h_tasks.p[li];
    a0 = (size_t) f.a_off; c0 = (size_t) f.col_off;
    a1 = (size_t) l.a_off + a_span(tasks[li]);
    c1 = (size_t) l.col_off + align_up((size_t) (tasks[li].b_right - tasks[li].b_left) + 2, 16);
}

// launches the kernels for order[lo .. hi) with ticket slot `slot` (3 counters per slot)
static int launch_range(gspaln_ctx* ctx, int lo, int hi, int slot, int& launches, const int* ready = nullptr)
{
    const bool local = ctx->prm.local != 0, spj = ctx->prm.spj != 0, dagp = ctx->prm.noll == 3;
    int* tick = ctx->d_ticket.p + 3 * slot;
    const int cnt = hi - lo;
    // packed int16x2 kernels first (problems the host marked eligible); the 32-bit kernels behind
    // them take the rest and whatever the packed ones hand back (status 6)
    if (ctx->pk_ok && ctx->n_trace) {
        pk_kernel_fn(true, spj, ctx->pk_np)<<<std::min(ctx->grid_run_trace, ctx->grid_trace_pk[ctx->pk_np == 8]), CTA_THREADS, ctx->smem_pk[ctx->pk_np == 8], ctx->stream>>>(
            ctx->d_prm.p, ctx->d_pen.p, ctx->d_tasks.p, ctx->d_order.p + lo, cnt, tick + 4,
            ctx->d_apool.p, ctx->d_cpool.p, ctx->d_band.p, (long long) ctx->band_slab,
            ctx->d_trace.p, (long long) ctx->trace_slab, ctx->d_skl.p, ctx->d_res.p, ready);
        ++launches;
    }
    if (ctx->pk_ok && ctx->n_score) {
        pk_kernel_fn(false, spj, ctx->pk_np)<<<std::min(ctx->grid_run_score, ctx->grid_score_pk[ctx->pk_np == 8]), CTA_THREADS, ctx->smem_pk[ctx->pk_np == 8], ctx->stream>>>(
            ctx->d_prm.p, ctx->d_pen.p, ctx->d_tasks.p, ctx->d_order.p + lo, cnt, tick + 5,
            ctx->d_apool.p, ctx->d_cpool.p, ctx->d_band.p, (long long) ctx->band_slab,
            ctx->d_trace.p, (long long) ctx->trace_slab, ctx->d_skl.p, ctx->d_res.p, ready);
        ++launches;
    }
    if (ctx->n_trace) {
        kernel_fn(true, local, spj, dagp)<<<std::min(ctx->grid_run_trace, ctx->grid_trace), CTA_THREADS, ctx->smem_bytes, ctx->stream>>>(
            ctx->d_prm.p, ctx->d_pen.p, ctx->d_tasks.p, ctx->d_order.p + lo, cnt, tick,
            ctx->d_apool.p, ctx->d_cpool.p, ctx->d_band.p, (long long) ctx->band_slab,
            ctx->d_trace.p, (long long) ctx->trace_slab, ctx->d_skl.p, ctx->d_res.p, ready);
        ++launches;
    }
    if (ctx->n_score) {
        kernel_fn(false, local, spj, dagp)<<<std::min(ctx->grid_run_score, ctx->grid_score), CTA_THREADS, ctx->smem_bytes, ctx->stream>>>(
            ctx->d_prm.p, ctx->d_pen.p, ctx->d_tasks.p, ctx->d_order.p + lo, cnt, tick + 1,
            ctx->d_apool.p, ctx->d_cpool.p, ctx->d_band.p, (long long) ctx->band_slab,
            ctx->d_trace.p, (long long) ctx->trace_slab, ctx->d_skl.p, ctx->d_res.p, ready);
        ++launches;
    }
    for (int c = 0; c < 3 && !dagp; ++c) {
        if (ctx->n_trace_c[c]) {
            thin_kernel_fn(true, local, spj, 4 >> c)<<<ctx->grid_run_trace_c[c], CTA_THREADS, ctx->smem_bytes, ctx->stream>>>(
                ctx->d_prm.p, ctx->d_pen.p, ctx->d_tasks.p, ctx->d_order.p + lo, cnt, tick + 12 + 2 * c,
                ctx->d_apool.p, ctx->d_cpool.p, ctx->d_band.p, (long long) ctx->band_slab_c[c],
                ctx->d_trace.p, (long long) ctx->trace_slab_c[c], ctx->d_skl.p, ctx->d_res.p, ready);
            ++launches;
        }
        if (ctx->n_score_c[c]) {
            thin_kernel_fn(false, local, spj, 4 >> c)<<<ctx->grid_run_score_c[c], CTA_THREADS, ctx->smem_bytes, ctx->stream>>>(
                ctx->d_prm.p, ctx->d_pen.p, ctx->d_tasks.p, ctx->d_order.p + lo, cnt, tick + 13 + 2 * c,
                ctx->d_apool.p, ctx->d_cpool.p, ctx->d_band.p, (long long) ctx->band_slab_c[c],
                ctx->d_trace.p, (long long) ctx->trace_slab_c[c], ctx->d_skl.p, ctx->d_res.p, ready);
            ++launches;
        }
    }
    if (ctx->n_udh_t) {
        auto kt = udh_team_kernel_fn(spj, local);
        kt<<<ctx->grid_run_udh_t, CTA_THREADS, ctx->smem_bytes, ctx->stream>>>(
            ctx->d_prm.p, ctx->d_pen.p, ctx->d_tasks.p, ctx->d_order.p + lo, cnt, tick + 3,
            ctx->d_apool.p, ctx->d_cpool.p, ctx->d_band.p, (long long) ctx->band_slab,
            ctx->d_udh.p, (long long) ctx->udh_slab, ctx->d_cpos.p, ctx->d_ures.p, ready);
        ++launches;
    }
    if (ctx->n_udh) {
        auto ku = udh_kernel_fn(spj, local);
        ku<<<ctx->grid_run_udh, CTA_THREADS, ctx->smem_bytes, ctx->stream>>>(
            ctx->d_prm.p, ctx->d_pen.p, ctx->d_tasks.p, ctx->d_order.p + lo, cnt, tick + 2,
            ctx->d_apool.p, ctx->d_cpool.p, ctx->d_band.p, (long long) ctx->band_slab,
            ctx->d_udh.p, (long long) ctx->udh_slab, ctx->d_cpos.p, ctx->d_ures.p, ready);
        ++launches;
    }
    if (ctx->n_ng_w) {
        dp_xild_kernel<false, NG_WIDE><<<ctx->grid_run_ng_w, 32 * NG_WIDE, 0, ctx->stream>>>(
            ctx->d_prm.p, ctx->d_ngtab.p, ctx->n_pen, ctx->d_tasks.p, ctx->d_order.p + lo, cnt, tick + 6,
            ctx->d_apool.p, ctx->d_cpool.p, ctx->d_ngwork.p, (long long) ctx->ng_slab,
            (long long) ctx->ng_width, ctx->ng_rec_cap, ctx->d_skl.p, ctx->d_res.p, ready);
        ++launches;
    }
    if (ctx->n_ngs_w) {
        dp_xild_kernel<true, NG_WIDE><<<ctx->grid_run_ngs_w, 32 * NG_WIDE, 0, ctx->stream>>>(
            ctx->d_prm.p, ctx->d_ngtab.p, ctx->n_pen, ctx->d_tasks.p, ctx->d_order.p + lo, cnt, tick + 7,
            ctx->d_apool.p, ctx->d_cpool.p, ctx->d_ngwork.p, (long long) ctx->ng_slab,
            (long long) ctx->ng_width, ctx->ng_rec_cap, ctx->d_skl.p, ctx->d_res.p, ready);
        ++launches;
    }
    if (ctx->n_ng) {
        dp_xild_kernel<false, 1><<<ctx->grid_run_ng, NG_THREADS, 0, ctx->stream>>>(
            ctx->d_prm.p, ctx->d_ngtab.p, ctx->n_pen, ctx->d_tasks.p, ctx->d_order.p + lo, cnt, tick + 8,
            ctx->d_apool.p, ctx->d_cpool.p, ctx->d_ngwork.p, (long long) ctx->ng_slab,
            (long long) ctx->ng_width, ctx->ng_rec_cap, ctx->d_skl.p, ctx->d_res.p, ready);
        ++launches;
    }
    if (ctx->n_ngs) {
        dp_xild_kernel<true, 1><<<ctx->grid_run_ngs, NG_THREADS, 0, ctx->stream>>>(
            ctx->d_prm.p, ctx->d_ngtab.p, ctx->n_pen, ctx->d_tasks.p, ctx->d_order.p + lo, cnt, tick + 9,
            ctx->d_apool.p, ctx->d_cpool.p, ctx->d_ngwork.p, (long long) ctx->ng_slab,
            (long long) ctx->ng_width, ctx->ng_rec_cap, ctx->d_skl.p, ctx->d_res.p, ready);
        ++launches;
    }
    if (ctx->n_xudh_w) {
        dp_xudh_kernel<XUDH_WIDE><<<ctx->grid_run_xudh_w, 32 * XUDH_WIDE, 0, ctx->stream>>>(
            ctx->d_prm.p, ctx->d_ngtab.p, ctx->n_pen, ctx->d_tasks.p, ctx->d_order.p + lo, cnt, tick + 11,
            ctx->d_apool.p, ctx->d_cpool.p, ctx->d_ngwork.p, (long long) ctx->ng_slab,
            (long long) ctx->xudh_width, ctx->d_cpos.p, ctx->d_ures.p, ready);
        ++launches;
    }
    if (ctx->n_xudh) {
        dp_xudh_kernel<1><<<ctx->grid_run_xudh, NG_THREADS, 0, ctx->stream>>>(
            ctx->d_prm.p, ctx->d_ngtab.p, ctx->n_pen, ctx->d_tasks.p, ctx->d_order.p + lo, cnt, tick + 10,
            ctx->d_apool.p, ctx->d_cpool.p, ctx->d_ngwork.p, (long long) ctx->ng_slab,
            (long long) ctx->xudh_width, ctx->d_cpos.p, ctx->d_ures.p, ready);
        ++launches;
    }
    CK(cudaGetLastError());
    return GSPALN_OK;
}

constexpr int MAX_CHUNKS = 16;

int gspaln_upload(gspaln_ctx* ctx, const gspaln_task* tasks, int n)
{
    if (!ctx || !tasks || n < 0) return GSPALN_EINVAL;
    int rc = plan_batch(ctx, tasks, n);
    if (rc != GSPALN_OK) return rc;
    if (n) pack_range(ctx, tasks, 0, n);
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_tasks.p, ctx->h_tasks.p, sizeof(DevTask) * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_order.p, ctx->h_order.p, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_apool.p, ctx->h_apool.p, ctx->a_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_cpool.p, ctx->h_cpool.p, sizeof(ColInfo) * ctx->c_elems, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->tim.h2d_ms = ms;
    ctx->tim.h2d_bytes = (int64_t) (sizeof(DevTask) * n + sizeof(int) * n + ctx->a_bytes + sizeof(ColInfo) * ctx->c_elems);
    return GSPALN_OK;
}

int gspaln_run(gspaln_ctx* ctx)
{
    if (!ctx) return GSPALN_EINVAL;
    CK(cudaSetDevice(ctx->device));
    const int n = ctx->n;
    int launches = 0;
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    if (n > 0) {
        CK(cudaMemsetAsync(ctx->d_ticket.p, 0, 32 * sizeof(int), ctx->stream));
        int rc = launch_range(ctx, 0, n, 0, launches);
        if (rc != GSPALN_OK) return rc;
    }
    CK(cudaEventRecord(ctx->ev[3], ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
    ctx->tim.kernel_ms = ms;
    ctx->tim.launches = launches;
    return GSPALN_OK;
}

int gspaln_download(gspaln_ctx* ctx, gspaln_result* results)
{
    if (!ctx || (!results && ctx->n)) return GSPALN_EINVAL;
    CK(cudaSetDevice(ctx->device));
    const int n = ctx->n;
    CK(cudaEventRecord(ctx->ev[4], ctx->stream));
    if (n) CK(cudaMemcpyAsync(ctx->h_res.p, ctx->d_res.p, sizeof(DevResult) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->skl_elems)
        CK(cudaMemcpyAsync(ctx->h_skl.p, ctx->d_skl.p, sizeof(int2) * ctx->skl_elems, cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->n_udh || ctx->n_udh_t || ctx->n_xudh || ctx->n_xudh_w) {
        CK(cudaMemcpyAsync(ctx->h_ures.p, ctx->d_ures.p, sizeof(DevUdhOut) * n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_cpos.p, ctx->d_cpos.p, sizeof(int) * ctx->cpos_elems, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaEventRecord(ctx->ev[5], ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]);
    ctx->tim.d2h_ms = ms;
    ctx->tim.d2h_bytes = (int64_t) (sizeof(DevResult) * n + sizeof(int2) * ctx->skl_elems);
    for (int i = 0; i < n; ++i) {
        const DevResult& r = ctx->h_res.p[i];
        gspaln_result& o = results[i];
        o.score = r.score; o.status = r.status; o.n_skl = r.n_skl; o.reserved = 0;
        o.cells = ctx->cells[i];
        const DevTask& d = ctx->h_tasks.p[i];
        if (d.kind == GSPALN_HIRSCHBERG_WIP || d.kind == GSPALN_HIRSCHBERG_NG) {
            const DevUdhOut& u = ctx->h_ures.p[i];
            o.score = u.score; o.status = u.status; o.n_skl = 0;
            o.ranges[0] = u.a_left; o.ranges[1] = u.a_right; o.ranges[2] = u.b_left; o.ranges[3] = u.b_right;
            if (o.cpos) memcpy(o.cpos, ctx->h_cpos.p + d.pad1, sizeof(int) * 10 * (size_t) (d.pad0 + 1));
            continue;
        }
        if (o.skl && d.skl_cap > 0) {
            const int cnt = std::min(r.n_skl, d.skl_cap);
            memcpy(o.skl, ctx->h_skl.p + d.skl_off, sizeof(int2) * (size_t) std::max(0, cnt));
        }
    }
    return GSPALN_OK;
}

// One-shot path.  The persistent kernels consume the problems in longest-first ticket order, so
// a large batch is streamed in along that order: the first chunk is packed and copied, the
// kernels start, and while they run the host threads pack the following chunks into pinned
// memory and a second stream copies them and advances a watermark the kernels wait on.
// Packing and H2D of everything but the first chunk hide behind the DP.
static int submit_once(gspaln_ctx* ctx, const gspaln_task* tasks, int n, gspaln_result* results);

// The path records of GSPALN_FORWARD_NG (the reference's Vmf) can number three per cell in theory
// and a fraction of one in practice; sizing every warp's store for the worst case leaves room for a
// few dozen problems in flight.  So the first run gives each problem a store for its typical need
// and the (rare) problems that report GSPALN_ST_VMF_OVERFLOW are run once more with the full bound.
int gspaln_submit(gspaln_ctx* ctx, const gspaln_task* tasks, int n, gspaln_result* results)
{
    if (!ctx || !tasks || n < 0) return GSPALN_EINVAL;
    int rc = submit_once(ctx, tasks, n, results);
    if (rc != GSPALN_OK || ctx->ng_full_records) return rc;
    std::vector<int> again;
    for (int i = 0; i < n; ++i)
        if (tasks[i].kind == GSPALN_FORWARD_NG && results[i].status == GSPALN_ST_VMF_OVERFLOW) again.push_back(i);
    if (again.empty()) return GSPALN_OK;
    static const bool dbg = getenv("GSPALN_LSP_DEBUG") != nullptr;
    if (dbg) fprintf(stderr, "gspaln: %zu of %d exact-ILD trace-backs need the full record store\n", again.size(), n);
    std::vector<gspaln_task> t2(again.size());
    std::vector<gspaln_result> r2(again.size());
    for (size_t k = 0; k < again.size(); ++k) { t2[k] = tasks[again[k]]; r2[k] = results[again[k]]; }
    const gspaln_timing first = ctx->tim;
    ctx->ng_full_records = true;
    rc = submit_once(ctx, t2.data(), (int) t2.size(), r2.data());
    ctx->ng_full_records = false;
    if (rc != GSPALN_OK) return rc;
    for (size_t k = 0; k < again.size(); ++k) results[again[k]] = r2[k];
    ctx->tim.kernel_ms += first.kernel_ms; ctx->tim.h2d_ms += first.h2d_ms; ctx->tim.d2h_ms += first.d2h_ms;
    ctx->tim.launches += first.launches; ctx->tim.h2d_bytes += first.h2d_bytes; ctx->tim.d2h_bytes += first.d2h_bytes;
    ctx->tim.cells += first.cells; ctx->tim.trace_bytes += first.trace_bytes;
    return GSPALN_OK;
}

static int submit_once(gspaln_ctx* ctx, const gspaln_task* tasks, int n, gspaln_result* results)
{
    int rc = plan_batch(ctx, tasks, n);
    if (rc != GSPALN_OK) return rc;
    int bounds[MAX_CHUNKS + 1];
    int nchunks = 1;
    bounds[0] = 0; bounds[1] = n;
    const size_t total = ctx->c_elems;
    // GSPALN_NO_STREAM=1: one chunk (profilers serialise the copy stream behind the running kernel,
    // which would leave the kernel waiting for its watermark until the in-kernel time-out)
    static const bool no_stream = getenv("GSPALN_NO_STREAM") != nullptr;
    if (!no_stream && n >= 256 && total >= (4u << 20)) {
        // boundaries by cumulative columns (what packing and H2D cost); small first chunk
        nchunks = 0;
        size_t acc = 0;
        const int want = (int) std::min<size_t>(MAX_CHUNKS, 2 + total / (6u << 20));
        size_t next = total / (2 * (size_t) want);
        for (int k = 0; k < n; ++k) {
            const gspaln_task& t = tasks[ctx->h_order.p[k]];
            acc += (size_t) (t.b_right - t.b_left) + 2;
            if (acc >= next && nchunks + 1 < want && k + 1 < n) {
                bounds[++nchunks] = k + 1;
                next = acc + (total - acc) / (size_t) (want - nchunks);
            }
        }
        bounds[++nchunks] = n;
    }
    if (ctx->h_marks.reserve(MAX_CHUNKS + 1) != cudaSuccess) return fail(ctx, GSPALN_ENOMEM, "pinned host allocation");
    int* d_ready = ctx->d_ticket.p + 32;
    CK(cudaMemsetAsync(ctx->d_ticket.p, 0, 40 * sizeof(int), ctx->stream));
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_tasks.p, ctx->h_tasks.p, sizeof(DevTask) * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_order.p, ctx->h_order.p, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
    int launches = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int lo = bounds[c], hi = bounds[c + 1];
        if (hi <= lo) continue;
        pack_range(ctx, tasks, lo, hi);
        size_t a0, a1, c0, c1;
        pool_span(ctx, tasks, lo, hi, a0, a1, c0, c1);
        cudaStream_t st = c == 0 ? ctx->stream : ctx->copy_stream;
        CK(cudaMemcpyAsync(ctx->d_apool.p + a0, ctx->h_apool.p + a0, a1 - a0, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->d_cpool.p + c0, ctx->h_cpool.p + c0, sizeof(ColInfo) * (c1 - c0), cudaMemcpyHostToDevice, st));
        ctx->h_marks.p[c] = hi;
        CK(cudaMemcpyAsync(d_ready, ctx->h_marks.p + c, sizeof(int), cudaMemcpyHostToDevice, st));
        if (c == 0) {
            // the kernels start once the first chunk (and the watermark reset) is in place; the
            // copy stream must not advance the watermark before that reset either
            CK(cudaEventRecord(ctx->ev[1], ctx->stream));
            CK(cudaEventRecord(ctx->ev_sync[0], ctx->stream));
            CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_sync[0], 0));
            CK(cudaEventRecord(ctx->ev[2], ctx->stream));
            rc = launch_range(ctx, 0, n, 0, launches, nchunks > 1 ? d_ready : nullptr);
            if (rc != GSPALN_OK) return rc;
        }
    }
    if (n == 0) { CK(cudaEventRecord(ctx->ev[1], ctx->stream)); CK(cudaEventRecord(ctx->ev[2], ctx->stream)); }
    CK(cudaEventRecord(ctx->ev[3], ctx->stream));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->tim.h2d_ms = ms;           // first chunk only: the rest overlaps the kernels
    cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
    ctx->tim.kernel_ms = ms;
    ctx->tim.launches = launches;
    ctx->tim.h2d_bytes = (int64_t) (sizeof(DevTask) * n + sizeof(int) * n + ctx->a_bytes + sizeof(ColInfo) * ctx->c_elems);
    return gspaln_download(ctx, results);
}

int gspaln_get_timing(const gspaln_ctx* ctx, gspaln_timing* out)
{
    if (!ctx || !out) return GSPALN_EINVAL;
    *out = ctx->tim;
    return GSPALN_OK;
}

}   // extern "C"
#include "gspaln_lsp.inl"

namespace {

// DNA x genome: Aln2s1::lspS_ng (src/fwd2s1.cc:1801-1897)
struct LspTraitsS {
    using Ctx = gspaln_ctx;
    using Task = gspaln_task;
    static constexpr int WPAD = 3;
    static constexpr bool SCALAR_MODE = true;               // -A0: forwardS_ng + hirschbergS_ng on the device
    static constexpr int KIND_SCALAR_UDH = GSPALN_HIRSCHBERG_NG;
    static void stripe(LspGeo& g, int sh)       // stripe(), src/aln2.cc:156-176
    {
        if (sh < 0) {
            const int shorter = std::min(g.a_right - g.a_left, g.b_right - g.b_left);
            sh = -sh * shorter / 100;
        }
        int up = g.b_right - g.a_right;
        int lw = g.b_left - g.a_left;
        if (up < lw) std::swap(up, lw);
        up += sh; lw -= sh;
        int q;
        if ((q = g.b_right - g.a_left) < up) up = q;
        if ((q = g.b_left - g.a_right) > lw) lw = q;
        g.up = up; g.lw = lw;
    }
    static bool small(int m, int nn) { return std::abs(nn - m) < 8 || m == 1 || nn == 1; }
    static float cvol(int m, int nn) { return (float) m * (nn + m); }
    static float cvol_hex(const LspGeo& g, int m, int nn)
    {
        const float k = (float) (g.lw - g.b_left + g.a_right), q = (float) (g.b_right - g.a_left - g.up);
        return (float) m * nn - (k * k + q * q) / 2;
    }
    static float coef_c(const gspaln_params& P) { return (float) ((P.noll + 1) * 4); }
    static bool is_local(const gspaln_params& P) { return P.local != 0; }
    static bool udh_ok(const gspaln_params& P) { return P.noll == 2; }   // no double-affine Hirschberg pass yet
    // blocks with fewer than 8 query rows: the scalar kernel, if its tables are there
    static bool scalar_ok(const gspaln_ctx* ctx, const gspaln_task& base, const LspGeo& g)
    {
        return !ctx->prm.spj || (ctx->ng_ready && base.int53 && g.b_right - g.b_left < ctx->n_pen);
    }
    static int trivial_score(const gspaln_params& P, const LspGeo& g, int m, int nn)
    {
        // PwdB::GapPenalty / GapExtPen / UnpPenalty (src/aln.h:275-287).  Double affine
        // (Noll == 3): beyond codonk1 = k1 residues the long-gap terms apply; k1 follows from
        // LongGOP = BasicGOP - diffu * k1 with diffu = LongGEP - BasicGEP (src/aln2.cc:109-114).
        const int diffu = P.lgep - P.gep;
        const int k1 = (P.noll == 3 && diffu) ? (P.gop - P.lgop) / diffu : INT_MAX;
        auto ext = [&](int i) { return i > k1 ? P.lgep : P.gep; };
        if (m) return (g.a_exgl || g.a_exgr) ? ext(m) : (m > k1 ? P.lgop + m * P.lgep : P.gop + m * P.gep);
        return (g.b_exgl || g.b_exgr) ? ext(nn) : (nn <= k1 ? nn * P.gep : nn * P.gep + diffu * (nn - k1));
    }
    static void diagonal(const gspaln_params& P, const gspaln_task& base, const LspGeo& g, int (&c4)[4], int& score)
    {
        // diagonalS_ng (src/fwd2s1.cc:1629-1665)
        const int NEVSEL = INT_MIN / 16 * 7;
        const int m = g.a_right - g.a_left, nn = g.b_right - g.b_left;
        const bool LocalL = P.local && g.a_exgl && g.b_exgl, LocalR = P.local && g.a_exgr && g.b_exgr;
        const int dlt = P.local ? 0 : (nn - m);
        const uint8_t* as = dlt < 0 ? base.b : base.a;
        const uint8_t* bs = dlt < 0 ? base.a : base.b;
        const int al = dlt < 0 ? g.b_left : g.a_left, ar = dlt < 0 ? g.b_right : g.a_right;
        const int bl = dlt < 0 ? g.a_left : g.b_left;
        int mL = al, mR = ar, scr = 0, maxh = NEVSEL;
        for (int mm = al, k = 0; mm++ < ar; ++k) {
            const int x = as[al + k], y = bs[bl + k];
            scr += dlt < 0 ? P.simmtx[y * P.simdim + x] : P.simmtx[x * P.simdim + y];
            if (LocalL && scr < 0) { scr = 0; mL = mm; }
            if (LocalR && scr > maxh) { maxh = scr; mR = mm; }
        }
        int r = bl - al;
        if (dlt < 0) r -= dlt;
        c4[0] = mL; c4[1] = mL + r; c4[2] = mR; c4[3] = mR + r;
        score = LocalR ? maxh : scr;
    }
    static bool bad_range(const gspaln_task&, const LspGeo&) { return false; }
    static bool beyond(const gspaln_task&, const LspGeo&) { return false; }     // lengths are not part of the DNA task
    static gspaln_task make_task(const gspaln_task& base, const LspGeo& g, int kind, int n_imd)
    {
        gspaln_task t = base;
        t.kind = kind;
        t.a_left = g.a_left; t.a_right = g.a_right; t.b_left = g.b_left; t.b_right = g.b_right;
        t.a_exgl = g.a_exgl; t.a_exgr = g.a_exgr; t.b_exgl = g.b_exgl; t.b_exgr = g.b_exgr;
        t.lw = g.lw; t.up = g.up;
        t.n_imd = n_imd;
        t.skl_cap = (kind == GSPALN_FORWARD_WIP || kind == GSPALN_FORWARD_NG)
                        ? (g.a_right - g.a_left) + (g.b_right - g.b_left) + 8 : 0;
        return t;
    }
    static int submit(gspaln_ctx* ctx, const gspaln_task* t, int n, gspaln_result* r) { return gspaln_submit(ctx, t, n, r); }
    static int64_t cells(const gspaln_task& t) { return task_cells(t); }
};

}   // namespace

extern "C" int gspaln_lsp(gspaln_ctx* ctx, const gspaln_task* tasks, int n,
                          const gspaln_lsp_opts* opts, gspaln_result* results)
{
    return lsp_driver<LspTraitsS>(ctx, tasks, n, opts, results);
}

// ---------------------------------------------------------------------------
// coalescing queue (include/gspaln.h): one dispatcher thread turns the single-problem calls of
// many host threads into batched gspaln_submit / gspaln_lsp calls (gspaln_host.hpp)
// ---------------------------------------------------------------------------
struct gspaln_queue : gspaln::CoalescingQueue<gspaln_ctx, gspaln_task, gspaln_result, gspaln_lsp_opts> {};

extern "C" {

int gspaln_queue_create(gspaln_queue** out, gspaln_ctx* ctx, int max_batch, int max_wait_us)
{
    if (!out || !ctx) return GSPALN_EINVAL;
    gspaln_queue* q = new gspaln_queue;
    q->ctx = ctx;
    q->submit_fn = gspaln_submit;
    q->lsp_fn = gspaln_lsp;
    q->einval = GSPALN_EINVAL;
    if (max_batch > 0) q->max_batch = max_batch;
    if (max_wait_us >= 0) q->max_wait_us = max_wait_us;
    q->start();
    *out = q;
    return GSPALN_OK;
}

int gspaln_queue_submit(gspaln_queue* q, const gspaln_task* task, gspaln_result* result)
{
    if (!q || !task || !result) return GSPALN_EINVAL;
    return q->submit(task, nullptr, result);
}

int gspaln_queue_submit_lsp(gspaln_queue* q, const gspaln_task* task, const gspaln_lsp_opts* opts,
                            gspaln_result* result)
{
    if (!q || !task || !opts || !result) return GSPALN_EINVAL;
    return q->submit(task, opts, result);
}

int gspaln_queue_stats(const gspaln_queue* q, int64_t* tasks, int64_t* batches)
{
    if (!q) return GSPALN_EINVAL;
    const_cast<gspaln_queue*>(q)->stats(tasks, batches);
    return GSPALN_OK;
}

void gspaln_queue_destroy(gspaln_queue* q)
{
    if (!q) return;
    q->shutdown();
    delete q;
}

}   // extern "C"
