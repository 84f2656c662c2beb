// gspaln_h.cu -- host side of the protein x genome part of the C-ABI (include/gspaln.h,
// gspaln_h_*): device pools, packing (the per-column records the DP rows consume are derived
// here from the caller's SGPT6 table), launches of dp_h1_kernel on the engine's own stream,
// CUDA-event timing.  No CPU implementation behind this API.
#include "../../include/gspaln.h"
#include "gspaln_h1.cuh"
#include "gspaln_h1_udh.cuh"
#include "gspaln_host.hpp"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

using namespace gspaln;

struct gspaln_h_ctx {
    int device = 0;
    int sm_count = 0;
    gspaln_h_params prm;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    DevBuf<DevParamsH> d_prm;
    DevBuf<int2> d_pen;
    int pen_cap = 0;
    size_t smem_bytes = 0;
    DevBuf<DevTaskH> d_tasks;
    DevBuf<int> d_order;
    DevBuf<int> d_ticket;
    DevBuf<unsigned char> d_apool;
    DevBuf<ColH> d_cpool;
    DevBuf<ColEnd> d_epool;
    DevBuf<unsigned> d_band;
    DevBuf<unsigned short> d_trace;
    DevBuf<unsigned char> d_rows;
    DevBuf<int2> d_skl;
    DevBuf<DevResult> d_res;
    DevBuf<int> d_ws;               // per-warp workspace of the Hirschberg pass
    DevBuf<int> d_cpos;
    DevBuf<DevUdhOutH> d_ures;
    PinBuf<int> h_cpos;
    PinBuf<DevUdhOutH> h_ures;
    PinBuf<DevTaskH> h_tasks;
    PinBuf<int> h_order;
    PinBuf<unsigned char> h_apool;
    PinBuf<ColH> h_cpool;
    PinBuf<ColEnd> h_epool;
    PinBuf<int2> h_skl;
    PinBuf<DevResult> h_res;
    int n = 0, n_trace = 0, n_score = 0, n_udh = 0;
    size_t a_bytes = 0, c_elems = 0, band_slab = 0, trace_slab = 0, row_slab = 0, skl_elems = 0;
    size_t ws_slab = 0, cpos_elems = 0;
    int grid_trace = 0, grid_score = 0, grid_run_trace = 0, grid_run_score = 0;
    int grid_udh = 0, grid_run_udh = 0;
    std::vector<int64_t> cells;
    gspaln_timing tim;
    std::string err;
};

namespace {

int fail(gspaln_h_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess)
{
    if (c) {
        c->err = what;
        if (e != cudaSuccess) { c->err += ": "; c->err += cudaGetErrorString(e); }
    }
    return code;
}

#define CKH(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, GSPALN_ECUDA, #call, e_); } while (0)

using KernelH = void (*)(const DevParamsH*, const int2*, const DevTaskH*, const int*, int, int*,
                         const unsigned char*, const ColH*, const ColEnd*, unsigned*, long long,
                         unsigned short*, long long, unsigned char*, long long, int2*, DevResult*);

KernelH kernel_h(bool trace, bool local, bool spj)
{
    static const KernelH tab[8] = {
        dp_h1_kernel<false, false, false>, dp_h1_kernel<false, false, true>,
        dp_h1_kernel<false, true, false>, dp_h1_kernel<false, true, true>,
        dp_h1_kernel<true, false, false>, dp_h1_kernel<true, false, true>,
        dp_h1_kernel<true, true, false>, dp_h1_kernel<true, true, true>,
    };
    return tab[(trace ? 4 : 0) | (local ? 2 : 0) | (spj ? 1 : 0)];
}

using KernelUdhH = void (*)(const DevParamsH*, const int2*, const DevTaskH*, const int*, int, int*,
                            const unsigned char*, const ColH*, const ColEnd*, int*, long long, int*,
                            DevUdhOutH*);

KernelUdhH kernel_udh_h(bool local, bool spj)
{
    static const KernelUdhH tab[4] = {dp_h1_udh_kernel<false, false>, dp_h1_udh_kernel<true, false>,
                                      dp_h1_udh_kernel<false, true>, dp_h1_udh_kernel<true, true>};
    return tab[(local ? 2 : 0) | (spj ? 1 : 0)];
}

int64_t task_cells_h(const gspaln_h_task& t)
{
    // rows m in (a_left, a_right], columns max(3m + lw - 1, b_left) < n <= min(3m + up, b_right)
    int64_t cells = 0;
    for (int m = t.a_left + 1; m <= t.a_right; ++m) {
        const int lo = std::max(3 * m + t.lw - 1, t.b_left);
        const int hi = std::min(3 * m + t.up, t.b_right);
        if (hi > lo) cells += hi - lo;
    }
    return cells;
}

// the column record of genome column c (what enters lane 0 of the reference's shift registers
// at step n == c: src/fwd2h1_wip_simd.h:115-117 (cv), 189-192 (profile), 208-222 / 270-283 (signals))
ColH derive_col(const gspaln_h_task& t, const gspaln_h_params& prm, int c)
{
    ColH o;
    memset(&o, 0, sizeof(o));
    unsigned flags = 0;
    auto sg = [&](int i) -> const gspaln_sgpt6* { return (i >= 0 && i <= t.b_len + 1) ? t.sg + i : nullptr; };
    if (prm.spj && c < t.b_right) {
        const gspaln_sgpt6* s = sg(c);
        const int ipen = prm.ipen;
        if (s) {
            auto put3 = [&](int phase) {
                const gspaln_sgpt6* q = sg(c - phase);
                o.s3[phase + 1] = q ? q->sig3 : 0;
                flags |= 1u << (phase + 1);
            };
            auto put5 = [&](int phase) {
                const gspaln_sgpt6* q = sg(c - phase);
                o.s5[phase + 1] = (short) ((q ? q->sig5 : 0) + ipen);
                flags |= 8u << (phase + 1);
            };
            if (s->phs3 == 2) { put3(-1); put3(1); }
            else if (s->phs3 > -2) put3(s->phs3);
            if (s->phs5 == 2) { put5(-1); put5(1); }
            else if (s->phs5 > -2) put5(s->phs5);
        }
    }
    if (c - 2 >= 0 && c - 2 < t.b_len) o.cv = t.sg[c - 2].sigE;
    o.prof = (c >= t.b_left + 3 && c <= t.b_right + 2) ? (unsigned char) (t.b[c - 2] & 31) : (unsigned char) ZROW;
    o.flags = (unsigned char) flags;
    return o;
}

}   // namespace

extern "C" {

int64_t gspaln_h_task_cells(const gspaln_h_task* t) { return t ? task_cells_h(*t) : 0; }

const char* gspaln_h_last_error(const gspaln_h_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int gspaln_h_create(gspaln_h_ctx** out, const gspaln_h_params* prm, int device)
{
    if (!out || !prm) return GSPALN_EINVAL;
    *out = nullptr;
    if (prm->simdim <= 0 || prm->simdim > ZROW || prm->nquant < 1 || prm->nquant > GSPALN_MAXQUANT ||
        prm->avmch <= 0 || (short) prm->gep > 0 || (short) prm->gw1 > 0 || (short) prm->gw2 > 0 ||
        (short) prm->gw3 > 0)
        return GSPALN_EINVAL;
    for (int j = 0; j < prm->nquant; ++j)
        if ((short) prm->quant_pen[j] > 0) return GSPALN_EINVAL;   // kernels rely on penalties <= 0
    int ndev = gspaln_device_count();
    if (ndev <= 0 || device < 0 || device >= ndev) return GSPALN_ENODEV;
    gspaln_h_ctx* ctx = new gspaln_h_ctx;
    ctx->device = device;
    ctx->prm = *prm;
    memset(&ctx->tim, 0, sizeof(ctx->tim));
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 6 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { gspaln_h_destroy(ctx); return GSPALN_ECUDA; }
    ctx->sm_count = prop.multiProcessorCount;

    DevParamsH P;
    memset(&P, 0, sizeof(P));
    P.g1 = (short) prm->gw1; P.g2 = (short) prm->gw2; P.g3 = (short) prm->gw3; P.ge = (short) prm->gep;
    P.gop = prm->gop; P.gep = prm->gep; P.lgep = prm->lgep; P.codonk1 = prm->codonk1;
    P.gw1 = prm->gw1; P.gw2 = prm->gw2; P.gw3 = prm->gw3;
    P.avmch = prm->avmch; P.local = (prm->lcl & 16) ? 1 : 0; P.spj = prm->spj ? 1 : 0; P.lcl = prm->lcl;
    for (int a = 0; a < prm->simdim && a < ZROW; ++a)
        for (int g = 0; g < prm->simdim && g < ZROW; ++g)
            P.mtxT[g * MTX_LD + a] = (short) prm->simmtx[a * prm->simdim + g];
    // binned intron-length penalty over the length counter (src/fwd2h1_wip_simd.h:226-236):
    // entry h = {penalty, lower clamp}; lengths <= llmt give exactly nevsel.  A second copy
    // that never yields a candidate serves the steps in which no lane carries an acceptor.
    const int mil = (short) prm->llmt;
    int cap = std::max(0, mil);
    for (int j = 0; j + 1 < prm->nquant; ++j) cap = std::max(cap, (int) (short) prm->quant_len[j]);
    cap += 1;
    std::vector<int2> pen(2 * (size_t) (cap + 1));
    for (int h = 0; h <= cap; ++h) {
        int pv = (short) prm->quant_pen[0];
        for (int j = 1; j < prm->nquant; ++j) if (h > (short) prm->quant_len[j - 1]) pv = (short) prm->quant_pen[j];
        const bool valid = h > mil;
        pen[h] = make_int2(valid ? pv : PEN_INVALID, valid ? -32768 : NEV);
        pen[cap + 1 + h] = make_int2(PEN_INVALID, -32768);
    }
    P.pen_cap = cap;
    ctx->pen_cap = cap;
    ctx->smem_bytes = sizeof(RingH) * RINGH * SPPH * WARPS_PER_CTA + sizeof(int2) * pen.size();
    if (ctx->smem_bytes > 100 * 1024) { gspaln_h_destroy(ctx); return GSPALN_EINVAL; }
    if (ctx->d_prm.reserve(1) != cudaSuccess || ctx->d_ticket.reserve(4) != cudaSuccess ||
        ctx->d_pen.reserve(pen.size()) != cudaSuccess ||
        cudaMemcpy(ctx->d_pen.p, pen.data(), sizeof(int2) * pen.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(ctx->d_prm.p, &P, sizeof(P), cudaMemcpyHostToDevice) != cudaSuccess) {
        gspaln_h_destroy(ctx);
        return GSPALN_ENOMEM;
    }
    const void* kt = reinterpret_cast<const void*>(kernel_h(true, P.local, P.spj));
    const void* ks = reinterpret_cast<const void*>(kernel_h(false, P.local, P.spj));
    cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ctx->smem_bytes);
    cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ctx->smem_bytes);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kt, CTA_THREADS, ctx->smem_bytes);
    ctx->grid_trace = std::max(1, occ) * ctx->sm_count;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ks, CTA_THREADS, ctx->smem_bytes);
    ctx->grid_score = std::max(1, occ) * ctx->sm_count;
    {
        const void* ku = reinterpret_cast<const void*>(kernel_udh_h(P.local, P.spj));
        cudaFuncSetAttribute(ku, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ctx->smem_bytes);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ku, CTA_THREADS, ctx->smem_bytes);
        ctx->grid_udh = std::max(1, occ) * ctx->sm_count;
    }
    if (cudaGetLastError() != cudaSuccess) { gspaln_h_destroy(ctx); return GSPALN_ECUDA; }
    *out = ctx;
    return GSPALN_OK;
}

void gspaln_h_destroy(gspaln_h_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx->d_prm.release(); ctx->d_pen.release(); ctx->d_tasks.release(); ctx->d_order.release();
    ctx->d_ticket.release(); ctx->d_apool.release(); ctx->d_cpool.release(); ctx->d_epool.release();
    ctx->d_band.release(); ctx->d_trace.release(); ctx->d_rows.release(); ctx->d_skl.release();
    ctx->d_res.release(); ctx->d_ws.release(); ctx->d_cpos.release(); ctx->d_ures.release();
    ctx->h_cpos.release(); ctx->h_ures.release();
    ctx->h_tasks.release(); ctx->h_order.release(); ctx->h_apool.release(); ctx->h_cpool.release();
    ctx->h_epool.release(); ctx->h_skl.release(); ctx->h_res.release();
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int gspaln_h_upload(gspaln_h_ctx* ctx, const gspaln_h_task* tasks, int n)
{
    if (!ctx || !tasks || n < 0) return GSPALN_EINVAL;
    CKH(cudaSetDevice(ctx->device));
    ctx->n = 0;
    size_t a_bytes = 0, c_elems = 0, band_slab = 0, trace_slab = 0, row_slab = 0, skl_elems = 0;
    size_t ws_slab = 0, cpos_elems = 0;
    ctx->cells.assign(n, 0);
    std::vector<DevTaskH> dt(n);
    int n_trace = 0, n_score = 0, n_udh = 0;
    for (int i = 0; i < n; ++i) {
        const gspaln_h_task& t = tasks[i];
        if (t.a_right < t.a_left || t.b_right < t.b_left || t.up < t.lw || t.a_left < 0 || t.b_left < 0 ||
            t.b_len < t.b_right ||
            (t.kind != GSPALN_FORWARD_WIP && t.kind != GSPALN_SCOREONLY_WIP && t.kind != GSPALN_HIRSCHBERG_WIP) ||
            (t.kind == GSPALN_HIRSCHBERG_WIP && (t.n_imd < 1 || t.a_right - t.a_left < 2)) ||
            !t.a || !t.b || !t.sg)
            return fail(ctx, GSPALN_EINVAL, "bad task");
        DevTaskH& d = dt[i];
        d.kind = t.kind;
        d.a_left = t.a_left; d.a_right = t.a_right; d.b_left = t.b_left; d.b_right = t.b_right;
        d.lw = t.lw; d.up = t.up;
        d.flags = (t.a_exgl & 3) | ((t.a_exgr & 3) << 2) | ((t.b_exgl & 3) << 4) | ((t.b_exgr & 3) << 6);
        d.skl_cap = t.kind == GSPALN_FORWARD_WIP ? std::max(0, t.skl_cap) : 0;
        d.b_len = t.b_len;
        const int mw = t.a_right - t.a_left, nw = t.b_right - t.b_left;
        const int width = t.up - t.lw + 7;
        d.a_off = (long long) a_bytes;      a_bytes += align_up((size_t) mw + 1, 16);
        d.col_off = (long long) c_elems;    c_elems += align_up((size_t) nw + COL_TAIL_H + 2, 4);
        band_slab = std::max(band_slab, align_up((size_t) width + BAND_PAD_H + 8, 32));
        row_slab = std::max(row_slab, 2 * align_up((size_t) width + 16, 64));
        d.skl_off = (long long) skl_elems;
        d.pad1 = 0;
        if (t.kind == GSPALN_FORWARD_WIP) {
            const size_t nstrips = (mw + NELEM - 1) / NELEM;
            trace_slab = std::max(trace_slab, align_up(nstrips * (size_t) (width + TRACE_PAD_H) * NELEM + 64, 128));
            skl_elems += (size_t) d.skl_cap;
            ++n_trace;
        } else if (t.kind == GSPALN_HIRSCHBERG_WIP) {
            d.pad1 = (long long) cpos_elems | ((long long) t.n_imd << 40);
            cpos_elems += (size_t) 10 * (t.n_imd + 1);
            ws_slab = std::max(ws_slab, align_up(4 * ((size_t) width + BAND_PAD_H) + (size_t) t.n_imd * 4 * width + 16, 64));
            ++n_udh;
        } else
            ++n_score;
        ctx->cells[i] = task_cells_h(t);
    }
    if (ctx->h_tasks.reserve(n + 1) != cudaSuccess || ctx->h_order.reserve(n + 1) != cudaSuccess ||
        ctx->h_apool.reserve(a_bytes + 16) != cudaSuccess || ctx->h_cpool.reserve(c_elems + 4) != cudaSuccess ||
        ctx->h_epool.reserve(c_elems + 4) != cudaSuccess ||
        ctx->h_res.reserve(n + 1) != cudaSuccess || ctx->h_skl.reserve(skl_elems + 1) != cudaSuccess ||
        ctx->h_cpos.reserve(cpos_elems + 1) != cudaSuccess || ctx->h_ures.reserve(n + 1) != cudaSuccess)
        return fail(ctx, GSPALN_ENOMEM, "pinned host allocation");
    if (ctx->d_tasks.reserve(n + 1) != cudaSuccess || ctx->d_order.reserve(n + 1) != cudaSuccess ||
        ctx->d_apool.reserve(a_bytes + 16) != cudaSuccess || ctx->d_cpool.reserve(c_elems + 4) != cudaSuccess ||
        ctx->d_epool.reserve(c_elems + 4) != cudaSuccess ||
        ctx->d_skl.reserve(skl_elems + 1) != cudaSuccess || ctx->d_res.reserve(n + 1) != cudaSuccess ||
        ctx->d_cpos.reserve(cpos_elems + 1) != cudaSuccess || ctx->d_ures.reserve(n + 1) != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, GSPALN_ENOMEM, "device allocation");
    }
    {
        auto ctas = [&](int full, int count) {
            return std::max(1, std::min(full, (count + WARPS_PER_CTA - 1) / WARPS_PER_CTA));
        };
        int gt = n_trace ? ctas(ctx->grid_trace, n_trace) : 0;
        int gs = n_score ? ctas(ctx->grid_score, n_score) : 0;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        free_b += ctx->d_trace.cap * 2 + ctx->d_band.cap * sizeof(unsigned) + ctx->d_rows.cap;
        const size_t budget = (size_t) (0.85 * (double) free_b);
        while (gt > 1 && (size_t) gt * WARPS_PER_CTA * (trace_slab * 2 + band_slab * 4 + row_slab) > budget) gt = gt * 3 / 4;
        const size_t warps = (size_t) std::max(gt, gs) * WARPS_PER_CTA;
        if (ctx->d_band.reserve(warps * band_slab + 32) != cudaSuccess ||
            ctx->d_rows.reserve(warps * row_slab + 64) != cudaSuccess ||
            ctx->d_trace.reserve((size_t) gt * WARPS_PER_CTA * trace_slab + 128) != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, GSPALN_ENOMEM, "device workspace allocation");
        }
        ctx->grid_run_trace = gt;
        ctx->grid_run_score = gs;
        const int gu = n_udh ? ctas(ctx->grid_udh, n_udh) : 0;
        if (ctx->d_ws.reserve((size_t) gu * WARPS_PER_CTA * ws_slab + 64) != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, GSPALN_ENOMEM, "device UDH workspace allocation");
        }
        ctx->grid_run_udh = gu;
    }
    {
        auto pack_range = [&](int lo, int hi) {
            for (int i = lo; i < hi; ++i) {
                const gspaln_h_task& t = tasks[i];
                const DevTaskH& d = dt[i];
                const int mw = t.a_right - t.a_left, nw = t.b_right - t.b_left;
                unsigned char* ap = ctx->h_apool.p + d.a_off;
                for (int j = 0; j < mw; ++j) ap[j] = (unsigned char) (t.a[t.a_left + j] & 31);
                ColH* col = ctx->h_cpool.p + d.col_off;
                ColEnd* ce = ctx->h_epool.p + d.col_off;
                for (int j = 0; j <= nw + COL_TAIL_H; ++j) {
                    const int c = t.b_left + j;
                    col[j] = derive_col(t, ctx->prm, c);
                    ColEnd e = {0, 0, 0, 0};
                    if (c <= t.b_len + 1) {
                        const gspaln_sgpt6& s = t.sg[c];
                        e.sigS = s.sigS; e.sigT = s.sigT; e.sigE = s.sigE; e.sig5 = s.sig5;
                    }
                    ce[j] = e;
                }
                ctx->h_tasks.p[i] = d;
            }
        };
        const size_t work = c_elems + a_bytes;
        int nthr = (int) std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 16);
        if (work < (1u << 19) || n < 2 * nthr) nthr = 1;
        if (nthr == 1) pack_range(0, n);
        else {
            std::vector<std::thread> pool;
            size_t acc = 0, per = (work + nthr - 1) / nthr;
            int lo = 0;
            for (int i = 0; i < n; ++i) {
                acc += (size_t) (tasks[i].b_right - tasks[i].b_left) + (tasks[i].a_right - tasks[i].a_left);
                if (acc >= per || i == n - 1) {
                    pool.emplace_back(pack_range, lo, i + 1);
                    lo = i + 1; acc = 0;
                }
            }
            for (auto& th : pool) th.join();
        }
    }
    std::iota(ctx->h_order.p, ctx->h_order.p + n, 0);
    std::stable_sort(ctx->h_order.p, ctx->h_order.p + n,
                     [&](int x, int y) { return ctx->cells[x] > ctx->cells[y]; });
    CKH(cudaEventRecord(ctx->ev[0], ctx->stream));
    CKH(cudaMemcpyAsync(ctx->d_tasks.p, ctx->h_tasks.p, sizeof(DevTaskH) * n, cudaMemcpyHostToDevice, ctx->stream));
    CKH(cudaMemcpyAsync(ctx->d_order.p, ctx->h_order.p, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
    CKH(cudaMemcpyAsync(ctx->d_apool.p, ctx->h_apool.p, a_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CKH(cudaMemcpyAsync(ctx->d_cpool.p, ctx->h_cpool.p, sizeof(ColH) * c_elems, cudaMemcpyHostToDevice, ctx->stream));
    CKH(cudaMemcpyAsync(ctx->d_epool.p, ctx->h_epool.p, sizeof(ColEnd) * c_elems, cudaMemcpyHostToDevice, ctx->stream));
    CKH(cudaEventRecord(ctx->ev[1], ctx->stream));
    CKH(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->tim.h2d_ms = ms;
    ctx->tim.h2d_bytes = (int64_t) (sizeof(DevTaskH) * n + sizeof(int) * n + a_bytes +
                                    (sizeof(ColH) + sizeof(ColEnd)) * c_elems);
    ctx->n = n; ctx->n_trace = n_trace; ctx->n_score = n_score; ctx->n_udh = n_udh;
    ctx->ws_slab = ws_slab; ctx->cpos_elems = cpos_elems;
    ctx->a_bytes = a_bytes; ctx->c_elems = c_elems; ctx->band_slab = band_slab;
    ctx->trace_slab = trace_slab; ctx->row_slab = row_slab; ctx->skl_elems = skl_elems;
    int64_t cells = 0, tb = 0;
    for (int i = 0; i < n; ++i) {
        cells += ctx->cells[i];
        if (tasks[i].kind == GSPALN_FORWARD_WIP) tb += 2 * ctx->cells[i];
    }
    ctx->tim.cells = cells;
    ctx->tim.trace_bytes = tb;
    return GSPALN_OK;
}

int gspaln_h_run(gspaln_h_ctx* ctx)
{
    if (!ctx) return GSPALN_EINVAL;
    CKH(cudaSetDevice(ctx->device));
    const int n = ctx->n;
    int launches = 0;
    CKH(cudaEventRecord(ctx->ev[2], ctx->stream));
    if (n > 0) {
        const bool local = (ctx->prm.lcl & 16) != 0, spj = ctx->prm.spj != 0;
        if (ctx->n_trace) {
            CKH(cudaMemsetAsync(ctx->d_ticket.p, 0, sizeof(int), ctx->stream));
            kernel_h(true, local, spj)<<<ctx->grid_run_trace, CTA_THREADS, ctx->smem_bytes, ctx->stream>>>(
                ctx->d_prm.p, ctx->d_pen.p, ctx->d_tasks.p, ctx->d_order.p, n, ctx->d_ticket.p,
                ctx->d_apool.p, ctx->d_cpool.p, ctx->d_epool.p, ctx->d_band.p, (long long) ctx->band_slab,
                ctx->d_trace.p, (long long) ctx->trace_slab, ctx->d_rows.p, (long long) ctx->row_slab,
                ctx->d_skl.p, ctx->d_res.p);
            ++launches;
        }
        if (ctx->n_score) {
            CKH(cudaMemsetAsync(ctx->d_ticket.p + 1, 0, sizeof(int), ctx->stream));
            kernel_h(false, local, spj)<<<ctx->grid_run_score, CTA_THREADS, ctx->smem_bytes, ctx->stream>>>(
                ctx->d_prm.p, ctx->d_pen.p, ctx->d_tasks.p, ctx->d_order.p, n, ctx->d_ticket.p + 1,
                ctx->d_apool.p, ctx->d_cpool.p, ctx->d_epool.p, ctx->d_band.p, (long long) ctx->band_slab,
                ctx->d_trace.p, (long long) ctx->trace_slab, ctx->d_rows.p, (long long) ctx->row_slab,
                ctx->d_skl.p, ctx->d_res.p);
            ++launches;
        }
        if (ctx->n_udh) {
            CKH(cudaMemsetAsync(ctx->d_ticket.p + 2, 0, sizeof(int), ctx->stream));
            kernel_udh_h(local, spj)<<<ctx->grid_run_udh, CTA_THREADS, ctx->smem_bytes, ctx->stream>>>(
                ctx->d_prm.p, ctx->d_pen.p, ctx->d_tasks.p, ctx->d_order.p, n, ctx->d_ticket.p + 2,
                ctx->d_apool.p, ctx->d_cpool.p, ctx->d_epool.p, ctx->d_ws.p, (long long) ctx->ws_slab,
                ctx->d_cpos.p, ctx->d_ures.p);
            ++launches;
        }
        CKH(cudaGetLastError());
    }
    CKH(cudaEventRecord(ctx->ev[3], ctx->stream));
    CKH(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
    ctx->tim.kernel_ms = ms;
    ctx->tim.launches = launches;
    return GSPALN_OK;
}

int gspaln_h_download(gspaln_h_ctx* ctx, gspaln_result* results)
{
    if (!ctx || (!results && ctx->n)) return GSPALN_EINVAL;
    CKH(cudaSetDevice(ctx->device));
    const int n = ctx->n;
    CKH(cudaEventRecord(ctx->ev[4], ctx->stream));
    if (n) CKH(cudaMemcpyAsync(ctx->h_res.p, ctx->d_res.p, sizeof(DevResult) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->skl_elems)
        CKH(cudaMemcpyAsync(ctx->h_skl.p, ctx->d_skl.p, sizeof(int2) * ctx->skl_elems, cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->n_udh) {
        CKH(cudaMemcpyAsync(ctx->h_ures.p, ctx->d_ures.p, sizeof(DevUdhOutH) * n, cudaMemcpyDeviceToHost, ctx->stream));
        CKH(cudaMemcpyAsync(ctx->h_cpos.p, ctx->d_cpos.p, sizeof(int) * ctx->cpos_elems, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CKH(cudaEventRecord(ctx->ev[5], ctx->stream));
    CKH(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]);
    ctx->tim.d2h_ms = ms;
    ctx->tim.d2h_bytes = (int64_t) (sizeof(DevResult) * n + sizeof(int2) * ctx->skl_elems);
    for (int i = 0; i < n; ++i) {
        const DevResult& r = ctx->h_res.p[i];
        gspaln_result& o = results[i];
        o.score = r.score; o.status = r.status; o.n_skl = r.n_skl; o.reserved = 0;
        o.cells = ctx->cells[i];
        const DevTaskH& d = ctx->h_tasks.p[i];
        if (d.kind == GSPALN_HIRSCHBERG_WIP) {
            const DevUdhOutH& u = ctx->h_ures.p[i];
            o.score = u.score; o.status = u.status; o.n_skl = 0;
            o.ranges[0] = u.a_left; o.ranges[1] = u.a_right; o.ranges[2] = u.b_left; o.ranges[3] = u.b_right;
            const int n_imd = (int) (d.pad1 >> 40);
            if (o.cpos) memcpy(o.cpos, ctx->h_cpos.p + (d.pad1 & ((1ll << 40) - 1)), sizeof(int) * 10 * (size_t) (n_imd + 1));
            continue;
        }
        if (o.skl && d.skl_cap > 0) {
            const int cnt = std::min(r.n_skl, d.skl_cap);
            memcpy(o.skl, ctx->h_skl.p + d.skl_off, sizeof(int2) * (size_t) std::max(0, cnt));
        }
    }
    return GSPALN_OK;
}

int gspaln_h_submit(gspaln_h_ctx* ctx, const gspaln_h_task* tasks, int n, gspaln_result* results)
{
    int rc = gspaln_h_upload(ctx, tasks, n);
    if (rc == GSPALN_OK) rc = gspaln_h_run(ctx);
    if (rc == GSPALN_OK) rc = gspaln_h_download(ctx, results);
    return rc;
}

int gspaln_h_get_timing(const gspaln_h_ctx* ctx, gspaln_timing* out)
{
    if (!ctx || !out) return GSPALN_EINVAL;
    *out = ctx->tim;
    return GSPALN_OK;
}

}   // extern "C"
#include "gspaln_lsp.inl"

namespace {

// protein x genome: Aln2h1::lspH_ng (src/fwd2h1.cc:2134-2230)
struct LspTraitsH {
    using Ctx = gspaln_h_ctx;
    using Task = gspaln_h_task;
    static constexpr int WPAD = 7;
    static void stripe(LspGeo& g, int sh)       // stripe31(), src/aln2.cc:178-199
    {
        if (sh < 0) {
            const int shorter = std::min(g.a_right - g.a_left, g.b_right - g.b_left);
            sh = -sh * shorter / 100;
        }
        sh *= 3;
        int up = g.b_right - 3 * g.a_right;
        int lw = g.b_left - 3 * g.a_left;
        if (up < lw) std::swap(up, lw);
        up += sh; lw -= sh;
        int q;
        if ((q = g.b_right - 3 * g.a_left) < up) up = q;
        if ((q = g.b_left - 3 * g.a_right) > lw) lw = q;
        g.up = up; g.lw = lw;
    }
    static bool small(int m, int nn) { return std::abs(nn - m) < NELEM || m == 1 || nn <= 3; }
    static float cvol(int m, int nn) { return (float) m * (nn + 3 * m); }
    static float coef_c(const gspaln_h_params&) { return 12.f; }     // (Noll + 1) * sizeof(int), Noll == 2
    static bool is_local(const gspaln_h_params& P) { return (P.lcl & 16) != 0; }
    static int trivial_score(const gspaln_h_params& P, const LspGeo& g, int m, int nn)
    {
        auto ext = [&](int i) { return i > P.codonk1 ? P.lgep : P.gep; };
        if (m) return (g.a_exgl || g.a_exgr) ? ext(m) : (m > P.codonk1 ? P.lgop + m * P.lgep : P.gop + m * P.gep);
        // PwdB::UnpPenalty3 (src/aln.h:289-301), i <= codonk1
        return (g.b_exgl || g.b_exgr) ? ext(nn)
                                      : (nn / 3) * P.gep + (nn % 3 == 1 ? P.gape1 : (nn % 3 == 2 ? P.gape2 : 0));
    }
    static void diagonal(const gspaln_h_params& P, const gspaln_h_task& t, const LspGeo& g, int (&c4)[4], int& score)
    {
        // diagonalH_ng (src/fwd2h1.cc:1963-1995): one codon per residue along the only diagonal
        const int NEVSEL = INT_MIN / 16 * 7;
        const bool local = (P.lcl & 16) != 0;
        const bool LocalL = local && g.a_exgl && g.b_exgl, LocalR = local && g.a_exgr && g.b_exgr;
        int scr = 0, maxh = NEVSEL, mL = g.a_left, mR = g.a_right;
        for (int mm = g.a_left, k = 0; mm < g.a_right; ++k) {
            const int col = g.b_left + 1 + 3 * k;
            scr += P.simmtx[(t.a[g.a_left + k] & 31) * P.simdim + (t.b[col] & 31)] + t.sg[col].sigE;
            ++mm;
            if (LocalL && scr < 0) { scr = 0; mL = mm; }
            if (LocalR && scr > maxh) { maxh = scr; mR = mm; }
        }
        c4[0] = mL; c4[1] = 3 * (mL - g.a_left) + g.b_left; c4[2] = mR; c4[3] = 3 * (mR - g.a_left) + g.b_left;
        score = LocalR ? maxh : scr;
    }
    static bool bad_range(const gspaln_h_task& t, const LspGeo& g)     // mimd_postwork, src/fwd2h1.cc:2060-2061
    {
        return g.a_right > t.a_len || g.b_right > t.b_len || g.a_left < 0 || g.b_left < 0;
    }
    static gspaln_h_task make_task(const gspaln_h_task& base, const LspGeo& g, int kind, int n_imd)
    {
        gspaln_h_task t = base;
        t.kind = kind;
        t.a_left = g.a_left; t.a_right = g.a_right; t.b_left = g.b_left; t.b_right = g.b_right;
        t.a_exgl = g.a_exgl; t.a_exgr = g.a_exgr; t.b_exgl = g.b_exgl; t.b_exgr = g.b_exgr;
        t.lw = g.lw; t.up = g.up;
        t.n_imd = n_imd;
        t.skl_cap = kind == GSPALN_FORWARD_WIP ? (g.a_right - g.a_left) + (g.b_right - g.b_left) + 8 : 0;
        return t;
    }
    static int submit(gspaln_h_ctx* ctx, const gspaln_h_task* t, int n, gspaln_result* r) { return gspaln_h_submit(ctx, t, n, r); }
    static int64_t cells(const gspaln_h_task& t) { return task_cells_h(t); }
};

}   // namespace

extern "C" int gspaln_h_lsp(gspaln_h_ctx* ctx, const gspaln_h_task* tasks, int n,
                            const gspaln_lsp_opts* opts, gspaln_result* results)
{
    return lsp_driver<LspTraitsH>(ctx, tasks, n, opts, results);
}
