// gspaln_h.cu -- host side of the protein x genome part of the C-ABI (include/gspaln.h,
// gspaln_h_*): device pools, packing (the per-column records the DP rows consume are derived
// here from the caller's SGPT6 table), launches of dp_h1_kernel on the engine's own stream,
// CUDA-event timing.  No CPU implementation behind this API.
#include "../../include/gspaln.h"
#include "gspaln_h1.cuh"
#include "gspaln_h1_udh.cuh"
#include "gspaln_hng.cuh"
#include "gspaln_hxudh.cuh"
#include "gspaln_host.hpp"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdio>
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
    // scalar kernel (GSPALN_FORWARD_NG): tables + frozen scalars on the device
    DevBuf<short> d_ngtab;              // sig53tab[544] | Penalty(0 .. n_pen - 1)
    DevBuf<unsigned char> d_ngspj;      // split-codon tables + aa2nuc
    DevBuf<int> d_ngmtx;                // simmtx[aa][tron]
    DevBuf<DevNgHParams> d_ngprm;
    int n_pen = 0;
    bool ng_ready = false;
    int ng_noll = 2;                // PwdB::Noll the exact-ILD tables were bound with
    gspaln_h_params prm;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;    // H2D of the chunks that follow the first one
    cudaEvent_t ev_sync[2] = {nullptr, nullptr};
    PinBuf<int> h_marks;
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
                         unsigned short*, long long, unsigned char*, long long, int2*, DevResult*, const int*);

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
                            DevUdhOutH*, const int*);

KernelUdhH kernel_udh_h(bool local, bool spj)
{
    static const KernelUdhH tab[4] = {dp_h1_udh_kernel<false, false>, dp_h1_udh_kernel<true, false>,
                                      dp_h1_udh_kernel<false, true>, dp_h1_udh_kernel<true, true>};
    return tab[(local ? 2 : 0) | (spj ? 1 : 0)];
}

int64_t task_cells_h(const gspaln_h_task& t)
{
    // rows m in (a_left, a_right], columns max(3m + lw - 1, b_left) < n <= min(3m + up, b_right)
    return band_cells(t.a_left, t.a_right, 3, (long long) t.lw - 1, t.b_left, t.up, t.b_right);
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
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ctx->ev_sync[i], cudaEventDisableTiming);
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
    if (ctx->d_prm.reserve(1) != cudaSuccess || ctx->d_ticket.reserve(64) != cudaSuccess ||
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
    ctx->d_ngtab.release(); ctx->d_ngspj.release(); ctx->d_ngmtx.release(); ctx->d_ngprm.release();
    ctx->h_tasks.release(); ctx->h_order.release(); ctx->h_apool.release(); ctx->h_cpool.release();
    ctx->h_epool.release(); ctx->h_skl.release(); ctx->h_res.release();
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    for (auto& e : ctx->ev_sync) if (e) cudaEventDestroy(e);
    ctx->h_marks.release();
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// ---- planning: validation, longest-first order, pool offsets along that order, workspaces, grids
static int plan_batch_h(gspaln_h_ctx* ctx, const gspaln_h_task* tasks, int n)
{
    CKH(cudaSetDevice(ctx->device));
    ctx->n = 0;
    ctx->cells.assign(n, 0);
    for (int i = 0; i < n; ++i) {
        const gspaln_h_task& t = tasks[i];
        if (t.a_right < t.a_left || t.b_right < t.b_left || t.up - t.lw + 7 < 0 || t.a_left < 0 || t.b_left < 0 ||
            t.b_len < t.b_right ||
            (t.kind != GSPALN_FORWARD_WIP && t.kind != GSPALN_SCOREONLY_WIP && t.kind != GSPALN_HIRSCHBERG_WIP) ||
            (t.kind == GSPALN_HIRSCHBERG_WIP && (t.n_imd < 1 || t.a_right - t.a_left < 2)) ||
            !t.a || !t.b || !t.sg) {
            char msg[256];
            snprintf(msg, sizeof(msg), "bad task %d: kind %d a (%d, %d] b (%d, %d] of %d band [%d, %d] n_imd %d",
                     i, t.kind, t.a_left, t.a_right, t.b_left, t.b_right, t.b_len, t.lw, t.up, t.n_imd);
            return fail(ctx, GSPALN_EINVAL, msg);
        }
        ctx->cells[i] = task_cells_h(t);
    }
    if (ctx->h_tasks.reserve(n + 1) != cudaSuccess || ctx->h_order.reserve(n + 1) != cudaSuccess)
        return fail(ctx, GSPALN_ENOMEM, "pinned host allocation");
    std::iota(ctx->h_order.p, ctx->h_order.p + n, 0);
    std::stable_sort(ctx->h_order.p, ctx->h_order.p + n,
                     [&](int x, int y) { return ctx->cells[x] > ctx->cells[y]; });
    size_t a_bytes = 0, c_elems = 0, band_slab = 0, trace_slab = 0, row_slab = 0, skl_elems = 0;
    size_t ws_slab = 0, cpos_elems = 0;
    int n_trace = 0, n_score = 0, n_udh = 0;
    for (int k = 0; k < n; ++k) {
        const int i = ctx->h_order.p[k];
        const gspaln_h_task& t = tasks[i];
        DevTaskH& d = ctx->h_tasks.p[i];
        d.kind = t.kind;
        d.a_left = t.a_left; d.a_right = t.a_right; d.b_left = t.b_left; d.b_right = t.b_right;
        d.lw = t.lw; d.up = t.up;
        d.flags = (t.a_exgl & 3) | ((t.a_exgr & 3) << 2) | ((t.b_exgl & 3) << 4) | ((t.b_exgr & 3) << 6);
        d.skl_cap = t.kind == GSPALN_FORWARD_WIP ? std::max(0, t.skl_cap) : 0;
        d.b_len = t.b_len;
        const int mw = t.a_right - t.a_left, nw = t.b_right - t.b_left;
        const int width = t.up - t.lw + 7;
        // 128-byte granules in every pool (inputs of later problems arrive while earlier ones are read)
        d.a_off = (long long) a_bytes;      a_bytes += align_up((size_t) mw + 1, 128);
        d.col_off = (long long) c_elems;    c_elems += align_up((size_t) nw + COL_TAIL_H + 2, 16);
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
    }
    if (ctx->h_apool.reserve(a_bytes + 16) != cudaSuccess || ctx->h_cpool.reserve(c_elems + 4) != cudaSuccess ||
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

// ---- packing of order[lo .. hi): the per-column records are derived here from SGPT6
static void pack_range_h(gspaln_h_ctx* ctx, const gspaln_h_task* tasks, int lo, int hi)
{
    auto pack_some = [&](int klo, int khi) {
        for (int k = klo; k < khi; ++k) {
            const int i = ctx->h_order.p[k];
            const gspaln_h_task& t = tasks[i];
            const DevTaskH& d = ctx->h_tasks.p[i];
            const int mw = t.a_right - t.a_left, nw = t.b_right - t.b_left;
            unsigned char* ap = ctx->h_apool.p + d.a_off;
            for (int j = 0; j < mw; ++j) ap[j] = (unsigned char) (t.a[t.a_left + j] & 31);
            ColH* col = ctx->h_cpool.p + d.col_off;
            ColEnd* ce = ctx->h_epool.p + d.col_off;
            const gspaln_sgpt6* sg = t.sg;
            const bool spj = ctx->prm.spj != 0;
            const int ipen = ctx->prm.ipen;
            const int c_sig_end = spj ? t.b_right : t.b_left;      // signals for c < b_right only
            for (int j = 0; j <= nw + COL_TAIL_H; ++j) {
                const int c = t.b_left + j;
                // the record of genome column c = what enters lane 0 of the reference's shift
                // registers at step n == c: src/fwd2h1_wip_simd.h:115-117 (cv), 189-192 (profile),
                // 208-222 / 270-283 (signals by splice phase; phs == 2 means both -1 and +1)
                ColH o;
                o.s3[0] = o.s3[1] = o.s3[2] = o.s5[0] = o.s5[1] = o.s5[2] = 0;
                unsigned flags = 0;
                if (c < c_sig_end) {
                    const gspaln_sgpt6& s0 = sg[c];
                    const int p3 = s0.phs3, p5 = s0.phs5;
                    if (p3 > -2 || p5 > -2) {
                        const short m3 = sg[c + 1].sig3, z3 = s0.sig3, q3 = c > 0 ? sg[c - 1].sig3 : (short) 0;
                        const short m5 = sg[c + 1].sig5, z5 = s0.sig5, q5 = c > 0 ? sg[c - 1].sig5 : (short) 0;
                        if (p3 == 2) { o.s3[0] = m3; o.s3[2] = q3; flags |= 5u; }
                        else if (p3 == -1) { o.s3[0] = m3; flags |= 1u; }
                        else if (p3 == 0) { o.s3[1] = z3; flags |= 2u; }
                        else if (p3 == 1) { o.s3[2] = q3; flags |= 4u; }
                        if (p5 == 2) { o.s5[0] = (short) (m5 + ipen); o.s5[2] = (short) (q5 + ipen); flags |= 40u; }
                        else if (p5 == -1) { o.s5[0] = (short) (m5 + ipen); flags |= 8u; }
                        else if (p5 == 0) { o.s5[1] = (short) (z5 + ipen); flags |= 16u; }
                        else if (p5 == 1) { o.s5[2] = (short) (q5 + ipen); flags |= 32u; }
                    }
                }
                o.cv = (c >= 2 && c - 2 < t.b_len) ? sg[c - 2].sigE : (short) 0;
                o.prof = (c >= t.b_left + 3 && c <= t.b_right + 2) ? (unsigned char) (t.b[c - 2] & 31) : (unsigned char) ZROW;
                o.flags = (unsigned char) flags;
                col[j] = o;
                ColEnd e = {0, 0, 0, 0};
                if (c <= t.b_len + 1) {
                    const gspaln_sgpt6& s = sg[c];
                    e.sigS = s.sigS; e.sigT = s.sigT; e.sigE = s.sigE; e.sig5 = s.sig5;
                }
                ce[j] = e;
            }
        }
    };
    size_t work = 0;
    for (int k = lo; k < hi; ++k) {
        const gspaln_h_task& t = tasks[ctx->h_order.p[k]];
        work += (size_t) (t.b_right - t.b_left) + (t.a_right - t.a_left);
    }
    int nthr = (int) std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 16);
    if (work < (1u << 19) || hi - lo < 2 * nthr) nthr = 1;
    if (nthr == 1) { pack_some(lo, hi); return; }
    std::vector<std::thread> pool;
    size_t acc = 0, per = (work + nthr - 1) / nthr;
    int from = lo;
    for (int k = lo; k < hi; ++k) {
        const gspaln_h_task& t = tasks[ctx->h_order.p[k]];
        acc += (size_t) (t.b_right - t.b_left) + (t.a_right - t.a_left);
        if (acc >= per || k == hi - 1) {
            pool.emplace_back(pack_some, from, k + 1);
            from = k + 1; acc = 0;
        }
    }
    for (auto& th : pool) th.join();
}

static void pool_span_h(const gspaln_h_ctx* ctx, const gspaln_h_task* tasks, int lo, int hi, size_t& a0, size_t& a1,
                        size_t& c0, size_t& c1)
{
    const DevTaskH& f = ctx->h_tasks.p[ctx->h_order.p[lo]];
    const int li = ctx->h_order.p[hi - 1];
    const DevTaskH& l = ctx->h_tasks.p[li];
    a0 = (size_t) f.a_off; c0 = (size_t) f.col_off;
    a1 = (size_t) l.a_off + align_up((size_t) (tasks[li].a_right - tasks[li].a_left) + 1, 128);
    c1 = (size_t) l.col_off + align_up((size_t) (tasks[li].b_right - tasks[li].b_left) + COL_TAIL_H + 2, 16);
}

static int launch_all_h(gspaln_h_ctx* ctx, int& launches, const int* ready)
{
    const int n = ctx->n;
    const bool local = (ctx->prm.lcl & 16) != 0, spj = ctx->prm.spj != 0;
    if (ctx->n_trace) {
        kernel_h(true, local, spj)<<<ctx->grid_run_trace, CTA_THREADS, ctx->smem_bytes, ctx->stream>>>(
            ctx->d_prm.p, ctx->d_pen.p, ctx->d_tasks.p, ctx->d_order.p, n, ctx->d_ticket.p,
            ctx->d_apool.p, ctx->d_cpool.p, ctx->d_epool.p, ctx->d_band.p, (long long) ctx->band_slab,
            ctx->d_trace.p, (long long) ctx->trace_slab, ctx->d_rows.p, (long long) ctx->row_slab,
            ctx->d_skl.p, ctx->d_res.p, ready);
        ++launches;
    }
    if (ctx->n_score) {
        kernel_h(false, local, spj)<<<ctx->grid_run_score, CTA_THREADS, ctx->smem_bytes, ctx->stream>>>(
            ctx->d_prm.p, ctx->d_pen.p, ctx->d_tasks.p, ctx->d_order.p, n, ctx->d_ticket.p + 1,
            ctx->d_apool.p, ctx->d_cpool.p, ctx->d_epool.p, ctx->d_band.p, (long long) ctx->band_slab,
            ctx->d_trace.p, (long long) ctx->trace_slab, ctx->d_rows.p, (long long) ctx->row_slab,
            ctx->d_skl.p, ctx->d_res.p, ready);
        ++launches;
    }
    if (ctx->n_udh) {
        kernel_udh_h(local, spj)<<<ctx->grid_run_udh, CTA_THREADS, ctx->smem_bytes, ctx->stream>>>(
            ctx->d_prm.p, ctx->d_pen.p, ctx->d_tasks.p, ctx->d_order.p, n, ctx->d_ticket.p + 2,
            ctx->d_apool.p, ctx->d_cpool.p, ctx->d_epool.p, ctx->d_ws.p, (long long) ctx->ws_slab,
            ctx->d_cpos.p, ctx->d_ures.p, ready);
        ++launches;
    }
    CKH(cudaGetLastError());
    return GSPALN_OK;
}

constexpr int MAX_CHUNKS_H = 16;

int gspaln_h_upload(gspaln_h_ctx* ctx, const gspaln_h_task* tasks, int n)
{
    if (!ctx || !tasks || n < 0) return GSPALN_EINVAL;
    int rc = plan_batch_h(ctx, tasks, n);
    if (rc != GSPALN_OK) return rc;
    if (n) pack_range_h(ctx, tasks, 0, n);
    CKH(cudaEventRecord(ctx->ev[0], ctx->stream));
    CKH(cudaMemcpyAsync(ctx->d_tasks.p, ctx->h_tasks.p, sizeof(DevTaskH) * n, cudaMemcpyHostToDevice, ctx->stream));
    CKH(cudaMemcpyAsync(ctx->d_order.p, ctx->h_order.p, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
    CKH(cudaMemcpyAsync(ctx->d_apool.p, ctx->h_apool.p, ctx->a_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CKH(cudaMemcpyAsync(ctx->d_cpool.p, ctx->h_cpool.p, sizeof(ColH) * ctx->c_elems, cudaMemcpyHostToDevice, ctx->stream));
    CKH(cudaMemcpyAsync(ctx->d_epool.p, ctx->h_epool.p, sizeof(ColEnd) * ctx->c_elems, cudaMemcpyHostToDevice, ctx->stream));
    CKH(cudaEventRecord(ctx->ev[1], ctx->stream));
    CKH(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->tim.h2d_ms = ms;
    ctx->tim.h2d_bytes = (int64_t) (sizeof(DevTaskH) * n + sizeof(int) * n + ctx->a_bytes +
                                    (sizeof(ColH) + sizeof(ColEnd)) * ctx->c_elems);
    return GSPALN_OK;
}

int gspaln_h_run(gspaln_h_ctx* ctx)
{
    if (!ctx) return GSPALN_EINVAL;
    CKH(cudaSetDevice(ctx->device));
    int launches = 0;
    CKH(cudaEventRecord(ctx->ev[2], ctx->stream));
    if (ctx->n > 0) {
        CKH(cudaMemsetAsync(ctx->d_ticket.p, 0, 3 * sizeof(int), ctx->stream));
        int rc = launch_all_h(ctx, launches, nullptr);
        if (rc != GSPALN_OK) return rc;
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

// One-shot path: the batch is streamed in behind the persistent kernels (see gspaln_submit)
static int h_submit_core(gspaln_h_ctx* ctx, const gspaln_h_task* tasks, int n, gspaln_result* results);
static int h_ng_submit(gspaln_h_ctx* ctx, const gspaln_h_task* tasks, int n, gspaln_result* results);

// tasks of kind GSPALN_FORWARD_NG run on the scalar kernel, the rest on the persistent kernels
int gspaln_h_submit(gspaln_h_ctx* ctx, const gspaln_h_task* tasks, int n, gspaln_result* results)
{
    if (!ctx || !tasks || n < 0) return GSPALN_EINVAL;
    int n_ng = 0;
    for (int i = 0; i < n; ++i) n_ng += tasks[i].kind == GSPALN_FORWARD_NG || tasks[i].kind == GSPALN_HIRSCHBERG_NG;
    if (!n_ng) return h_submit_core(ctx, tasks, n, results);
    std::vector<gspaln_h_task> t_ng, t_rest;
    std::vector<gspaln_result> r_ng, r_rest;
    std::vector<int> i_ng, i_rest;
    for (int i = 0; i < n; ++i) {
        const bool ng = tasks[i].kind == GSPALN_FORWARD_NG || tasks[i].kind == GSPALN_HIRSCHBERG_NG;
        (ng ? t_ng : t_rest).push_back(tasks[i]);
        (ng ? r_ng : r_rest).push_back(results[i]);
        (ng ? i_ng : i_rest).push_back(i);
    }
    int rc = GSPALN_OK;
    gspaln_timing tim_rest;
    memset(&tim_rest, 0, sizeof(tim_rest));
    if (!t_rest.empty()) {
        rc = h_submit_core(ctx, t_rest.data(), (int) t_rest.size(), r_rest.data());
        if (rc != GSPALN_OK) return rc;
        tim_rest = ctx->tim;
    }
    {
        // trace-back problems, then Hirschberg passes (two kernels over the same kind of pools)
        std::vector<gspaln_h_task> part[2];
        std::vector<gspaln_result> rpart[2];
        std::vector<size_t> ipart[2];
        for (size_t k = 0; k < t_ng.size(); ++k) {
            const int w = t_ng[k].kind == GSPALN_HIRSCHBERG_NG;
            part[w].push_back(t_ng[k]); rpart[w].push_back(r_ng[k]); ipart[w].push_back(k);
        }
        gspaln_timing acc;
        memset(&acc, 0, sizeof(acc));
        for (int w = 0; w < 2; ++w) {
            if (part[w].empty()) continue;
            rc = h_ng_submit(ctx, part[w].data(), (int) part[w].size(), rpart[w].data());
            if (rc != GSPALN_OK) return rc;
            acc.kernel_ms += ctx->tim.kernel_ms; acc.launches += ctx->tim.launches; acc.cells += ctx->tim.cells;
            for (size_t k = 0; k < ipart[w].size(); ++k) r_ng[ipart[w][k]] = rpart[w][k];
        }
        ctx->tim = acc;
    }
    ctx->tim.kernel_ms += tim_rest.kernel_ms; ctx->tim.h2d_ms += tim_rest.h2d_ms; ctx->tim.d2h_ms += tim_rest.d2h_ms;
    ctx->tim.launches += tim_rest.launches; ctx->tim.h2d_bytes += tim_rest.h2d_bytes;
    ctx->tim.d2h_bytes += tim_rest.d2h_bytes; ctx->tim.cells += tim_rest.cells;
    for (size_t k = 0; k < i_ng.size(); ++k) results[i_ng[k]] = r_ng[k];
    for (size_t k = 0; k < i_rest.size(); ++k) results[i_rest[k]] = r_rest[k];
    return GSPALN_OK;
}

int gspaln_h_set_ng_tables(gspaln_h_ctx* ctx, const int16_t* sig53tab, const int16_t* penalty,
                           int32_t n_penalty, const uint8_t* spj_tabs, int32_t minl,
                           int32_t extragop, int32_t gw3l, int32_t noll)
{
    if (!ctx || !sig53tab || !penalty || !spj_tabs || n_penalty < 1 || (noll != 2 && noll != 3)) return GSPALN_EINVAL;
    CKH(cudaSetDevice(ctx->device));
    const gspaln_h_params& q = ctx->prm;
    const size_t nm = (size_t) q.simdim * q.simdim;
    if (ctx->d_ngtab.reserve(544 + (size_t) n_penalty) != cudaSuccess || ctx->d_ngspj.reserve(800) != cudaSuccess ||
        ctx->d_ngmtx.reserve(nm + 1) != cudaSuccess || ctx->d_ngprm.reserve(1) != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, GSPALN_ENOMEM, "device allocation");
    }
    CKH(cudaMemcpy(ctx->d_ngtab.p, sig53tab, 544 * sizeof(short), cudaMemcpyHostToDevice));
    CKH(cudaMemcpy(ctx->d_ngtab.p + 544, penalty, (size_t) n_penalty * sizeof(short), cudaMemcpyHostToDevice));
    CKH(cudaMemcpy(ctx->d_ngspj.p, spj_tabs, 796, cudaMemcpyHostToDevice));
    CKH(cudaMemcpy(ctx->d_ngmtx.p, q.simmtx, nm * sizeof(int), cudaMemcpyHostToDevice));
    DevNgHParams P;
    memset(&P, 0, sizeof(P));
    P.gop = q.gop; P.gep = q.gep; P.lgop = q.lgop; P.lgep = q.lgep; P.codonk1 = q.codonk1;
    P.gw1 = q.gw1; P.gw2 = q.gw2; P.gw3 = q.gw3; P.gw3l = gw3l; P.gape1 = q.gape1; P.gape2 = q.gape2;
    P.extragop = extragop; P.local = (q.lcl & 16) ? 1 : 0; P.spj = q.spj ? 1 : 0; P.noll = noll; P.minl = minl;
    P.lcl2 = (q.lcl & 2) ? 1 : 0;
    ctx->ng_noll = noll;
    P.simdim = q.simdim; P.n_penalty = n_penalty;
    P.mtx = ctx->d_ngmtx.p; P.penalty = ctx->d_ngtab.p + 544; P.sig53tab = ctx->d_ngtab.p; P.spj_tabs = ctx->d_ngspj.p;
    CKH(cudaMemcpy(ctx->d_ngprm.p, &P, sizeof(P), cudaMemcpyHostToDevice));
    ctx->n_pen = n_penalty;
    ctx->ng_ready = true;
    return GSPALN_OK;
}

// the exact-ILD kernel (gspaln_hng.cuh): one warp per problem, raw inputs copied per problem with a margin
static int h_ng_submit(gspaln_h_ctx* ctx, const gspaln_h_task* tasks, int n, gspaln_result* results)
{
    if (!ctx->ng_ready) return fail(ctx, GSPALN_EINVAL, "GSPALN_FORWARD_NG needs gspaln_h_set_ng_tables");
    CKH(cudaSetDevice(ctx->device));
    const bool udh = n > 0 && tasks[0].kind == GSPALN_HIRSCHBERG_NG;    // (the caller sends one kind at a time)
    std::vector<DevNgHTask> dt(n);
    std::vector<DevUdhHTask> du(udh ? n : 0);
    size_t cpos_elems = 0;
    int n_wide = 0;
    std::vector<unsigned char> apool, bpool;
    std::vector<short> sgpool;
    std::vector<unsigned short> ipool;
    std::vector<int> cippool;       // Cip_score words of the tasks that carry them
    size_t skl_elems = 0, work_bytes = 0;
    int64_t cells_total = 0;
    for (int i = 0; i < n; ++i) {
        const gspaln_h_task& t = tasks[i];
        if (!t.a || !t.b || !t.sg || !t.int53 || t.a_right < t.a_left || t.b_right < t.b_left ||
            t.up - t.lw + 7 < 0 || t.b_right - t.b_left >= ctx->n_pen ||
            (udh && (t.n_imd < 1 || t.a_right - t.a_left < 2)))
            return fail(ctx, GSPALN_EINVAL, "bad GSPALN_FORWARD_NG / GSPALN_HIRSCHBERG_NG task");
        DevNgHTask& d = dt[i];
        d.a_left = t.a_left; d.a_right = t.a_right; d.b_left = t.b_left; d.b_right = t.b_right;
        d.lw = t.lw; d.up = t.up;
        d.a_exgl = t.a_exgl; d.a_exgr = t.a_exgr; d.b_exgl = t.b_exgl; d.b_exgr = t.b_exgr;
        d.skl_cap = std::max(0, t.skl_cap);
        d.wide = t.a_right - t.a_left >= HNG_WIDE_ROWS; d.pad_ = 0;      // a CTA of warps per problem
        n_wide += d.wide;
        const int width = t.up - t.lw + 7;
        const int64_t cells = gspaln_h_task_cells(&t);
        cells_total += cells;
        d.rec_cap = (int) std::min<int64_t>(4 * cells + 4 * width + 64 + (32 * HNG_WIDE + 1) * HNG_CHUNK, INT_MAX / 4);
        // query residues a_left - 1 .. a_right, genome columns b_left - 4 .. b_right + 4 (zeros outside
        // the sequences, as the terminal residues of the reference's arrays)
        d.a_lo = t.a_left - 1; d.a_off = (long long) apool.size();
        for (int p = d.a_lo; p <= t.a_right; ++p) apool.push_back(p >= 0 && p < t.a_len ? t.a[p] : 0);
        d.b_lo = t.b_left - 4; d.b_off = (long long) bpool.size(); d.sg_off = (long long) ipool.size();
        for (int p = d.b_lo; p <= t.b_right + 4; ++p) {
            bpool.push_back(p >= 0 && p < t.b_len ? t.b[p] : 0);
            const bool in = p >= 0 && p <= t.b_len + 1;
            const gspaln_sgpt6 z = {0, 0, 0, 0, 0, 0, -2, -2};
            const gspaln_sgpt6& g = in ? t.sg[p] : z;
            const short rec[8] = {g.sig5, g.sig3, g.sigS, g.sigT, g.sigE, g.sigI, g.phs5, g.phs3};
            sgpool.insert(sgpool.end(), rec, rec + 8);
            ipool.push_back(in ? t.int53[p] : 0);
        }
        d.cip_off = -1;
        if (t.cip) {
            // coding positions 3 a_left - 1 .. 3 a_right + 1 (3 m - phase of every row and splice phase)
            d.cip_off = (long long) cippool.size();
            for (int c = 3 * t.a_left - 1; c <= 3 * t.a_right + 1; ++c) cippool.push_back(c >= 0 ? t.cip[c] : 0);
        }
        d.skl_off = (long long) skl_elems; skl_elems += (size_t) d.skl_cap;
        d.work_off = (long long) work_bytes;
        if (udh) {
            // band rows of 32-byte cells + hlnk | vlnk | lwrb | uprb (noll x width ints each) per intermediate row
            work_bytes += align_up((size_t) 3 * (width + 8) * sizeof(HuSlot) +
                                   (size_t) t.n_imd * 4 * ctx->ng_noll * (size_t) width * sizeof(int) + 64, 32);
            du[i].g = d;
            du[i].n_req = t.n_imd; du[i].pad = 0;
            du[i].cpos_off = (long long) cpos_elems;
            cpos_elems += (size_t) 10 * (t.n_imd + 1);
        } else
            work_bytes += align_up((size_t) 3 * (width + 8) * sizeof(HCell) + (size_t) d.rec_cap * 12 + 16, 16);
    }
    unsigned char *d_a = nullptr, *d_b = nullptr, *d_work = nullptr;
    short* d_sg = nullptr; unsigned short* d_i = nullptr; DevNgHTask* d_t = nullptr; int* d_cip = nullptr;
    int2* d_skl = nullptr; DevResult* d_res = nullptr; int* d_tick = nullptr;
    DevUdhHTask* d_tu = nullptr; int* d_cpos = nullptr; DevUdhOut* d_ures = nullptr;
    auto freeall = [&] {
        cudaFree(d_a); cudaFree(d_b); cudaFree(d_work); cudaFree(d_sg); cudaFree(d_i); cudaFree(d_t); cudaFree(d_cip);
        cudaFree(d_skl); cudaFree(d_res); cudaFree(d_tick); cudaFree(d_tu); cudaFree(d_cpos); cudaFree(d_ures);
    };
    cudaError_t e = cudaSuccess;
    auto up = [&](void** dp, const void* hp, size_t bytes) {
        if (e != cudaSuccess) return;
        e = cudaMalloc(dp, bytes + 16);
        if (e == cudaSuccess && bytes) e = cudaMemcpy(*dp, hp, bytes, cudaMemcpyHostToDevice);
    };
    up((void**) &d_a, apool.data(), apool.size());
    up((void**) &d_b, bpool.data(), bpool.size());
    up((void**) &d_sg, sgpool.data(), sgpool.size() * sizeof(short));
    up((void**) &d_i, ipool.data(), ipool.size() * sizeof(unsigned short));
    if (udh) up((void**) &d_tu, du.data(), du.size() * sizeof(DevUdhHTask));
    else up((void**) &d_t, dt.data(), dt.size() * sizeof(DevNgHTask));
    up((void**) &d_cip, cippool.data(), cippool.size() * sizeof(int));
    if (udh && e == cudaSuccess) e = cudaMalloc((void**) &d_cpos, (cpos_elems + 1) * sizeof(int));
    if (udh && e == cudaSuccess) e = cudaMalloc((void**) &d_ures, (size_t) (n + 1) * sizeof(DevUdhOut));
    if (e == cudaSuccess) e = cudaMalloc((void**) &d_work, work_bytes + 16);
    if (e == cudaSuccess) e = cudaMalloc((void**) &d_skl, (skl_elems + 1) * sizeof(int2));
    if (e == cudaSuccess) e = cudaMalloc((void**) &d_res, (size_t) (n + 1) * sizeof(DevResult));
    if (e == cudaSuccess) e = cudaMalloc((void**) &d_tick, 2 * sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(d_tick, 0, 2 * sizeof(int));
    if (e != cudaSuccess) { freeall(); cudaGetLastError(); return fail(ctx, GSPALN_ENOMEM, "scalar kernel buffers", e); }
    cudaEventRecord(ctx->ev[2], ctx->stream);
    // two classes by query length: one warp per problem (HNG_WARPS problems per CTA), or a CTA of
    // HNG_WIDE warps per problem; every kernel walks the whole task list and takes its class
    const int n_thin = n - n_wide;
    const int grid = std::max(1, std::min((n_thin + HNG_WARPS - 1) / HNG_WARPS, 4 * ctx->sm_count));
    const int grid_w = std::max(1, std::min(n_wide, 2 * ctx->sm_count));
    int launches = 0;
    if (udh) {
        if (n_wide) { dp_hxudh_kernel<HNG_WIDE><<<grid_w, 32 * HNG_WIDE, 0, ctx->stream>>>(ctx->d_ngprm.p, d_tu, n, d_tick + 1, d_a, d_b, d_sg, d_i, d_cip, d_work, d_cpos, d_ures); ++launches; }
        if (n_thin) { dp_hxudh_kernel<1><<<grid, HNG_THREADS, 0, ctx->stream>>>(ctx->d_ngprm.p, d_tu, n, d_tick, d_a, d_b, d_sg, d_i, d_cip, d_work, d_cpos, d_ures); ++launches; }
    } else {
        if (n_wide) { dp_hxild_kernel<HNG_WIDE><<<grid_w, 32 * HNG_WIDE, 0, ctx->stream>>>(ctx->d_ngprm.p, d_t, n, d_tick + 1, d_a, d_b, d_sg, d_i, d_cip, d_work, d_skl, d_res); ++launches; }
        if (n_thin) { dp_hxild_kernel<1><<<grid, HNG_THREADS, 0, ctx->stream>>>(ctx->d_ngprm.p, d_t, n, d_tick, d_a, d_b, d_sg, d_i, d_cip, d_work, d_skl, d_res); ++launches; }
    }
    cudaEventRecord(ctx->ev[3], ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (udh) {
        std::vector<DevUdhOut> hu(n);
        std::vector<int> hc(cpos_elems + 1);
        if (e == cudaSuccess) e = cudaMemcpy(hu.data(), d_ures, (size_t) n * sizeof(DevUdhOut), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(hc.data(), d_cpos, cpos_elems * sizeof(int), cudaMemcpyDeviceToHost);
        float ums = 0;
        cudaEventElapsedTime(&ums, ctx->ev[2], ctx->ev[3]);
        freeall();
        if (e != cudaSuccess) return fail(ctx, GSPALN_ECUDA, "scalar Hirschberg kernel", e);
        memset(&ctx->tim, 0, sizeof(ctx->tim));
        ctx->tim.kernel_ms = ums; ctx->tim.launches = launches; ctx->tim.cells = cells_total;
        for (int i = 0; i < n; ++i) {
            gspaln_result& o = results[i];
            o.score = hu[i].score; o.status = hu[i].status; o.n_skl = 0; o.reserved = 0;
            o.cells = gspaln_h_task_cells(&tasks[i]);
            o.ranges[0] = hu[i].a_left; o.ranges[1] = hu[i].a_right; o.ranges[2] = hu[i].b_left; o.ranges[3] = hu[i].b_right;
            if (o.cpos) memcpy(o.cpos, hc.data() + du[i].cpos_off, sizeof(int) * 10 * (size_t) (tasks[i].n_imd + 1));
        }
        return GSPALN_OK;
    }
    std::vector<DevResult> hres(n);
    std::vector<int2> hskl(skl_elems + 1);
    if (e == cudaSuccess) e = cudaMemcpy(hres.data(), d_res, (size_t) n * sizeof(DevResult), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && skl_elems) e = cudaMemcpy(hskl.data(), d_skl, skl_elems * sizeof(int2), cudaMemcpyDeviceToHost);
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
    freeall();
    if (e != cudaSuccess) return fail(ctx, GSPALN_ECUDA, "scalar kernel", e);
    memset(&ctx->tim, 0, sizeof(ctx->tim));
    ctx->tim.kernel_ms = ms; ctx->tim.launches = launches; ctx->tim.cells = cells_total;
    for (int i = 0; i < n; ++i) {
        gspaln_result& o = results[i];
        o.score = hres[i].score; o.status = hres[i].status; o.n_skl = hres[i].n_skl; o.reserved = 0;
        o.cells = gspaln_h_task_cells(&tasks[i]);
        if (o.skl && dt[i].skl_cap > 0)
            memcpy(o.skl, hskl.data() + dt[i].skl_off, sizeof(int2) * (size_t) std::max(0, std::min(o.n_skl, dt[i].skl_cap)));
    }
    return GSPALN_OK;
}

static int h_submit_core(gspaln_h_ctx* ctx, const gspaln_h_task* tasks, int n, gspaln_result* results)
{
    if (!ctx || !tasks || n < 0) return GSPALN_EINVAL;
    int rc = plan_batch_h(ctx, tasks, n);
    if (rc != GSPALN_OK) return rc;
    int bounds[MAX_CHUNKS_H + 1];
    int nchunks = 1;
    bounds[0] = 0; bounds[1] = n;
    const size_t total = ctx->c_elems;
    // GSPALN_NO_STREAM=1: one chunk (profilers serialise the copy stream behind the running kernel,
    // which would leave the kernel waiting for its watermark until the in-kernel time-out)
    static const bool no_stream = getenv("GSPALN_NO_STREAM") != nullptr;
    if (!no_stream && n >= 256 && total >= (2u << 20)) {
        nchunks = 0;
        size_t acc = 0;
        const int want = (int) std::min<size_t>(MAX_CHUNKS_H, 2 + total / (2u << 20));
        size_t next = total / (2 * (size_t) want);
        for (int k = 0; k < n; ++k) {
            const gspaln_h_task& t = tasks[ctx->h_order.p[k]];
            acc += (size_t) (t.b_right - t.b_left) + COL_TAIL_H + 2;
            if (acc >= next && nchunks + 1 < want && k + 1 < n) {
                bounds[++nchunks] = k + 1;
                next = acc + (total - acc) / (size_t) (want - nchunks);
            }
        }
        bounds[++nchunks] = n;
    }
    if (ctx->h_marks.reserve(MAX_CHUNKS_H + 1) != cudaSuccess) return fail(ctx, GSPALN_ENOMEM, "pinned host allocation");
    int* d_ready = ctx->d_ticket.p + 32;
    CKH(cudaMemsetAsync(ctx->d_ticket.p, 0, 40 * sizeof(int), ctx->stream));
    CKH(cudaEventRecord(ctx->ev[0], ctx->stream));
    CKH(cudaMemcpyAsync(ctx->d_tasks.p, ctx->h_tasks.p, sizeof(DevTaskH) * n, cudaMemcpyHostToDevice, ctx->stream));
    CKH(cudaMemcpyAsync(ctx->d_order.p, ctx->h_order.p, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
    int launches = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int lo = bounds[c], hi = bounds[c + 1];
        if (hi <= lo) continue;
        pack_range_h(ctx, tasks, lo, hi);
        size_t a0, a1, c0, c1;
        pool_span_h(ctx, tasks, lo, hi, a0, a1, c0, c1);
        cudaStream_t st = c == 0 ? ctx->stream : ctx->copy_stream;
        CKH(cudaMemcpyAsync(ctx->d_apool.p + a0, ctx->h_apool.p + a0, a1 - a0, cudaMemcpyHostToDevice, st));
        CKH(cudaMemcpyAsync(ctx->d_cpool.p + c0, ctx->h_cpool.p + c0, sizeof(ColH) * (c1 - c0), cudaMemcpyHostToDevice, st));
        CKH(cudaMemcpyAsync(ctx->d_epool.p + c0, ctx->h_epool.p + c0, sizeof(ColEnd) * (c1 - c0), cudaMemcpyHostToDevice, st));
        ctx->h_marks.p[c] = hi;
        CKH(cudaMemcpyAsync(d_ready, ctx->h_marks.p + c, sizeof(int), cudaMemcpyHostToDevice, st));
        if (c == 0) {
            CKH(cudaEventRecord(ctx->ev[1], ctx->stream));
            CKH(cudaEventRecord(ctx->ev_sync[0], ctx->stream));
            CKH(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_sync[0], 0));
            CKH(cudaEventRecord(ctx->ev[2], ctx->stream));
            rc = launch_all_h(ctx, launches, nchunks > 1 ? d_ready : nullptr);
            if (rc != GSPALN_OK) return rc;
        }
    }
    if (n == 0) { CKH(cudaEventRecord(ctx->ev[1], ctx->stream)); CKH(cudaEventRecord(ctx->ev[2], ctx->stream)); }
    CKH(cudaEventRecord(ctx->ev[3], ctx->stream));
    CKH(cudaStreamSynchronize(ctx->copy_stream));
    CKH(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->tim.h2d_ms = ms;
    cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
    ctx->tim.kernel_ms = ms;
    ctx->tim.launches = launches;
    ctx->tim.h2d_bytes = (int64_t) (sizeof(DevTaskH) * n + sizeof(int) * n + ctx->a_bytes +
                                    (sizeof(ColH) + sizeof(ColEnd)) * ctx->c_elems);
    return gspaln_h_download(ctx, results);
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
    static constexpr bool SCALAR_MODE = true;               // -A0: forwardH_ng + hirschbergH_ng on the device
    static constexpr int KIND_SCALAR_UDH = GSPALN_HIRSCHBERG_NG;
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
    static float cvol_hex(const LspGeo& g, int m, int nn)
    {
        const float k = (float) (g.lw - g.b_left + 3 * g.a_right), q = (float) (g.b_right - 3 * g.a_left - g.up);
        return (float) m * nn - (k * k + q * q) / 6;
    }
    static float coef_c(const gspaln_h_params&) { return 12.f; }     // (Noll + 1) * sizeof(int), Noll == 2
    static bool is_local(const gspaln_h_params& P) { return (P.lcl & 16) != 0; }
    static bool udh_ok(const gspaln_h_params&) { return true; }
    // no scalar forwardH_ng on the device: blocks with fewer than 8 rows stay unsupported
    static bool scalar_ok(const gspaln_h_ctx* ctx, const gspaln_h_task& base, const LspGeo& g)
    {
        return ctx->ng_ready && base.int53 && g.b_right - g.b_left < ctx->n_pen;
    }
    static int trivial_score(const gspaln_h_params& P, const LspGeo& g, int m, int nn)
    {
        auto ext = [&](int i) { return i > P.codonk1 ? P.lgep : P.gep; };
        if (m) return (g.a_exgl || g.a_exgr) ? ext(m) : (m > P.codonk1 ? P.lgop + m * P.lgep : P.gop + m * P.gep);
        // PwdB::UnpPenalty3 (src/aln.h:290-301); beyond codonk1 (= 3 k1 nt) the long-gap slope
        // enters as -diffu (d - k1), diffu = LongGEP - BasicGEP
        if (g.b_exgl || g.b_exgr) return ext(nn);
        const int d = nn / 3;
        const int unp = d * P.gep + (nn % 3 == 1 ? P.gape1 : (nn % 3 == 2 ? P.gape2 : 0));
        return nn <= P.codonk1 ? unp : unp - (P.lgep - P.gep) * (d - P.codonk1 / 3);
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
    static bool beyond(const gspaln_h_task& t, const LspGeo& g) { return g.b_right > t.b_len || g.a_right > t.a_len; }
    static gspaln_h_task make_task(const gspaln_h_task& base, const LspGeo& g, int kind, int n_imd)
    {
        gspaln_h_task t = base;
        t.kind = kind;
        t.a_left = g.a_left; t.a_right = g.a_right; t.b_left = g.b_left; t.b_right = g.b_right;
        t.a_exgl = g.a_exgl; t.a_exgr = g.a_exgr; t.b_exgl = g.b_exgl; t.b_exgr = g.b_exgr;
        t.lw = g.lw; t.up = g.up;
        t.n_imd = n_imd;
        t.skl_cap = (kind == GSPALN_FORWARD_WIP || kind == GSPALN_FORWARD_NG) ? (g.a_right - g.a_left) + (g.b_right - g.b_left) + 8 : 0;
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

// coalescing queue of the protein path (same dispatcher as the DNA one, gspaln_host.hpp)
struct gspaln_h_queue : gspaln::CoalescingQueue<gspaln_h_ctx, gspaln_h_task, gspaln_result, gspaln_lsp_opts> {};

extern "C" {

int gspaln_h_queue_create(gspaln_h_queue** out, gspaln_h_ctx* ctx, int max_batch, int max_wait_us)
{
    if (!out || !ctx) return GSPALN_EINVAL;
    gspaln_h_queue* q = new gspaln_h_queue;
    q->ctx = ctx;
    q->submit_fn = gspaln_h_submit;
    q->lsp_fn = gspaln_h_lsp;
    q->einval = GSPALN_EINVAL;
    if (max_batch > 0) q->max_batch = max_batch;
    if (max_wait_us >= 0) q->max_wait_us = max_wait_us;
    q->start();
    *out = q;
    return GSPALN_OK;
}

int gspaln_h_queue_submit(gspaln_h_queue* q, const gspaln_h_task* task, gspaln_result* result)
{
    if (!q || !task || !result) return GSPALN_EINVAL;
    return q->submit(task, nullptr, result);
}

int gspaln_h_queue_submit_lsp(gspaln_h_queue* q, const gspaln_h_task* task, const gspaln_lsp_opts* opts,
                              gspaln_result* result)
{
    if (!q || !task || !opts || !result) return GSPALN_EINVAL;
    return q->submit(task, opts, result);
}

int gspaln_h_queue_stats(const gspaln_h_queue* q, int64_t* tasks, int64_t* batches)
{
    if (!q) return GSPALN_EINVAL;
    const_cast<gspaln_h_queue*>(q)->stats(tasks, batches);
    return GSPALN_OK;
}

void gspaln_h_queue_destroy(gspaln_h_queue* q)
{
    if (!q) return;
    q->shutdown();
    delete q;
}

}   // extern "C"
