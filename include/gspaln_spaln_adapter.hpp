// gspaln_spaln_adapter.hpp -- header-only adapter that lets the reference (ogotoh/spaln)
// call libgspaln at its SimdAln2s1 seam.  It is compiled INSIDE a Spaln translation unit:
// it expects the reference's own headers (aln.h -> seq.h, codepot.h, mfile.h) to be included
// first and uses their types (Seq, PwdB, WINDOW, SKL, Mfile, IntronPrm, algmode).
//
// Replaces (all paths relative to the reference tree):
//   SimdAln2s1 ctor + forwardS1_wip(Mfile*)   src/fwd2s1_simd.h:191-333, src/fwd2s1_wip_simd.h:233-474
//   SimdAln2s1 ctor + scoreonlyS1_wip()       src/fwd2s1_wip_simd.h:42-231
//   Aln2s1::lspS_ng(wdw) as a whole           src/fwd2s1.cc:1801-1897 (driver: trace-back vs
//                                             Hirschberg dispatch, post-work, scalar small blocks)
// See INTEGRATION.md for the three-line patch of src/fwd2s1.cc.
#ifndef GSPALN_SPALN_ADAPTER_HPP
#define GSPALN_SPALN_ADAPTER_HPP

#include "gspaln.h"

#include <cstdlib>
#include <vector>

namespace gspaln {

class SpalnEngine {
    gspaln_ctx* ctx_ = nullptr;
    std::vector<short> sig5_, sig3_;
    std::vector<unsigned short> int53_;
    std::vector<int> skl_;

    static void die(const char* what, int rc, const gspaln_ctx* c)
    {
        // the reference's own error convention: fatal() == message + exit(1) (src/adddef.h:173-182)
        fatal("gspaln: %s failed (%d): %s\n", what, rc, c ? gspaln_last_error(c) : "");
    }

    void fill(gspaln_task& t, const Seq** seqs, const WINDOW& wdw, int kind)
    {
        const Seq* a = seqs[0];
        const Seq* b = seqs[1];
        t.kind = kind;
        t.int53 = 0;    // only the scalar kernel (GSPALN_FORWARD_NG) reads the INT53 array
        t.a = a->at(0);
        t.b = b->at(0);
        t.a_left = a->left; t.a_right = a->right;
        t.b_left = b->left; t.b_right = b->right;
        t.a_exgl = a->inex.exgl; t.a_exgr = a->inex.exgr;
        t.b_exgl = b->inex.exgl; t.b_exgr = b->inex.exgr;
        t.lw = wdw.lw; t.up = wdw.up;
        // Exinon::data_n is an array of {short sig5, sig3; char phs5, phs3}; the kernels take
        // the two signal columns (src/codepot.h:27-32,104)
        const int n = b->right + 2;
        sig5_.assign(n, 0); sig3_.assign(n, 0);
        if (b->inex.intr && b->exin)
            for (int i = b->left; i <= b->right; ++i) {
                const SGPT2* g = b->exin->score_n(i);
                sig5_[i] = g->sig5; sig3_[i] = g->sig3;
            }
        t.sig5 = sig5_.data(); t.sig3 = sig3_.data();
        t.skl_cap = 0;
    }

public:
    // freezes the globals the reference kernels read into gspaln_params
    explicit SpalnEngine(const PwdB* pwd, int device = 0, bool spliced = true)
    {
        gspaln_params p = gspaln_params();
        p.gop = pwd->BasicGOP; p.gep = pwd->BasicGEP;
        p.lgop = pwd->LongGOP; p.lgep = pwd->LongGEP;
        p.noll = pwd->Noll;
        p.ipen = (spliced && pwd->IntPen) ? pwd->IntPen->Penalty() : 0;
        p.llmt = IntronPrm.llmt;
        p.nquant = IntronPrm.nquant;
        for (int j = 0; j < p.nquant && j < GSPALN_MAXQUANT && pwd->IntPen && pwd->IntPen->qm; ++j) {
            p.quant_len[j] = pwd->IntPen->qm[j].len;
            p.quant_pen[j] = pwd->IntPen->qm[j].pen;
        }
        p.avmch = (int) pwd->simmtx->AvTrc();
        p.local = (algmode.lcl & 16) ? 1 : 0;
        p.spj = spliced ? 1 : 0;
        p.simdim = pwd->simmtx->dim;
        p.gappen1 = pwd->GapPenalty(1);
        for (int q = 0; q < p.simdim; ++q)
            for (int g = 0; g < p.simdim; ++g)
                p.simmtx[q * p.simdim + g] = pwd->simmtx->mtx[q][g];
        int rc = gspaln_create(&ctx_, &p, device);
        if (rc != GSPALN_OK) die("gspaln_create", rc, ctx_);
    }
    ~SpalnEngine() { gspaln_destroy(ctx_); }
    SpalnEngine(const SpalnEngine&) = delete;
    SpalnEngine& operator=(const SpalnEngine&) = delete;

    // == SimdAln2s1(seqs, pwd, wdw, spjcs, cip, 1).forwardS1_wip(mfd)
    VTYPE forwardS1_wip(const Seq** seqs, const WINDOW& wdw, Mfile* mfd)
    {
        gspaln_task t;
        gspaln_result r;
        fill(t, seqs, wdw, GSPALN_FORWARD_WIP);
        int cap = (t.a_right - t.a_left) + (t.b_right - t.b_left) + 8;
        for (;;) {
            skl_.assign(2 * (size_t) cap, 0);
            t.skl_cap = cap;
            r.skl = skl_.data();
            int rc = gspaln_submit(ctx_, &t, 1, &r);
            if (rc != GSPALN_OK) die("gspaln_submit", rc, ctx_);
            if (r.status != GSPALN_ST_SKL_OVERFLOW) break;
            cap = r.n_skl + 8;
        }
        if (r.status == GSPALN_ST_BAD_TRACE) fatal("Unexpected dir\n");     // src/rhomb_coord.h:216
        for (int i = 0; i < r.n_skl; ++i) {
            SKL wsk = {skl_[2 * i], skl_[2 * i + 1]};
            mfd->write((UPTR) &wsk);
        }
        return (VTYPE) r.score;
    }

    // Tables of the scalar exact-ILD kernel (the reference runs Aln2s1::forwardS_ng for blocks with
    // fewer than 8 query rows, src/fwd2s1.cc:1676): sig53tab = Exinon::sig53tab[0] (544 shorts),
    // the length penalty is evaluated here with the reference's own IntronPenalty::Penalty().
    void enable_scalar(const PwdB* pwd, const STYPE* sig53tab, int max_segment)
    {
        std::vector<short> pen((size_t) max_segment + 1);
        for (int n = 0; n <= max_segment; ++n) pen[n] = pwd->IntPen->Penalty(n);
        std::vector<short> tab(sig53tab, sig53tab + 544);
        int rc = gspaln_set_ng_tables(ctx_, tab.data(), pen.data(), (int) pen.size(), pwd->codonk1);
        if (rc != GSPALN_OK) die("gspaln_set_ng_tables", rc, ctx_);
    }

    // == Aln2s1::lspS_ng(wdw) with the corners appended to mfd (src/fwd2s1.cc:1801-1897).
    // int53: the INT53 array of seqs[1]->exin indexed by column (may be 0: blocks with fewer than
    // 8 rows are then reported as unsupported).  Returns false if the problem needs a kernel that
    // is not on the device (the caller falls back to the stock lspS_ng); *scr receives the score.
    bool lspS_ng(const Seq** seqs, const WINDOW& wdw, Mfile* mfd, const INT53* int53, VTYPE* scr)
    {
        gspaln_task t;
        gspaln_result r;
        fill(t, seqs, wdw, GSPALN_FORWARD_WIP);
        const Seq* b = seqs[1];
        if (int53) {
            // INT53 is four 4-bit fields in one INT (src/codepot.h:49-54): the low 16 bits are the
            // dinc5 | dinc3 << 4 | cano5 << 8 | cano3 << 12 word of gspaln_task.int53
            int53_.assign((size_t) b->right + 2, 0);
            for (int i = b->left; i <= b->right; ++i)
                int53_[i] = (unsigned short) (int53[i].dinc5 | (int53[i].dinc3 << 4) |
                                              (int53[i].cano5 << 8) | (int53[i].cano3 << 12));
            t.int53 = int53_.data();
        }
        gspaln_lsp_opts o = {MaxVmfSpace, (int) alprm.sh, (int) alprm.ubh, (int) algmode.alg};
        int cap = (t.a_right - t.a_left) + (t.b_right - t.b_left) + 8;
        for (;;) {
            skl_.assign(2 * (size_t) cap, 0);
            t.skl_cap = cap;
            r.skl = skl_.data();
            r.cpos = 0;
            int rc = gspaln_lsp(ctx_, &t, 1, &o, &r);
            if (rc != GSPALN_OK) die("gspaln_lsp", rc, ctx_);
            if (r.status != GSPALN_ST_SKL_OVERFLOW) break;
            cap = r.n_skl + 8;
        }
        if (r.status == GSPALN_ST_UNSUPPORTED) return false;
        if (r.status == GSPALN_ST_BAD_TRACE) fatal("Unexpected dir\n");
        for (int i = 0; i < r.n_skl; ++i) {
            SKL wsk = {skl_[2 * i], skl_[2 * i + 1]};
            mfd->write((UPTR) &wsk);
        }
        *scr = (VTYPE) r.score;
        return true;
    }

    // == SimdAln2s1(seqs, pwd, wdw, spjcs, cip, 1).scoreonlyS1_wip()
    VTYPE scoreonlyS1_wip(const Seq** seqs, const WINDOW& wdw)
    {
        gspaln_task t;
        gspaln_result r;
        fill(t, seqs, wdw, GSPALN_SCOREONLY_WIP);
        r.skl = 0;
        int rc = gspaln_submit(ctx_, &t, 1, &r);
        if (rc != GSPALN_OK) die("gspaln_submit", rc, ctx_);
        return (VTYPE) r.score;
    }
};

}   // namespace gspaln
#endif
