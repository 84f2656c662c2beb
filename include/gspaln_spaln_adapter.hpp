// gspaln_spaln_adapter.hpp -- header-only adapter that lets the reference (ogotoh/spaln)
// call libgspaln at its SimdAln2s1 / SimdAln2h1 seams.  It is compiled INSIDE a Spaln translation
// unit: it expects the reference's own headers (aln.h -> seq.h, codepot.h, mfile.h) to be
// included first and uses their types (Seq, PwdB, WINDOW, SKL, Mfile, IntronPrm, algmode).
//
// Replaces (all paths relative to the reference tree):
//   SimdAln2s1 ctor + forwardS1_wip(Mfile*)   src/fwd2s1_simd.h:191-333, src/fwd2s1_wip_simd.h:233-474
//   SimdAln2s1 ctor + scoreonlyS1_wip()       src/fwd2s1_wip_simd.h:42-231
//   Aln2s1::lspS_ng(wdw) as a whole           src/fwd2s1.cc:1801-1897 (driver: trace-back vs
//                                             Hirschberg dispatch, post-work, scalar small blocks)
//   SimdAln2h1 ctor + forwardH1_wip(Mfile*)   src/fwd2h1_simd.h:196-382, src/fwd2h1_wip_simd.h:50-336
//   Aln2h1::lspH_ng(wdw) as a whole           src/fwd2h1.cc:2134-2230
// Every call goes through the coalescing queue of the C-ABI (gspaln_queue_* / gspaln_h_queue_*):
// Spaln's pthread workers (src/spaln.cc:1363-1468) each block with one problem and the queue's
// dispatcher runs what is pending as one device batch, so the engines are safe to share between
// threads.  See INTEGRATION.md for the patch of src/fwd2s1.cc / src/fwd2h1.cc
// (include/gspaln_spaln_dropin.hpp holds the hooks).
#ifndef GSPALN_SPALN_ADAPTER_HPP
#define GSPALN_SPALN_ADAPTER_HPP

#include "gspaln.h"

#include <cstdlib>
#include <cstring>
#include <vector>

namespace gspaln {

// the corner list of one result -> the caller's Mfile, in the order received (alignment end
// first: the order Anti_rhomb_coord::traceback uses, src/rhomb_coord.h:222-235)
inline void write_corners(Mfile* mfd, const std::vector<int>& skl, int n)
{
    for (int i = 0; i < n; ++i) {
        SKL wsk = {skl[2 * i], skl[2 * i + 1]};
        mfd->write((UPTR) &wsk);
    }
}

// INT53 is four 4-bit fields in one INT (src/codepot.h:49-54): the low 16 bits are the
// dinc5 | dinc3 << 4 | cano5 << 8 | cano3 << 12 word of gspaln_task.int53
inline void pack_int53(std::vector<unsigned short>& out, const INT53* int53, const Seq* b)
{
    out.assign((size_t) b->right + 2, 0);
    for (int i = b->left; i <= b->right; ++i)
        out[i] = (unsigned short) (int53[i].dinc5 | (int53[i].dinc3 << 4) |
                                   (int53[i].cano5 << 8) | (int53[i].cano3 << 12));
}

inline gspaln_lsp_opts lsp_opts_now()
{
    gspaln_lsp_opts o = {MaxVmfSpace, (int) alprm.sh, (int) alprm.ubh, (int) algmode.alg};
    return o;
}

// ===================================================================================== DNA
class SpalnEngine {
    gspaln_ctx* ctx_ = nullptr;
    gspaln_queue* q_ = nullptr;
    std::vector<short> sig53tab_;       // the table the scalar kernel was bound with

    static void die(const char* what, int rc, const gspaln_ctx* c)
    {
        // the reference's own error convention: fatal() == message + exit(1) (src/adddef.h:173-182)
        fatal("gspaln: %s failed (%d): %s\n", what, rc, c ? gspaln_last_error(c) : "");
    }

    struct Scratch {                    // per call: the engine is shared between threads
        std::vector<short> sig5, sig3;
        std::vector<unsigned short> int53;
        std::vector<int> skl;
        std::vector<int32_t> cip;
    };

    // gspaln_task.cip: Cip_score::cip_score(m) of the rows of the current range (src/gsinfo.h:127-139)
    static void fill_cip(gspaln_task& t, Scratch& s, const Seq* a, const Cip_score* cip)
    {
        if (!cip) return;
        s.cip.assign((size_t) a->right + 1, 0);
        for (int m = a->left; m <= a->right; ++m) s.cip[m] = (int32_t) cip->cip_score(m);
        t.cip = s.cip.data();
    }

    static void fill(gspaln_task& t, Scratch& s, const Seq** seqs, const WINDOW& wdw, int kind)
    {
        const Seq* a = seqs[0];
        const Seq* b = seqs[1];
        memset(&t, 0, sizeof(t));
        t.kind = kind;
        t.int53 = 0;    // only the scalar kernels read the INT53 array
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
        s.sig5.assign(n, 0); s.sig3.assign(n, 0);
        if (b->inex.intr && b->exin)
            for (int i = b->left; i <= b->right; ++i) {
                const SGPT2* g = b->exin->score_n(i);
                s.sig5[i] = g->sig5; s.sig3[i] = g->sig3;
            }
        t.sig5 = s.sig5.data(); t.sig3 = s.sig3.data();
        t.skl_cap = 0;
    }

public:
    // the frozen copy of the globals the reference kernels read (gspaln_params)
    static gspaln_params freeze(const PwdB* pwd, bool spliced)
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
        return p;
    }

    explicit SpalnEngine(const PwdB* pwd, int device = 0, bool spliced = true)
    {
        gspaln_params p = freeze(pwd, spliced);
        int rc = gspaln_create(&ctx_, &p, device);
        if (rc != GSPALN_OK) die("gspaln_create", rc, ctx_);
        rc = gspaln_queue_create(&q_, ctx_, 256, 50);
        if (rc != GSPALN_OK) die("gspaln_queue_create", rc, ctx_);
    }
    ~SpalnEngine() { gspaln_queue_destroy(q_); gspaln_destroy(ctx_); }
    SpalnEngine(const SpalnEngine&) = delete;
    SpalnEngine& operator=(const SpalnEngine&) = delete;

    // == SimdAln2s1(seqs, pwd, wdw, spjcs, cip, 1).forwardS1_wip(mfd)
    VTYPE forwardS1_wip(const Seq** seqs, const WINDOW& wdw, Mfile* mfd)
    {
        gspaln_task t;
        gspaln_result r;
        Scratch s;
        fill(t, s, seqs, wdw, GSPALN_FORWARD_WIP);
        int cap = 256;
        for (;;) {
            s.skl.assign(2 * (size_t) cap, 0);
            t.skl_cap = cap;
            memset(&r, 0, sizeof(r));
            r.skl = s.skl.data();
            int rc = gspaln_queue_submit(q_, &t, &r);
            if (rc != GSPALN_OK) die("gspaln_submit", rc, ctx_);
            if (r.status != GSPALN_ST_SKL_OVERFLOW) break;
            cap = r.n_skl + 8;
        }
        if (r.status == GSPALN_ST_BAD_TRACE) fatal("Unexpected dir\n");     // src/rhomb_coord.h:216
        write_corners(mfd, s.skl, r.n_skl);
        return (VTYPE) r.score;
    }

    // Tables of the scalar exact-ILD kernel (the reference runs Aln2s1::forwardS_ng for blocks with
    // fewer than 8 query rows, src/fwd2s1.cc:1676): sig53tab = Exinon::sig53tab[0] (544 shorts),
    // the length penalty is evaluated here with the reference's own IntronPenalty::Penalty().
    void enable_scalar(const PwdB* pwd, const STYPE* sig53tab, int max_segment)
    {
        std::vector<short> pen((size_t) max_segment + 1);
        for (int n = 0; n <= max_segment; ++n) pen[n] = pwd->IntPen->Penalty(n);
        sig53tab_.assign(sig53tab, sig53tab + 544);
        int rc = gspaln_set_ng_tables(ctx_, sig53tab_.data(), pen.data(), (int) pen.size(), pwd->codonk1);
        if (rc != GSPALN_OK) die("gspaln_set_ng_tables", rc, ctx_);
    }
    // a segment whose Exinon carries another dinucleotide table (Seq::many != 1) must not use
    // the bound one
    bool same_sig53tab(const STYPE* tab) const
    {
        return !sig53tab_.empty() && !memcmp(sig53tab_.data(), tab, 544 * sizeof(short));
    }

    // == Aln2s1::lspS_ng(wdw) with the corners appended to mfd (src/fwd2s1.cc:1801-1897).
    // int53: the INT53 array of seqs[1]->exin indexed by column (may be 0: blocks with fewer than
    // 8 rows are then reported as unsupported).  Returns false if the problem needs a kernel that
    // is not on the device (the caller falls back to the stock lspS_ng); *scr receives the score.
    bool lspS_ng(const Seq** seqs, const WINDOW& wdw, Mfile* mfd, const INT53* int53, VTYPE* scr,
                 const Cip_score* cip = 0)
    {
        gspaln_task t;
        gspaln_result r;
        Scratch s;
        fill(t, s, seqs, wdw, GSPALN_FORWARD_WIP);
        if (int53) {
            pack_int53(s.int53, int53, seqs[1]);
            t.int53 = s.int53.data();
        }
        fill_cip(t, s, seqs[0], cip);   // read by the exact-ILD kernel of blocks with < 8 rows
        const gspaln_lsp_opts o = lsp_opts_now();
        int cap = 256;
        for (;;) {
            s.skl.assign(2 * (size_t) cap, 0);
            t.skl_cap = cap;
            memset(&r, 0, sizeof(r));
            r.skl = s.skl.data();
            int rc = gspaln_queue_submit_lsp(q_, &t, &o, &r);
            if (rc != GSPALN_OK) die("gspaln_lsp", rc, ctx_);
            if (r.status != GSPALN_ST_SKL_OVERFLOW) break;
            cap = r.n_skl + 8;
        }
        if (r.status == GSPALN_ST_UNSUPPORTED) return false;
        if (r.status == GSPALN_ST_BAD_TRACE) fatal("Unexpected dir\n");
        if (r.status != GSPALN_ST_OK) die("gspaln_lsp (problem status)", r.status, ctx_);
        write_corners(mfd, s.skl, r.n_skl);
        *scr = (VTYPE) r.score;
        return true;
    }

    // == SimdAln2s1(seqs, pwd, wdw, spjcs, cip, 1).scoreonlyS1_wip()
    VTYPE scoreonlyS1_wip(const Seq** seqs, const WINDOW& wdw)
    {
        gspaln_task t;
        gspaln_result r;
        Scratch s;
        fill(t, s, seqs, wdw, GSPALN_SCOREONLY_WIP);
        memset(&r, 0, sizeof(r));
        int rc = gspaln_queue_submit(q_, &t, &r);
        if (rc != GSPALN_OK) die("gspaln_submit", rc, ctx_);
        return (VTYPE) r.score;
    }

    // == Aln2s1::trcbkalignS_ng(wdw) on its scalar branch (src/fwd2s1.cc:1676-1706): forwardS_ng +
    // Vmf::traceback + the start-point adjustment -- every trace-back of `-A0`, and the blocks with
    // fewer than 8 query rows of the other modes.  Needs enable_scalar() and the segment's INT53
    // array.  false: the kernel ran out of path records (the caller runs the stock code).
    bool forwardS_ng(const Seq** seqs, const WINDOW& wdw, Mfile* mfd, const INT53* int53, VTYPE* scr,
                     const Cip_score* cip = 0)
    {
        gspaln_task t;
        gspaln_result r;
        Scratch s;
        fill(t, s, seqs, wdw, GSPALN_FORWARD_NG);
        if (int53) {
            pack_int53(s.int53, int53, seqs[1]);
            t.int53 = s.int53.data();
        }
        fill_cip(t, s, seqs[0], cip);
        int cap = 256;
        for (;;) {
            s.skl.assign(2 * (size_t) cap, 0);
            t.skl_cap = cap;
            memset(&r, 0, sizeof(r));
            r.skl = s.skl.data();
            int rc = gspaln_queue_submit(q_, &t, &r);
            if (rc != GSPALN_OK) die("gspaln_submit", rc, ctx_);
            if (r.status != GSPALN_ST_SKL_OVERFLOW) break;
            cap = r.n_skl + 8;
        }
        if (r.status != GSPALN_ST_OK) return false;
        write_corners(mfd, s.skl, r.n_skl);
        *scr = (VTYPE) r.score;
        return true;
    }

    // == Aln2s1::scorealoneS_ng(wdw): what HomScoreS_ng runs for queries shorter than 4 residues
    // (src/fwd2s1.cc:2704-2705).  Needs enable_scalar() and the segment's INT53 array.
    VTYPE scorealoneS_ng(const Seq** seqs, const WINDOW& wdw, const INT53* int53, const Cip_score* cip = 0)
    {
        gspaln_task t;
        gspaln_result r;
        Scratch s;
        fill(t, s, seqs, wdw, GSPALN_SCOREALONE_NG);
        pack_int53(s.int53, int53, seqs[1]);
        t.int53 = s.int53.data();
        fill_cip(t, s, seqs[0], cip);
        memset(&r, 0, sizeof(r));
        int rc = gspaln_queue_submit(q_, &t, &r);
        if (rc != GSPALN_OK) die("gspaln_submit", rc, ctx_);
        return (VTYPE) r.score;
    }

    void queue_stats(int64_t* tasks, int64_t* batches) const { gspaln_queue_stats(q_, tasks, batches); }
};

// ===================================================================================== protein
class SpalnEngineH {
    gspaln_h_ctx* ctx_ = nullptr;
    gspaln_h_queue* q_ = nullptr;
    std::vector<short> sig53tab_;

    static void die(const char* what, int rc, const gspaln_h_ctx* c)
    {
        fatal("gspaln: %s failed (%d): %s\n", what, rc, c ? gspaln_h_last_error(c) : "");
    }

    struct Scratch {
        std::vector<unsigned short> int53;
        std::vector<int> skl;
        std::vector<int32_t> cip;
    };

    // gspaln_h_task.cip: Cip_score::cip_score(c) by coding position c = 3 m - phase (src/fwd2h1.cc:352-354)
    static void fill_cip(gspaln_h_task& t, Scratch& s, const Seq* a, const Cip_score* cip)
    {
        if (!cip) return;
        s.cip.assign((size_t) 3 * a->right + 2, 0);
        for (int c = std::max(0, 3 * a->left - 1); c <= 3 * a->right + 1; ++c)
            s.cip[c] = (int32_t) cip->cip_score(c);
        t.cip = s.cip.data();
    }

    // one task == what the SimdAln2h1 constructor dereferences (src/fwd2h1_simd.h:196-382): the
    // SGPT6 array is taken as it is (gspaln_sgpt6 has the layout of src/codepot.h:34-43)
    static void fill(gspaln_h_task& t, const Seq** seqs, const WINDOW& wdw, int kind)
    {
        const Seq* a = seqs[0];
        const Seq* b = seqs[1];
        memset(&t, 0, sizeof(t));
        t.kind = kind;
        t.a = a->at(0);
        t.b = b->at(0);
        t.sg = reinterpret_cast<const gspaln_sgpt6*>(b->exin->score_p(0));
        t.b_len = b->len; t.a_len = a->len;
        t.a_left = a->left; t.a_right = a->right;
        t.b_left = b->left; t.b_right = b->right;
        t.a_exgl = a->inex.exgl; t.a_exgr = a->inex.exgr;
        t.b_exgl = b->inex.exgl; t.b_exgr = b->inex.exgr;
        t.lw = wdw.lw; t.up = wdw.up;
    }

public:
    static gspaln_h_params freeze(const PwdB* pwd, bool spliced)
    {
        gspaln_h_params p = gspaln_h_params();
        p.gop = pwd->BasicGOP; p.gep = pwd->BasicGEP;
        p.lgop = pwd->LongGOP; p.lgep = pwd->LongGEP;
        p.codonk1 = pwd->codonk1;
        p.gw1 = pwd->GapW1; p.gw2 = pwd->GapW2; p.gw3 = pwd->GapW3;
        p.gape1 = pwd->GapE1; p.gape2 = pwd->GapE2;
        p.ipen = (spliced && pwd->IntPen) ? pwd->IntPen->Penalty() : 0;
        p.llmt = IntronPrm.llmt;
        p.nquant = IntronPrm.nquant;
        for (int j = 0; j < p.nquant && j < GSPALN_MAXQUANT && pwd->IntPen && pwd->IntPen->qm; ++j) {
            p.quant_len[j] = pwd->IntPen->qm[j].len;
            p.quant_pen[j] = pwd->IntPen->qm[j].pen;
        }
        p.avmch = (int) pwd->simmtx->AvTrc();
        p.lcl = (int) algmode.lcl;
        p.spj = spliced ? 1 : 0;
        p.simdim = pwd->simmtx->dim;
        for (int q = 0; q < p.simdim; ++q)
            for (int g = 0; g < p.simdim; ++g)
                p.simmtx[q * p.simdim + g] = pwd->simmtx->mtx[q][g];
        return p;
    }

    explicit SpalnEngineH(const PwdB* pwd, int device = 0, bool spliced = true)
    {
        gspaln_h_params p = freeze(pwd, spliced);
        int rc = gspaln_h_create(&ctx_, &p, device);
        if (rc != GSPALN_OK) die("gspaln_h_create", rc, ctx_);
        rc = gspaln_h_queue_create(&q_, ctx_, 256, 50);
        if (rc != GSPALN_OK) die("gspaln_h_queue_create", rc, ctx_);
    }
    ~SpalnEngineH() { gspaln_h_queue_destroy(q_); gspaln_h_destroy(ctx_); }
    SpalnEngineH(const SpalnEngineH&) = delete;
    SpalnEngineH& operator=(const SpalnEngineH&) = delete;

    // tables of the scalar kernel forwardH_ng (blocks with fewer than 8 query rows,
    // src/fwd2h1.cc:2007): spj_tabs = the split-codon tables of SpJunc::spjseq + aa2nuc in the
    // layout of gspaln_h_set_ng_tables (the caller's TU sees the static tables of codepot.h)
    void enable_scalar(const PwdB* pwd, const STYPE* sig53tab, const unsigned char* spj_tabs, int max_segment)
    {
        std::vector<short> pen((size_t) max_segment + 1);
        for (int n = 0; n <= max_segment; ++n) pen[n] = pwd->IntPen->Penalty(n);
        sig53tab_.assign(sig53tab, sig53tab + 544);
        int rc = gspaln_h_set_ng_tables(ctx_, sig53tab_.data(), pen.data(), (int) pen.size(), spj_tabs,
                                        IntronPrm.minl, pwd->ExtraGOP, pwd->GapW3L, pwd->Noll);
        if (rc != GSPALN_OK) die("gspaln_h_set_ng_tables", rc, ctx_);
    }
    bool same_sig53tab(const STYPE* tab) const
    {
        return !sig53tab_.empty() && !memcmp(sig53tab_.data(), tab, 544 * sizeof(short));
    }

    // == SimdAln2h1(seqs, pwd, wdw, spjcs, cip, 1).forwardH1_wip(mfd); mfd == 0: score only
    // (HomScoreH_ng, src/fwd2h1.cc:3293-3307)
    VTYPE forwardH1_wip(const Seq** seqs, const WINDOW& wdw, Mfile* mfd)
    {
        gspaln_h_task t;
        gspaln_result r;
        Scratch s;
        fill(t, seqs, wdw, mfd ? GSPALN_FORWARD_WIP : GSPALN_SCOREONLY_WIP);
        int cap = mfd ? 256 : 0;
        for (;;) {
            s.skl.assign(2 * (size_t) cap + 2, 0);
            t.skl_cap = cap;
            memset(&r, 0, sizeof(r));
            r.skl = mfd ? s.skl.data() : 0;
            int rc = gspaln_h_queue_submit(q_, &t, &r);
            if (rc != GSPALN_OK) die("gspaln_h_submit", rc, ctx_);
            if (r.status != GSPALN_ST_SKL_OVERFLOW) break;
            cap = r.n_skl + 8;
        }
        if (r.status == GSPALN_ST_BAD_TRACE) fatal("Unexpected dir\n");
        if (mfd) write_corners(mfd, s.skl, r.n_skl);
        return (VTYPE) r.score;
    }

    // == Aln2h1::trcbkalignH_ng(wdw) on its scalar branch (src/fwd2h1.cc:2007-2037): forwardH_ng +
    // Vmf::traceback + the start-point adjustment; mfd == 0: the score alone (HomScoreH_ng under
    // `-A0` and for queries shorter than 8 residues, src/fwd2h1.cc:3297-3298 -- the score does not
    // depend on the path records).  false: out of path records (the caller runs the stock code).
    bool forwardH_ng(const Seq** seqs, const WINDOW& wdw, Mfile* mfd, const INT53* int53, VTYPE* scr,
                     const Cip_score* cip = 0)
    {
        gspaln_h_task t;
        gspaln_result r;
        Scratch s;
        fill(t, seqs, wdw, GSPALN_FORWARD_NG);
        pack_int53(s.int53, int53, seqs[1]);
        t.int53 = s.int53.data();
        fill_cip(t, s, seqs[0], cip);
        int cap = 256;
        for (;;) {
            s.skl.assign(2 * (size_t) cap, 0);
            t.skl_cap = cap;
            memset(&r, 0, sizeof(r));
            r.skl = s.skl.data();
            int rc = gspaln_h_queue_submit(q_, &t, &r);
            if (rc != GSPALN_OK) die("gspaln_h_submit", rc, ctx_);
            if (r.status != GSPALN_ST_SKL_OVERFLOW || !mfd) break;
            cap = r.n_skl + 8;
        }
        if (r.status != GSPALN_ST_OK && !(r.status == GSPALN_ST_SKL_OVERFLOW && !mfd)) return false;
        if (mfd) write_corners(mfd, s.skl, r.n_skl);
        *scr = (VTYPE) r.score;
        return true;
    }

    // == Aln2h1::lspH_ng(wdw) with the corners appended to mfd (src/fwd2h1.cc:2134-2230); false:
    // the problem needs a kernel that is not on the device (the caller runs the stock lspH_ng)
    bool lspH_ng(const Seq** seqs, const WINDOW& wdw, Mfile* mfd, const INT53* int53, VTYPE* scr,
                 const Cip_score* cip = 0)
    {
        gspaln_h_task t;
        gspaln_result r;
        Scratch s;
        fill(t, seqs, wdw, GSPALN_FORWARD_WIP);
        if (int53) {
            pack_int53(s.int53, int53, seqs[1]);
            t.int53 = s.int53.data();
        }
        fill_cip(t, s, seqs[0], cip);
        const gspaln_lsp_opts o = lsp_opts_now();
        int cap = 256;
        for (;;) {
            s.skl.assign(2 * (size_t) cap, 0);
            t.skl_cap = cap;
            memset(&r, 0, sizeof(r));
            r.skl = s.skl.data();
            int rc = gspaln_h_queue_submit_lsp(q_, &t, &o, &r);
            if (rc != GSPALN_OK) die("gspaln_h_lsp", rc, ctx_);
            if (r.status != GSPALN_ST_SKL_OVERFLOW) break;
            cap = r.n_skl + 8;
        }
        if (r.status == GSPALN_ST_UNSUPPORTED) return false;
        if (r.status == GSPALN_ST_BAD_TRACE) fatal("Unexpected dir\n");
        if (r.status != GSPALN_ST_OK) die("gspaln_h_lsp (problem status)", r.status, ctx_);
        write_corners(mfd, s.skl, r.n_skl);
        *scr = (VTYPE) r.score;
        return true;
    }

    void queue_stats(int64_t* tasks, int64_t* batches) const { gspaln_h_queue_stats(q_, tasks, batches); }
};

}   // namespace gspaln
#endif
