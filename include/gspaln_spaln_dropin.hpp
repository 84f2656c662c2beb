// gspaln_spaln_dropin.hpp -- the hooks that route Spaln's DP seams through libgspaln.
//
// Included at the top of src/fwd2s1.cc and src/fwd2h1.cc (after their own #includes); the
// one-line hooks below are then placed at the head of five reference functions -- see
// INTEGRATION.md for the patch and oracle/Makefile (`make dropin`) for the sed script that applies
// it to a scratch copy when the test binaries are built:
//
//   VTYPE Aln2s1::lspS_ng(const WINDOW& wdw)            { GSPALN_HOOK_LSPS   ...   src/fwd2s1.cc:1801
//   VTYPE Aln2s1::trcbkalignS_ng(wdw, spj, mc)          { GSPALN_HOOK_TRCBKS ...   src/fwd2s1.cc:1667
//   VTYPE HomScoreS_ng(const Seq* seqs[], const PwdB*)  { GSPALN_HOOK_HOMS   ...   src/fwd2s1.cc:2696
//   VTYPE Aln2h1::lspH_ng(const WINDOW& wdw)            { GSPALN_HOOK_LSPH   ...   src/fwd2h1.cc:2134
//   VTYPE Aln2h1::trcbkalignH_ng(wdw, spj, mc)          { GSPALN_HOOK_TRCBKH ...   src/fwd2h1.cc:1997
//   VTYPE HomScoreH_ng(const Seq* seqs[], const PwdB*)  { GSPALN_HOOK_HOMH   ...   src/fwd2h1.cc:3288
//
// A hook returns the engine's result when the call is one the device covers (-A2 / -A3, no
// conserved-intron annotation on the query) and falls through to the stock code otherwise, so a
// patched binary behaves like the stock one for every other option set.  Engines are created on
// first use, one per (PwdB, spliced) pair, and shared by all pthread workers through the
// coalescing queues of the C-ABI.
//
// GSPALN_HARVEST (test builds only): instead of replacing the calls, the lsp hooks run the STOCK
// code and append every top-level lsp*_ng call -- inputs, frozen parameters, score, corners -- to
// the file named by $GSPALN_HARVEST_FILE; tests/test_gpu_realdata.py replays that file through
// gspaln_lsp / gspaln_h_lsp.
#ifndef GSPALN_SPALN_DROPIN_HPP
#define GSPALN_SPALN_DROPIN_HPP

#include "gspaln_spaln_adapter.hpp"

#include <cstdio>
#include <atomic>
#include <chrono>
#include <map>
#include <mutex>

namespace gspaln {
namespace dropin {

// Exinon keeps its INT53 array and sig53tab private (src/codepot.h:72-80).  A maintainer adds two
// one-line accessors next to Exinon::isDonor; this header must work on the unmodified class, and
// an explicit instantiation may name private members.
template <typename Tag, typename Tag::type M> struct Rob {
    friend typename Tag::type get(Tag) { return M; }
};
struct ExinonInt53 { typedef INT53* Exinon::*type; friend type get(ExinonInt53); };
struct ExinonTab { typedef STYPE** Exinon::*type; friend type get(ExinonTab); };
template struct Rob<ExinonInt53, &Exinon::int53>;
template struct Rob<ExinonTab, &Exinon::sig53tab>;

inline const INT53* int53_of(const Seq* b) { return b->exin ? b->exin->*get(ExinonInt53()) : 0; }
inline const STYPE* sig53tab_of(const Seq* b)
{
    if (!b->exin) return 0;
    STYPE** t = b->exin->*get(ExinonTab());
    return t ? t[0] : 0;
}

constexpr int MAX_SEGMENT = 1 << 20;    // Penalty(n) is tabulated up to this intron length

inline int device_index()
{
    const char* e = getenv("GSPALN_DEVICE");
    return e ? atoi(e) : 0;
}

inline std::mutex& registry_mutex() { static std::mutex m; return m; }

// Calls answered by the device, by hook.  GSPALN_DROPIN_STATS=1 prints them to stderr at exit (what
// the whole-program tests read to prove that a run went through the CUDA kernels).
struct HookStats {
    enum { LSP, TRCBK_WIP, TRCBK_EXACT, HOM_WIP, HOM_EXACT, EXACT_NO_TABLES, EXACT_OVERFLOW, N };
    std::atomic<long long> n[2][N];
    std::atomic<long long> setup_us[2];     // engine creation (CUDA context included), exact-ILD tables
    HookStats()
    {
        for (auto& row : n) for (auto& c : row) c = 0;
        setup_us[0] = setup_us[1] = 0;
        if (getenv("GSPALN_DROPIN_STATS")) atexit(report);
    }
    static HookStats& get() { static HookStats s; return s; }
    static void report()
    {
        static const char* names[N] = {"lsp", "trcbk_wip", "trcbk_exact", "homscore_wip", "homscore_exact",
                                       "exact_no_tables", "exact_overflow"};
        HookStats& s = get();
        for (int p = 0; p < 2; ++p) {
            fprintf(stderr, "gspaln drop-in (%s):", p ? "protein" : "dna");
            for (int k = 0; k < N; ++k) fprintf(stderr, " %s=%lld", names[k], (long long) s.n[p][k]);
            fputc('\n', stderr);
        }
        fprintf(stderr, "gspaln drop-in set-up: engines %.2f s, exact-ILD tables %.2f s\n",
                1e-6 * (double) s.setup_us[0], 1e-6 * (double) s.setup_us[1]);
    }
};
inline bool counted(bool ok, int protein, int what)
{
    if (ok) ++HookStats::get().n[protein][what];
    return ok;
}

// split-codon tables of SpJunc::spjseq (src/codepot.h:130-190) and aa2nuc (src/seq.cc:76) in the
// layout of gspaln_h_set_ng_tables
inline void spj_tables(unsigned char* out)
{
    int k = 0;
    for (int i = 0; i < 257; ++i) { out[k++] = spj_tron_tab[i][0]; out[k++] = spj_tron_tab[i][1]; }
    for (int i = 0; i < 64; ++i) { out[k++] = spj_amb_tron_tab[i][0]; out[k++] = spj_amb_tron_tab[i][1]; }
    for (int i = 0; i < 64; ++i) { out[k++] = spj_tron_amb_tab[i][0]; out[k++] = spj_tron_amb_tab[i][1]; }
    for (int i = 0; i < 26; ++i) out[k++] = aa2nuc[i];
}

inline SpalnEngine* engineS(const PwdB* pwd, const Seq* b)
{
    static std::map<std::pair<const PwdB*, bool>, SpalnEngine*> reg;
    std::lock_guard<std::mutex> lk(registry_mutex());
    const bool spliced = b->inex.intr;
    SpalnEngine*& e = reg[std::make_pair(pwd, spliced)];
    if (!e) {
        const auto t0 = std::chrono::steady_clock::now();
        e = new SpalnEngine(pwd, device_index(), spliced);
        const auto t1 = std::chrono::steady_clock::now();
        if (spliced && sig53tab_of(b) && pwd->IntPen) e->enable_scalar(pwd, sig53tab_of(b), MAX_SEGMENT);
        const auto t2 = std::chrono::steady_clock::now();
        HookStats::get().setup_us[0] += std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count();
        HookStats::get().setup_us[1] += std::chrono::duration_cast<std::chrono::microseconds>(t2 - t1).count();
    }
    return e;
}

inline SpalnEngineH* engineH(const PwdB* pwd, const Seq* b)
{
    static std::map<std::pair<const PwdB*, bool>, SpalnEngineH*> reg;
    std::lock_guard<std::mutex> lk(registry_mutex());
    const bool spliced = b->inex.intr;
    SpalnEngineH*& e = reg[std::make_pair(pwd, spliced)];
    if (!e) {
        const auto t0 = std::chrono::steady_clock::now();
        e = new SpalnEngineH(pwd, device_index(), spliced);
        const auto t1 = std::chrono::steady_clock::now();
        if (spliced && sig53tab_of(b) && pwd->IntPen) {
            unsigned char tabs[796];
            spj_tables(tabs);
            e->enable_scalar(pwd, sig53tab_of(b), tabs, MAX_SEGMENT);
        }
        const auto t2 = std::chrono::steady_clock::now();
        HookStats::get().setup_us[0] += std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count();
        HookStats::get().setup_us[1] += std::chrono::duration_cast<std::chrono::microseconds>(t2 - t1).count();
    }
    return e;
}

// What the device covers.  Drivers: lspS_ng / lspH_ng under -A2 / -A3 (`_wip` kernels) and under
// -A0, the reference's default (exact-ILD kernels + the scalar Hirschberg passes hirschbergS_ng /
// hirschbergH_ng).  Kernels behind trcbkalign*_ng / HomScore*_ng: the `_wip` kernels for
// simd >= 2, the exact-ILD kernels for every call the reference sends to its scalar code -- all
// of `-A0` and blocks with fewer than 8 query rows in any mode.  The int16 exact-ILD kernels of
// `-A1` stay with the stock code.  Cip_score (queries annotated with intron positions) is read by
// the exact-ILD kernels only and travels with the task.
inline bool exact_tables_ok(const Seq* b, bool same_tab)
{
    return !b->inex.intr || (int53_of(b) && same_tab && b->right - b->left < MAX_SEGMENT);
}

// ---------------------------------------------------------------------------------- DNA hooks
inline bool lspS(const Seq** seqs, const PwdB* pwd, const WINDOW& wdw, Mfile* mfd, int simd,
                 const Cip_score* cip, VTYPE* scr)
{
    if (simd == 1) return false;
    const Seq* b = seqs[1];
    SpalnEngine* e = engineS(pwd, b);
    const INT53* i53 = (b->inex.intr && e->same_sig53tab(sig53tab_of(b))) ? int53_of(b) : 0;
    if (simd == 0 && (!i53 || b->right - b->left >= MAX_SEGMENT)) return false;    // -A0 runs on the exact-ILD tables
    return counted(e->lspS_ng(seqs, wdw, mfd, i53, scr, cip), 0, HookStats::LSP);
}

inline bool trcbkS(const Seq** seqs, const PwdB* pwd, const WINDOW& wdw, Mfile* mfd, int simd,
                   const Cip_score* cip, const RANGE* mc, VTYPE* scr)
{
    // trcbkalignS_ng (src/fwd2s1.cc:1667-1710): its SIMD branch for -A2 / -A3, its scalar branch
    // (forwardS_ng + Vmf) for -A0 and for m < 8; the cut-range variant and -A1's forwardS1 stay
    // with the stock code
    if (mc || wdw.width < 0) return false;
    const Seq* b = seqs[1];
    const int m = seqs[0]->right - seqs[0]->left;
    if (simd >= 2 && m >= 8) {
        *scr = engineS(pwd, b)->forwardS1_wip(seqs, wdw, mfd);
        return counted(true, 0, HookStats::TRCBK_WIP);
    }
    if (simd != 0 && m >= 8) return false;
    if (m < 1 || b->right <= b->left || !b->inex.intr) return false;
    SpalnEngine* e = engineS(pwd, b);
    if (!exact_tables_ok(b, e->same_sig53tab(sig53tab_of(b)))) return !counted(true, 0, HookStats::EXACT_NO_TABLES);
    if (counted(e->forwardS_ng(seqs, wdw, mfd, int53_of(b), scr, cip), 0, HookStats::TRCBK_EXACT)) return true;
    return !counted(true, 0, HookStats::EXACT_OVERFLOW);
}

inline bool homscoreS(const Seq** seqs, const PwdB* pwd, VTYPE* scr)
{
    const int simd = algmode.alg & 3;
    const Seq* a = seqs[0];
    const Seq* b = seqs[1];
    if (simd == 1 && a->right - a->left >= 4) return false;     // -A1: scoreonlyS1, stock code
    if (simd == 3) IntronPrm.nquant = 1;        // what the Aln2s1 constructor does (src/fwd2s1.cc:125)
    WINDOW wdw;
    stripe(seqs, &wdw, alprm.sh);
    SpalnEngine* e = engineS(pwd, b);
    if (simd == 0 || a->right - a->left < 4) {  // src/fwd2s1.cc:2704-2705
        if (b->inex.intr && !(int53_of(b) && e->same_sig53tab(sig53tab_of(b)))) return false;
        if (!b->inex.intr || b->right - b->left >= MAX_SEGMENT) return false;
        if (a->sigII) {                         // the Aln2s1 constructor's Cip_score (src/fwd2s1.cc:124)
            const Cip_score cs(a);
            *scr = e->scorealoneS_ng(seqs, wdw, int53_of(b), &cs);
        } else
            *scr = e->scorealoneS_ng(seqs, wdw, int53_of(b));
        return counted(true, 0, HookStats::HOM_EXACT);
    }
    *scr = e->scoreonlyS1_wip(seqs, wdw);
    return counted(true, 0, HookStats::HOM_WIP);
}

// ------------------------------------------------------------------------------ protein hooks
inline bool lspH(const Seq** seqs, const PwdB* pwd, const WINDOW& wdw, Mfile* mfd, int simd,
                 const Cip_score* cip, VTYPE* scr)
{
    const Seq* b = seqs[1];
    if (simd == 1 || !b->exin || !b->exin->data_p) return false;
    SpalnEngineH* e = engineH(pwd, b);
    const INT53* i53 = (b->inex.intr && e->same_sig53tab(sig53tab_of(b))) ? int53_of(b) : 0;
    if (simd == 0 && (!i53 || b->right - b->left >= MAX_SEGMENT)) return false;    // -A0 runs on the exact-ILD tables
    return counted(e->lspH_ng(seqs, wdw, mfd, i53, scr, cip), 1, HookStats::LSP);
}

inline bool trcbkH(const Seq** seqs, const PwdB* pwd, const WINDOW& wdw, Mfile* mfd, int simd,
                   const Cip_score* cip, bool spj, const RANGE* mc, VTYPE* scr)
{
    // trcbkalignH_ng (src/fwd2h1.cc:1997-2041): SIMD branch for -A2 / -A3, scalar branch (forwardH_ng
    // + Vmf) for -A0 and m < 8, there only with the splice switch the engine was frozen with
    const Seq* b = seqs[1];
    if (mc || wdw.width < 0 || !b->exin || !b->exin->data_p) return false;
    const int m = seqs[0]->right - seqs[0]->left;
    if (simd >= 2 && m >= 8) {
        *scr = engineH(pwd, b)->forwardH1_wip(seqs, wdw, mfd);
        return counted(true, 1, HookStats::TRCBK_WIP);
    }
    if ((simd != 0 && m >= 8) || m < 1 || b->right <= b->left || !b->inex.intr || !spj) return false;
    SpalnEngineH* e = engineH(pwd, b);
    if (!exact_tables_ok(b, e->same_sig53tab(sig53tab_of(b)))) return !counted(true, 1, HookStats::EXACT_NO_TABLES);
    if (counted(e->forwardH_ng(seqs, wdw, mfd, int53_of(b), scr, cip), 1, HookStats::TRCBK_EXACT)) return true;
    return !counted(true, 1, HookStats::EXACT_OVERFLOW);
}

inline bool homscoreH(const Seq** seqs, const PwdB* pwd, VTYPE* scr)
{
    const int simd = algmode.alg & 3;
    const Seq* a = seqs[0];
    const Seq* b = seqs[1];
    if (!b->exin || !b->exin->data_p) return false;
    if (simd == 3) IntronPrm.nquant = 1;
    WINDOW wdw;
    stripe31(seqs, &wdw, alprm.sh);
    if (simd == 0 || a->right - a->left < 8) {  // forwardH_ng's score (src/fwd2h1.cc:3297-3298)
        if (!b->inex.intr || a->right <= a->left || b->right <= b->left || wdw.up - wdw.lw + 7 < 0) return false;
        SpalnEngineH* e = engineH(pwd, b);
        if (!exact_tables_ok(b, e->same_sig53tab(sig53tab_of(b)))) return false;
        if (a->sigII) {
            const Cip_score cs(a);
            return counted(e->forwardH_ng(seqs, wdw, 0, int53_of(b), scr, &cs), 1, HookStats::HOM_EXACT);
        }
        return counted(e->forwardH_ng(seqs, wdw, 0, int53_of(b), scr), 1, HookStats::HOM_EXACT);
    }
    if (simd < 2) return false;                 // -A1: forwardH1, stock code
    *scr = engineH(pwd, b)->forwardH1_wip(seqs, wdw, 0);
    return counted(true, 1, HookStats::HOM_WIP);
}

#ifdef GSPALN_HARVEST
// ------------------------------------------------------------------------------------ harvest
// File = records { char tag[4]; int32 nbytes; payload }.
//   "PRMS" / "PRMH"  gspaln_params / gspaln_h_params, gspaln_lsp_opts, sig53tab[544], n_pen,
//                    penalty[n_pen], codonk1 | (H) spj_tabs[796], minl, ExtraGOP, GapW3L, Noll
//   "CALL"           a_len, b_len, a_left, a_right, b_left, b_right, exgl / exgr x 4, lw, up,
//                    a0, a1, b0, b1 (the slices that follow: residues [a0, a1) and [b0, b1),
//                    table columns [b0, b1 + 1]; everything outside reads as zero), a codes, b codes,
//                    sig5 + sig3 (DNA) or SGPT6 records (protein), INT53 words, score, n_skl, corners
struct Harvest {
    FILE* fp = 0;
    std::mutex mu;
    bool params_written = false;
    long long n_calls = 0;
    Harvest()
    {
        const char* fn = getenv("GSPALN_HARVEST_FILE");
        if (fn) fp = fopen(fn, "wb");
    }
    ~Harvest() { if (fp) fclose(fp); }
    void rec(const char* tag, const std::vector<char>& payload)
    {
        int n = (int) payload.size();
        fwrite(tag, 1, 4, fp);
        fwrite(&n, 4, 1, fp);
        fwrite(payload.data(), 1, payload.size(), fp);
    }
};
inline Harvest& harvest() { static Harvest h; return h; }

template <typename T> inline void put(std::vector<char>& v, const T* p, size_t n)
{
    const char* c = reinterpret_cast<const char*>(p);
    v.insert(v.end(), c, c + n * sizeof(T));
}
inline void put_i(std::vector<char>& v, int x) { put(v, &x, 1); }

// call with the harvest mutex held.  rng / exg: the ranges and end-gap flags at call time
inline void harvest_call(bool protein, const Seq** seqs, const PwdB* pwd, const WINDOW& wdw,
                         const RANGE* rng, const INT* exg, VTYPE scr, const SKL* skl, int n_skl)
{
    Harvest& H = harvest();
    if (!H.fp) return;
    const Seq* a = seqs[0];
    const Seq* b = seqs[1];
    const bool spliced = b->inex.intr;
    if (!H.params_written) {
        std::vector<char> v;
        const int n_pen = 1 << 20;      // (-A0 runs every block on the exact intron-length table)
        std::vector<short> pen(n_pen, 0);
        if (pwd->IntPen) for (int n = 0; n < n_pen; ++n) pen[n] = pwd->IntPen->Penalty(n);
        std::vector<short> tab(544, 0);
        if (sig53tab_of(b)) tab.assign(sig53tab_of(b), sig53tab_of(b) + 544);
        const gspaln_lsp_opts o = lsp_opts_now();
        if (protein) {
            gspaln_h_params p = SpalnEngineH::freeze(pwd, spliced);
            put(v, &p, 1); put(v, &o, 1); put(v, tab.data(), 544); put_i(v, n_pen); put(v, pen.data(), n_pen);
            unsigned char tabs[796];
            spj_tables(tabs);
            put(v, tabs, 796);
            put_i(v, IntronPrm.minl); put_i(v, pwd->ExtraGOP); put_i(v, pwd->GapW3L); put_i(v, pwd->Noll);
        } else {
            gspaln_params p = SpalnEngine::freeze(pwd, spliced);
            put(v, &p, 1); put(v, &o, 1); put(v, tab.data(), 544); put_i(v, n_pen); put(v, pen.data(), n_pen);
            put_i(v, pwd->codonk1);
        }
        H.rec(protein ? "PRMH" : "PRMS", v);
        H.params_written = true;
    }
    const int margin = protein ? 192 : 32;
    const int a0 = std::max(0, rng[0].left - 8), a1 = std::min(a->len, rng[0].right + 8);
    const int b0 = std::max(0, rng[1].left - margin), b1 = std::min(b->len, rng[1].right + margin);
    std::vector<char> v;
    put_i(v, a->len); put_i(v, b->len);
    put_i(v, rng[0].left); put_i(v, rng[0].right); put_i(v, rng[1].left); put_i(v, rng[1].right);
    for (int k = 0; k < 4; ++k) put_i(v, (int) exg[k]);
    put_i(v, wdw.lw); put_i(v, wdw.up);
    put_i(v, a0); put_i(v, a1); put_i(v, b0); put_i(v, b1);
    put(v, a->at(a0), a1 - a0);
    put(v, b->at(b0), b1 - b0);
    const INT53* i53 = int53_of(b);
    const int ncol = b1 + 1 - b0 + 1;       // columns b0 .. b1 + 1
    std::vector<unsigned short> w53((size_t) ncol, 0);
    auto pack53 = [&](int n) {
        const INT53& w = i53[n];
        return (unsigned short) (w.dinc5 | (w.dinc3 << 4) | (w.cano5 << 8) | (w.cano3 << 12));
    };
    // the Exinon tables cover the range the segment had when it was built; outside: zeros
    if (protein) {
        std::vector<gspaln_sgpt6> sg((size_t) ncol);
        memset(sg.data(), 0, sg.size() * sizeof(gspaln_sgpt6));
        const SGPT6* lo = b->exin->begin_p();
        const SGPT6* hi = b->exin->end_p();
        for (int n = b0; n <= b1 + 1; ++n) {
            const SGPT6* g = b->exin->score_p(n);
            if (g >= lo && g <= hi) {
                memcpy(&sg[n - b0], g, sizeof(gspaln_sgpt6));
                if (i53) w53[n - b0] = pack53(n);
            }
        }
        put(v, sg.data(), sg.size());
    } else {
        std::vector<short> s5((size_t) ncol, 0), s3((size_t) ncol, 0);
        const SGPT2* lo = b->exin ? b->exin->begin_n() : 0;
        const SGPT2* hi = b->exin ? b->exin->end_n() : 0;
        for (int n = b0; b->exin && n <= b1 + 1; ++n) {
            const SGPT2* g = b->exin->score_n(n);
            if (g >= lo && g <= hi) {
                s5[n - b0] = g->sig5; s3[n - b0] = g->sig3;
                if (i53) w53[n - b0] = pack53(n);
            }
        }
        put(v, s5.data(), s5.size()); put(v, s3.data(), s3.size());
    }
    put(v, w53.data(), w53.size());
    put_i(v, (int) scr); put_i(v, n_skl);
    for (int k = 0; k < n_skl; ++k) { put_i(v, skl[k].m); put_i(v, skl[k].n); }
    H.rec("CALL", v);
    ++H.n_calls;
}

// the records a stock lsp*_ng call appended to mfd, without disturbing it: Mfile has no read
// accessor, but a copy can be flushed
inline void harvest_after(bool protein, const Seq** seqs, const PwdB* pwd, const WINDOW& wdw,
                          const RANGE* rng, const INT* exg, VTYPE scr, Mfile* mfd, size_t mark)
{
    Mfile cp(*mfd);
    const size_t n = cp.size();
    SKL* all = (SKL*) cp.flush();
    std::lock_guard<std::mutex> lk(harvest().mu);
    harvest_call(protein, seqs, pwd, wdw, rng, exg, scr, all + mark, (int) (n - mark));
    delete[] all;
}

#define GSPALN_HARVEST_BODY(PROT, SELF_CALL)                                                         \
    {                                                                                                \
        static thread_local int gspaln_depth_ = 0;                                                   \
        if (!gspaln_depth_ && simd != 1 && !cip && gspaln::dropin::harvest().fp) {                   \
            RANGE rng_[2] = {{a->left, a->right}, {b->left, b->right}};                              \
            const INT exg_[4] = {a->inex.exgl, a->inex.exgr, b->inex.exgl, b->inex.exgr};            \
            const size_t mark_ = mfd->size();                                                        \
            ++gspaln_depth_;                                                                         \
            const VTYPE s_ = SELF_CALL;                                                              \
            --gspaln_depth_;                                                                         \
            gspaln::dropin::harvest_after(PROT, seqs, pwd, wdw, rng_, exg_, s_, mfd, mark_);         \
            return s_;                                                                               \
        }                                                                                            \
    }
#define GSPALN_HOOK_LSPS GSPALN_HARVEST_BODY(false, lspS_ng(wdw))
#define GSPALN_HOOK_LSPH GSPALN_HARVEST_BODY(true, lspH_ng(wdw))
#define GSPALN_HOOK_TRCBKS
#define GSPALN_HOOK_TRCBKH
#define GSPALN_HOOK_HOMS
#define GSPALN_HOOK_HOMH

#else   // the drop-in proper

#define GSPALN_HOOK_LSPS   { VTYPE s_; if (gspaln::dropin::lspS(seqs, pwd, wdw, mfd, simd, cip, &s_)) return s_; }
#define GSPALN_HOOK_TRCBKS { VTYPE s_; if (gspaln::dropin::trcbkS(seqs, pwd, wdw, mfd, simd, cip, mc, &s_)) return s_; }
#define GSPALN_HOOK_HOMS   { VTYPE s_; if (gspaln::dropin::homscoreS(seqs, pwd, &s_)) return s_; }
#define GSPALN_HOOK_LSPH   { VTYPE s_; if (gspaln::dropin::lspH(seqs, pwd, wdw, mfd, simd, cip, &s_)) return s_; }
#define GSPALN_HOOK_TRCBKH { VTYPE s_; if (gspaln::dropin::trcbkH(seqs, pwd, wdw, mfd, simd, cip, spj, mc, &s_)) return s_; }
#define GSPALN_HOOK_HOMH   { VTYPE s_; if (gspaln::dropin::homscoreH(seqs, pwd, &s_)) return s_; }

#endif  // GSPALN_HARVEST

}   // namespace dropin
}   // namespace gspaln
#endif
