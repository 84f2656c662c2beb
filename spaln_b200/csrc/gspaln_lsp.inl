// gspaln_lsp.inl -- host driver: Aln2s1::lspS_ng / Aln2h1::lspH_ng over a batch (included by
// gspaln.cu and gspaln_h.cu, which supply the traits of their sequence type).
//
// Reference: lspS_ng src/fwd2s1.cc:1801-1897, trcbkalignS_ng 1667-1710 (SIMD branch),
// mimd_postwork 1714-1756, rcsv_postwork 1758-1799, diagonalS_ng 1629-1665,
// stripe src/aln2.cc:156-176.  The reference recurses problem by problem; here
// every level of that recursion becomes one device batch (all Hirschberg passes
// and all block re-alignments of a level run together).  The protein driver
// (src/fwd2h1.cc:1963-2230) differs only in what the traits carry: stripe31, the rhombic
// volume m (n + 3 m), the lane count in the thresholds, the trivial-case penalties,
// diagonalH_ng and the range check of its mimd_postwork.

namespace {

struct LspGeo {
    int a_left, a_right, b_left, b_right;
    int a_exgl, a_exgr, b_exgl, b_exgr;
    int lw, up;
};

struct LspPiece {
    int kind;                   // 0 literal corners, 1 forward-task result, 2 child item
    int ref;
    std::vector<int2> lit;
};

struct LspItem {
    int root;                   // original problem
    LspGeo g;
    int score = 0;
    bool recursive = false;
    int n_imd = 0;              // intermediate rows of the Hirschberg pass
    int n_req = 0;              // ... before lspS_ng's even-division correction (what the scalar pass is given)
    std::vector<LspPiece> pieces;
};

struct LspFwd {                 // one trcbkalignS_ng call
    int root;
    int kind = GSPALN_FORWARD_WIP;  // GSPALN_FORWARD_NG for blocks with fewer than 8 query rows
    LspGeo g;
    int score = 0;
    std::vector<int2> skl;
};

}   // namespace

template <class TR>
int lsp_driver(typename TR::Ctx* ctx, const typename TR::Task* tasks, int n,
               const gspaln_lsp_opts* opts, gspaln_result* results)
{
    using Task = typename TR::Task;
    if (!ctx || !opts || n < 0 || (n && (!tasks || !results))) return GSPALN_EINVAL;
    const auto& P = ctx->prm;
    const int NEVSEL = INT_MIN / 16 * 7;
    const int EOU = INT_MAX - 2;
    // algmode.alg & 3: 2 / 3 = the `_wip` kernels; 0 = the scalar mode (-A0: exact-ILD kernels,
    // hexagonal volume, blocks banded by the diagonal bounds of the pass); 1 (-A1) is not on the device
    const int simd = opts->alg & 3;
    if (simd == 1 || (simd == 0 && !TR::SCALAR_MODE)) {
        for (int i = 0; i < n; ++i) {
            results[i].score = NEVSEL; results[i].status = GSPALN_ST_UNSUPPORTED; results[i].n_skl = 0;
            results[i].reserved = 0; results[i].cells = TR::cells(tasks[i]);
        }
        return GSPALN_OK;
    }
    // band of a block after a Hirschberg pass (src/fwd2s1.cc:1736-1741)
    auto window = [&](LspGeo& g, const int* row) {
        if (simd) TR::stripe(g, opts->sh);
        else { g.lw = row[8]; g.up = row[9]; }
    };
    std::vector<LspItem> items;
    std::vector<LspFwd> fwds;
    std::vector<int> status(n, GSPALN_ST_OK);
    std::vector<int> pending;
    items.reserve(2 * (size_t) n);
    for (int i = 0; i < n; ++i) {
        const Task& t = tasks[i];
        LspItem it;
        it.root = i;
        it.g = LspGeo{t.a_left, t.a_right, t.b_left, t.b_right, t.a_exgl, t.a_exgr, t.b_exgl, t.b_exgr, t.lw, t.up};
        items.push_back(it);
        pending.push_back(i);
    }
    auto lit2 = [](LspItem& it, int m0, int n0, int m1, int n1) {
        LspPiece p; p.kind = 0; p.ref = -1;
        p.lit.push_back(make_int2(m0, n0)); p.lit.push_back(make_int2(m1, n1));
        it.pieces.push_back(std::move(p));
    };
    // returns the index of the queued forward task or -1 (nothing to do / unsupported)
    auto queue_trcbk = [&](int item, const LspGeo& g) -> int {
        if (g.up - g.lw + TR::WPAD < 0) return -1;
        const bool scalar = simd == 0 || g.a_right - g.a_left < 8;     // src/fwd2s1.cc:1676, src/fwd2h1.cc:2007
        if ((scalar && !TR::scalar_ok(ctx, tasks[items[item].root], g)) || g.b_right < g.b_left ||
            g.a_left < 0 || g.b_left < 0 || TR::beyond(tasks[items[item].root], g)) {
            status[items[item].root] = GSPALN_ST_UNSUPPORTED;
            return -1;
        }
        LspFwd f; f.root = items[item].root; f.g = g;
        if (scalar) f.kind = GSPALN_FORWARD_NG;
        fwds.push_back(std::move(f));
        LspPiece p; p.kind = 1; p.ref = (int) fwds.size() - 1;
        items[item].pieces.push_back(std::move(p));
        return p.ref;
    };
    size_t fwd_done = 0;
    int64_t cells_total = 0;
    // Corner buffers of the block re-alignments: a block of m rows and n columns can in theory
    // return m + n corners, in practice a few dozen.  The first submit gives every forward task
    // LSP_SKL_CAP corners (the pool of the whole level is copied back, so its size matters); a
    // task that reports GSPALN_ST_SKL_OVERFLOW is run again with the size it asked for.
    constexpr int LSP_SKL_CAP = 192;
    std::vector<std::vector<int>> retry_buf;
    auto retry_overflow = [&](std::vector<Task>& bt, std::vector<gspaln_result>& br, size_t first) -> int {
        std::vector<size_t> again;
        for (size_t k = first; k < bt.size(); ++k)
            if (br[k].status == GSPALN_ST_SKL_OVERFLOW) again.push_back(k);
        if (again.empty()) return GSPALN_OK;
        std::vector<Task> t2;
        std::vector<gspaln_result> r2(again.size());
        const size_t keep = retry_buf.size();
        retry_buf.resize(keep + again.size());
        for (size_t j = 0; j < again.size(); ++j) {
            Task t = bt[again[j]];
            t.skl_cap = br[again[j]].n_skl + 8;
            retry_buf[keep + j].assign(2 * (size_t) t.skl_cap, 0);
            r2[j].skl = retry_buf[keep + j].data();
            r2[j].cpos = nullptr;
            t2.push_back(t);
        }
        int rc = TR::submit(ctx, t2.data(), (int) t2.size(), r2.data());
        if (rc != GSPALN_OK) return rc;
        for (size_t j = 0; j < again.size(); ++j) {
            bt[again[j]].skl_cap = t2[j].skl_cap;
            br[again[j]] = r2[j];
        }
        return GSPALN_OK;
    };
    const bool dbg = getenv("GSPALN_LSP_DEBUG") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto msec = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    double t_classify = 0, t_submit = 0, t_post = 0;
    auto t_all = now();
    float kernel_ms = 0, h2d_ms = 0, d2h_ms = 0;
    int launches = 0;
    int64_t h2d_bytes = 0, d2h_bytes = 0;

    while (!pending.empty()) {
        auto t0 = now();
        // ---- classify (lspS_ng head) and build this level's batch
        std::vector<int> udh_items;         // items waiting for a Hirschberg pass
        std::vector<int> score_from_fwd;    // (item, fwd) pairs: item score = forward score
        std::vector<std::pair<int, int>> item_fwd;
        for (int id : pending) {
            LspItem& it = items[id];
            const LspGeo& g = it.g;
            const Task& base = tasks[it.root];
            const int m = g.a_right - g.a_left, nn = g.b_right - g.b_left;
            if (m < 0 || nn < 0 || g.a_left < 0 || g.b_left < 0 || TR::beyond(base, g)) {
                // a range a Hirschberg pass narrowed to outside the sequences (the reference reads
                // foreign memory from here on, e.g. fhlastH1's start point right of b_right)
                status[it.root] = GSPALN_ST_UNSUPPORTED;
                it.score = NEVSEL;
                continue;
            }
            if (!m && !nn) { it.score = 0; continue; }
            if (!m || !nn) {
                lit2(it, g.a_left, g.b_left, g.a_right, g.b_right);
                it.score = TR::trivial_score(P, g, m, nn);
                continue;
            }
            if (g.up == g.lw) {
                int c4[4], scr = 0;
                TR::diagonal(P, base, g, c4, scr);
                lit2(it, c4[0], c4[1], c4[2], c4[3]);
                it.score = scr;
                continue;
            }
            bool trcbk = TR::small(m, nn);
            int n_imd = 1, n_req = 1;
            bool recursive = (opts->alg & 4) != 0;
            const float coef_B = 2.f, coef_C = TR::coef_c(P);
            float cvol = TR::cvol(m, nn);                       // rhombic (simd >= 2)
            if (simd < 2) cvol = TR::cvol_hex(g, m, nn);         // hexagonal (src/fwd2s1.cc:1830-1833, src/fwd2h1.cc:2161-2164)
            if (!trcbk && coef_B * cvol < opts->max_vmf_space) trcbk = true;
            if (!trcbk && !recursive) {
                const double z = 2. * m * coef_B / coef_C;
                const int imd1 = (int) (pow(z, 1. / 3) + 0.5) - 1;
                const float spc = coef_C * nn * imd1 + coef_B * cvol / (imd1 + 1) / (imd1 + 1);
                if (spc > opts->max_vmf_space) recursive = true;
                else {
                    const int imd3 = m / NELEM;
                    n_imd = opts->ubh ? opts->ubh : std::min(imd1, imd3);
                    n_req = n_imd;
                    const int imd_intvl = (m + n_imd) / (n_imd + 1);
                    if (imd_intvl * n_imd == m) --n_imd;
                    if (n_imd == 0) trcbk = true;
                }
            }
            if (trcbk) {
                const int f = queue_trcbk(id, g);
                if (f >= 0) item_fwd.push_back({id, f}); else it.score = NEVSEL;
                continue;
            }
            if (simd ? !TR::udh_ok(P) : !TR::scalar_ok(ctx, base, g)) {     // the Hirschberg route needs a pass that is not on the device
                status[it.root] = GSPALN_ST_UNSUPPORTED;
                it.score = NEVSEL;
                continue;
            }
            it.recursive = recursive;
            it.n_imd = n_imd;
            it.n_req = n_req;
            udh_items.push_back(id);
        }
        pending.clear();

        // ---- one device batch: all Hirschberg passes + all queued trace-back problems
        std::vector<Task> batch;
        std::vector<gspaln_result> bres;
        std::vector<std::vector<int>> cposbuf(udh_items.size());
        for (size_t k = 0; k < udh_items.size(); ++k) {
            const LspItem& it = items[udh_items[k]];
            if (simd) batch.push_back(TR::make_task(tasks[it.root], it.g, GSPALN_HIRSCHBERG_WIP, it.n_imd));
            else batch.push_back(TR::make_task(tasks[it.root], it.g, TR::KIND_SCALAR_UDH, it.n_req));
            cposbuf[k].assign(10 * (size_t) (std::max(it.n_imd, it.n_req) + 1), 0);
        }
        const size_t fwd_first = fwd_done;
        size_t skl_ints = 0;
        for (size_t f = fwd_first; f < fwds.size(); ++f) {
            batch.push_back(TR::make_task(tasks[fwds[f].root], fwds[f].g, fwds[f].kind, 0));
            batch.back().skl_cap = std::min(batch.back().skl_cap, LSP_SKL_CAP);
            skl_ints += 2 * (size_t) batch.back().skl_cap;
        }
        std::vector<int> sklflat(skl_ints + 2);
        bres.resize(batch.size());
        for (size_t k = 0; k < udh_items.size(); ++k) { bres[k].skl = nullptr; bres[k].cpos = cposbuf[k].data(); }
        {
            size_t at = 0;
            for (size_t f = fwd_first; f < fwds.size(); ++f) {
                const size_t k = udh_items.size() + (f - fwd_first);
                bres[k].skl = sklflat.data() + at;
                bres[k].cpos = nullptr;
                at += 2 * (size_t) batch[k].skl_cap;
            }
        }
        auto t1 = now();
        t_classify += msec(t0, t1);
        if (!batch.empty()) {
            int rc = TR::submit(ctx, batch.data(), (int) batch.size(), bres.data());
            if (rc != GSPALN_OK) return rc;
            t_submit += msec(t1, now());
            if (dbg) fprintf(stderr, "gspaln lsp: level of %zu tasks (%zu Hirschberg passes): %.1f ms, kernels %.1f ms, %.3g cells\n",
                             batch.size(), udh_items.size(), msec(t1, now()), ctx->tim.kernel_ms, (double) ctx->tim.cells);
            kernel_ms += ctx->tim.kernel_ms; h2d_ms += ctx->tim.h2d_ms; d2h_ms += ctx->tim.d2h_ms;
            launches += ctx->tim.launches; h2d_bytes += ctx->tim.h2d_bytes; d2h_bytes += ctx->tim.d2h_bytes;
            cells_total += ctx->tim.cells;
            rc = retry_overflow(batch, bres, udh_items.size());
            if (rc != GSPALN_OK) return rc;
        }
        // forward results
        for (size_t f = fwd_first; f < fwds.size(); ++f) {
            const gspaln_result& r = bres[udh_items.size() + (f - fwd_first)];
            fwds[f].score = r.score;
            if (r.status != GSPALN_ST_OK) status[fwds[f].root] = r.status;
            const int cnt = std::min(r.n_skl, batch[udh_items.size() + (f - fwd_first)].skl_cap);
            const int* s = r.skl;
            fwds[f].skl.resize(std::max(cnt, 0));
            for (int k = 0; k < cnt; ++k) fwds[f].skl[k] = make_int2(s[2 * k], s[2 * k + 1]);
        }
        fwd_done = fwds.size();
        for (auto& pr : item_fwd) items[pr.first].score = fwds[pr.second].score;

        auto t2 = now();
        // ---- post-work of the Hirschberg passes: next level
        for (size_t k = 0; k < udh_items.size(); ++k) {
            const int id = udh_items[k];
            const gspaln_result& r = bres[k];
            const int* cpos = cposbuf[k].data();
            items[id].score = r.score;
            if (r.status != GSPALN_ST_OK) {
                // the pass did not run to its end (e.g. its inputs never arrived): no post-work on
                // records that were not written
                status[items[id].root] = r.status;
                items[id].score = NEVSEL;
                continue;
            }
            if (!(r.score > NEVSEL)) continue;
            LspGeo g = items[id].g;
            g.a_left = r.ranges[0]; g.a_right = r.ranges[1]; g.b_left = r.ranges[2]; g.b_right = r.ranges[3];
            const int n_imd = items[id].n_imd;
            if (cpos[0] == EOU) {
                lit2(items[id], g.a_left, g.b_left, g.a_right, g.b_right);
            } else if (items[id].recursive) {
                // rcsv_postwork
                g.a_exgl = g.b_exgl = g.a_exgr = g.b_exgr = 0;
                int c = 0;
                if (cpos[c++] < EOU) {
                    LspPiece p; p.kind = 0; p.ref = -1;
                    while (cpos[++c] < EOU) p.lit.push_back(make_int2(cpos[0], cpos[c]));
                    items[id].pieces.push_back(std::move(p));
                    const int aright = g.a_right, bright = g.b_right;
                    LspGeo g1 = g;
                    g1.a_right = cpos[0]; g1.b_right = cpos[c - 1];
                    window(g1, cpos);
                    LspGeo g2 = g1;
                    g2.a_left = cpos[0]; g2.b_exgl = cpos[1]; g2.b_left = cpos[2];
                    g2.a_right = aright; g2.b_right = bright;
                    window(g2, cpos + 10);
                    for (const LspGeo& gg : {g1, g2}) {
                        LspItem child; child.root = items[id].root; child.g = gg;
                        items.push_back(child);
                        const int cid = (int) items.size() - 1;
                        LspPiece cp; cp.kind = 2; cp.ref = cid;
                        items[id].pieces.push_back(std::move(cp));
                        pending.push_back(cid);
                    }
                } else if (TR::is_local(P)) {
                    TR::stripe(g, opts->sh);
                    queue_trcbk(id, g);
                }
            } else {
                // mimd_postwork
                const int aleft = g.a_left, bleft = g.b_left;
                g.a_exgl = g.b_exgl = g.a_exgr = g.b_exgr = 0;
                int i = n_imd;
                while (--i >= 0 && cpos[10 * i] == EOU) ;
                for ( ; i >= 0 && cpos[10 * i] != EOU; --i) {
                    int c = 0;
                    g.a_left = cpos[10 * i + c];
                    g.b_exgl = cpos[10 * i + (++c)];
                    g.b_left = cpos[10 * i + (++c)];
                    if (TR::bad_range(tasks[items[id].root], g)) { i = -2; break; }    // mimd_postwork returns
                    if (g.b_left < 0 || g.b_left > g.b_right) break;
                    LspPiece p; p.kind = 0; p.ref = -1;
                    while (cpos[10 * i + (++c)] < EOU) p.lit.push_back(make_int2(g.a_left, cpos[10 * i + c]));
                    if (!p.lit.empty()) items[id].pieces.push_back(std::move(p));
                    window(g, cpos + 10 * (i + 1));
                    queue_trcbk(id, g);
                    g.a_right = g.a_left;
                    g.b_right = cpos[10 * i + c - 1];
                }
                if (i != -2 && ((i < 0 && cpos[0] != EOU) || cpos[2] != EOU)) {
                    g.a_left = aleft;
                    g.b_left = bleft;
                    window(g, cpos);
                    queue_trcbk(id, g);
                }
            }
        }
        t_post += msec(t2, now());
        // block re-alignments queued by the post-work run with the next level
        if (pending.empty() && fwd_done < fwds.size()) pending.push_back(-1);
        if (!pending.empty() && pending.back() == -1) {
            pending.pop_back();
            // a level with only forward tasks: loop once more with no items to classify
            std::vector<Task> b2;
            std::vector<gspaln_result> r2;
            size_t ints2 = 0;
            for (size_t f = fwd_done; f < fwds.size(); ++f) {
                b2.push_back(TR::make_task(tasks[fwds[f].root], fwds[f].g, fwds[f].kind, 0));
                b2.back().skl_cap = std::min(b2.back().skl_cap, LSP_SKL_CAP);
                ints2 += 2 * (size_t) b2.back().skl_cap;
            }
            std::vector<int> flat2(ints2 + 2);
            r2.resize(b2.size());
            {
                size_t at = 0;
                for (size_t f = 0; f < b2.size(); ++f) {
                    r2[f].skl = flat2.data() + at;
                    r2[f].cpos = nullptr;
                    at += 2 * (size_t) b2[f].skl_cap;
                }
            }
            auto t3 = now();
            int rc = TR::submit(ctx, b2.data(), (int) b2.size(), r2.data());
            if (rc != GSPALN_OK) return rc;
            t_submit += msec(t3, now());
            if (dbg) fprintf(stderr, "gspaln lsp: %zu block trace-backs: %.1f ms, kernels %.1f ms, %.3g cells\n",
                             b2.size(), msec(t3, now()), ctx->tim.kernel_ms, (double) ctx->tim.cells);
            kernel_ms += ctx->tim.kernel_ms; h2d_ms += ctx->tim.h2d_ms; d2h_ms += ctx->tim.d2h_ms;
            launches += ctx->tim.launches; h2d_bytes += ctx->tim.h2d_bytes; d2h_bytes += ctx->tim.d2h_bytes;
            cells_total += ctx->tim.cells;
            rc = retry_overflow(b2, r2, 0);
            if (rc != GSPALN_OK) return rc;
            for (size_t f = 0; f < b2.size(); ++f) {
                LspFwd& F = fwds[fwd_done + f];
                F.score = r2[f].score;
                if (r2[f].status != GSPALN_ST_OK) status[F.root] = r2[f].status;
                const int cnt = std::min(r2[f].n_skl, b2[f].skl_cap);
                F.skl.resize(std::max(cnt, 0));
                for (int k = 0; k < cnt; ++k) F.skl[k] = make_int2(r2[f].skl[2 * k], r2[f].skl[2 * k + 1]);
            }
            fwd_done = fwds.size();
        }
    }

    // ---- assemble the corner lists in the reference's write order (depth first); the problems
    // are independent, so the loop is dealt to a few host threads
    auto assemble = [&](int i) {
        gspaln_result& o = results[i];
        std::vector<int2> out;
        std::vector<std::pair<int, size_t>> stack;      // (item, next piece)
        stack.push_back({i, 0});
        while (!stack.empty()) {
            auto& top = stack.back();
            const LspItem& it = items[top.first];
            if (top.second >= it.pieces.size()) { stack.pop_back(); continue; }
            const LspPiece& p = it.pieces[top.second++];
            if (p.kind == 0) out.insert(out.end(), p.lit.begin(), p.lit.end());
            else if (p.kind == 1) out.insert(out.end(), fwds[p.ref].skl.begin(), fwds[p.ref].skl.end());
            else stack.push_back({p.ref, 0});
        }
        o.score = items[i].score;
        o.status = status[i];
        o.n_skl = (int) out.size();
        o.reserved = 0;
        o.cells = TR::cells(tasks[i]);
        const int cap = tasks[i].skl_cap;
        if (o.status == GSPALN_ST_OK && o.n_skl > cap) o.status = GSPALN_ST_SKL_OVERFLOW;
        if (o.skl)
            for (int k = 0; k < std::min(o.n_skl, cap); ++k) { o.skl[2 * k] = out[k].x; o.skl[2 * k + 1] = out[k].y; }
    };
    {
        const int nthr = n >= 512 ? (int) std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u) : 1;
        if (nthr == 1) {
            for (int i = 0; i < n; ++i) assemble(i);
        } else {
            std::vector<std::thread> pool;
            for (int w = 0; w < nthr; ++w)
                pool.emplace_back([&, w] { for (int i = w; i < n; i += nthr) assemble(i); });
            for (auto& th : pool) th.join();
        }
    }
    if (dbg)
        fprintf(stderr, "gspaln lsp: %d problems, %zu forward tasks; classify %.1f ms, submits %.1f ms "
                "(kernels %.1f), post-work %.1f ms, total %.1f ms\n", n, fwds.size(), t_classify, t_submit,
                kernel_ms, t_post, msec(t_all, now()));
    ctx->tim.kernel_ms = kernel_ms; ctx->tim.h2d_ms = h2d_ms; ctx->tim.d2h_ms = d2h_ms;
    ctx->tim.launches = launches; ctx->tim.h2d_bytes = h2d_bytes; ctx->tim.d2h_bytes = d2h_bytes;
    ctx->tim.cells = cells_total;
    return GSPALN_OK;
}
