// gspaln_host.hpp -- small host-side helpers shared by the two halves of the C-ABI
// (gspaln.cu: DNA path, gspaln_h.cu: protein path): grow-only device / pinned buffers.
#pragma once
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

namespace gspaln {

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <typename T>
struct PinBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// sum over rows m in (a_left, a_right] of max(0, min(k m + U, R) - max(k m + L, B)): the band cells
// of a problem as the scalar reference loops count them (k = 1 DNA, 3 protein), in closed form
inline long long band_cells(int a_left, int a_right, int k, long long L, long long B, long long U, long long R)
{
    auto fdiv = [](long long x, long long y) { long long q = x / y; return (x % y != 0 && ((x < 0) != (y < 0))) ? q - 1 : q; };
    // breakpoints: hi = k m + U while m <= (R - U) / k; lo = k m + L once m >= (B - L) / k (ceil)
    const long long mh = fdiv(R - U, k);                // last row with hi == k m + U
    const long long ml = -fdiv(-(B - L), k);            // first row with lo == k m + L
    long long cuts[4] = {(long long) a_left + 1, mh + 1, ml, (long long) a_right + 1};
    if (cuts[1] > cuts[2]) std::swap(cuts[1], cuts[2]);
    long long total = 0;
    for (int s = 0; s < 3; ++s) {
        long long from = std::max(cuts[s], (long long) a_left + 1), to = std::min(cuts[s + 1] - 1, (long long) a_right);
        if (from > to) continue;
        const bool hi_lin = from <= mh && to <= mh, lo_lin = from >= ml;
        // f(m) = c0 + c1 m on [from, to]
        const long long c1 = (hi_lin ? k : 0) - (lo_lin ? k : 0);
        const long long c0 = (hi_lin ? U : R) - (lo_lin ? L : B);
        if (c1 > 0) from = std::max(from, fdiv(-c0, c1) + 1);
        else if (c1 < 0) to = std::min(to, fdiv(c0 - 1, -c1));
        else if (c0 <= 0) continue;
        if (from > to) continue;
        const long long cnt = to - from + 1;
        total += c0 * cnt + c1 * (from + to) * cnt / 2;
    }
    return total;
}

// ---------------------------------------------------------------------------
// Coalescing queue of the literal drop-in (include/gspaln.h: gspaln_queue_* / gspaln_h_queue_*).
// Spaln issues its DP problems one at a time from each pthread worker (src/spaln.cc:1363-1468);
// workers block in submit() with ONE problem, one dispatcher thread turns whatever is queued into
// batched calls of the engine: kernel-level problems (`SubmitFn`: gspaln_submit) and whole
// lsp*_ng calls (`LspFn`: gspaln_lsp, grouped by their options) are batched separately.
// ---------------------------------------------------------------------------
template <class Ctx, class Task, class Result, class LspOpts>
struct CoalescingQueue {
    using SubmitFn = int (*)(Ctx*, const Task*, int, Result*);
    using LspFn = int (*)(Ctx*, const Task*, int, const LspOpts*, Result*);
    struct Item {
        const Task* task;
        Result* result;
        const LspOpts* opts = nullptr;     // non-null: a driver (lsp) call
        int rc = 0;
        bool done = false;
    };
    Ctx* ctx = nullptr;
    SubmitFn submit_fn = nullptr;
    LspFn lsp_fn = nullptr;
    int einval = -1;
    int max_batch = 256, max_wait_us = 200;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<Item*> pending;
    bool stop = false;
    int64_t n_tasks = 0, n_batches = 0;
    std::thread worker;

    // runs items[idx...] as one engine call; a batch the engine refuses as a whole (one malformed
    // task) is re-run problem by problem so that only the offender sees the error
    void run_group(const std::vector<Item*>& take, const std::vector<size_t>& idx, std::vector<int>& rcs)
    {
        std::vector<Task> tasks(idx.size());
        std::vector<Result> results(idx.size());
        for (size_t k = 0; k < idx.size(); ++k) { tasks[k] = *take[idx[k]]->task; results[k] = *take[idx[k]]->result; }
        const LspOpts* o = take[idx[0]]->opts;
        auto call = [&](const Task* t, int n, Result* r) {
            return o ? lsp_fn(ctx, t, n, o, r) : submit_fn(ctx, t, n, r);
        };
        int rc = call(tasks.data(), (int) tasks.size(), results.data());
        for (size_t k = 0; k < idx.size(); ++k) rcs[idx[k]] = rc;
        if (rc == einval && idx.size() > 1)
            for (size_t k = 0; k < idx.size(); ++k) rcs[idx[k]] = call(&tasks[k], 1, &results[k]);
        for (size_t k = 0; k < idx.size(); ++k) *take[idx[k]]->result = results[k];
    }

    void run()
    {
        std::vector<Item*> take;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stop || !pending.empty(); });
                if (stop && pending.empty()) return;
                // stragglers: the other workers are usually a few microseconds behind
                if ((int) pending.size() < max_batch && max_wait_us > 0)
                    cv_work.wait_for(lk, std::chrono::microseconds(max_wait_us),
                                     [&] { return stop || (int) pending.size() >= max_batch; });
                take.clear();
                while (!pending.empty() && (int) take.size() < max_batch) {
                    take.push_back(pending.front());
                    pending.pop_front();
                }
            }
            std::vector<int> rcs(take.size(), 0);
            std::vector<char> seen(take.size(), 0);
            int groups = 0;
            for (size_t i = 0; i < take.size(); ++i) {
                if (seen[i]) continue;
                std::vector<size_t> idx;
                for (size_t j = i; j < take.size(); ++j) {
                    if (seen[j]) continue;
                    const bool same = (!take[i]->opts && !take[j]->opts) ||
                        (take[i]->opts && take[j]->opts && !memcmp(take[i]->opts, take[j]->opts, sizeof(LspOpts)));
                    if (same) { idx.push_back(j); seen[j] = 1; }
                }
                run_group(take, idx, rcs);
                ++groups;
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                for (size_t i = 0; i < take.size(); ++i) { take[i]->rc = rcs[i]; take[i]->done = true; }
                n_tasks += (int64_t) take.size();
                n_batches += groups;
            }
            cv_done.notify_all();
        }
    }

    void start() { worker = std::thread([this] { run(); }); }

    int submit(const Task* task, const LspOpts* opts, Result* result)
    {
        Item it;
        it.task = task; it.result = result; it.opts = opts;
        std::unique_lock<std::mutex> lk(mu);
        if (stop) return einval;
        pending.push_back(&it);
        cv_work.notify_one();
        cv_done.wait(lk, [&] { return it.done; });
        return it.rc;
    }

    void stats(int64_t* tasks, int64_t* batches)
    {
        std::lock_guard<std::mutex> lk(mu);
        if (tasks) *tasks = n_tasks;
        if (batches) *batches = n_batches;
    }

    void shutdown()
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_work.notify_all();
        if (worker.joinable()) worker.join();
    }
};

}   // namespace gspaln
