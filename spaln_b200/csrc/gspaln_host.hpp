// gspaln_host.hpp -- small host-side helpers shared by the two halves of the C-ABI
// (gspaln.cu: DNA path, gspaln_h.cu: protein path): grow-only device / pinned buffers.
#pragma once
#include <algorithm>
#include <cstddef>
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

}   // namespace gspaln
