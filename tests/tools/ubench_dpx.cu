// micro-benchmark: issue rate of the packed int16x2 DPX forms against their 32-bit twins on sm_100a
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/tools/ubench_dpx tests/tools/ubench_dpx.cu (binary is git-ignored; results: profiles/r01_ubench_dpx.txt)
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(unsigned* out, unsigned seed, int iters)
{
    unsigned a[8], b = seed ^ threadIdx.x, c = seed * 7u + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + i * 977u + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) a[i] = (unsigned) max((int) a[i], (int) b) + 1u;              // VIMNMX (+IADD)
            if (OP == 1) a[i] = __vmaxs2(a[i], b) + 1u;                                 // VIMNMX.S16x2
            if (OP == 2) a[i] = __vadd2(a[i], b);                                       // VIADD.16x2
            if (OP == 3) a[i] = __viaddmax_s16x2(a[i], b, c);                           // VIADDMNMX.S16x2
            if (OP == 4) a[i] = (unsigned) __viaddmax_s32((int) a[i], (int) b, (int) c); // VIADDMNMX
            if (OP == 5) a[i] = __byte_perm(a[i], b, 0x5410 + (it & 1));                // PRMT
            if (OP == 6) { bool p, q; a[i] = __vibmax_s16x2(a[i], b, &p, &q); c += p ? 1u : 0u; b += q ? 3u : 0u; }
        }
    }
    unsigned r = b ^ c;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int OP> void run(const char* name, unsigned* d)
{
    const int iters = 4096, grid = 148 * 8, block = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<grid, block>>>(d, 1234u, iters);
    cudaEventRecord(e0);
    k<OP><<<grid, block>>>(d, 1234u, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double) grid * block * iters * 8;
    printf("%-28s %8.3f ms  %7.1f Gops/s  (%.1f lane-ops/clk/SM at 1.965 GHz)\n", name, ms, ops / ms / 1e6,
           ops / ms / 1e6 / 148 / 1.965);
}
int main()
{
    unsigned* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("max.s32 + add (2 instr)", d);
    run<1>("vmaxs2 + add (2 instr)", d);
    run<2>("vadd2", d);
    run<3>("viaddmax_s16x2", d);
    run<4>("viaddmax_s32", d);
    run<5>("prmt", d);
    run<6>("vibmax_s16x2 + 2 pred uses", d);
    return 0;
}
