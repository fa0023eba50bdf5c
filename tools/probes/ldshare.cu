// ldshare.cu -- does it cost more when G CTAs read the SAME 64 KB (one exchange matrix) at the same time than when each
// reads its own copy?  512 threads x 16 independent 16-byte strong loads per pass, static data resident in L2, all CTAs
// released together by a grid barrier before every pass.
#include <cstdio>
#include <cstdint>
#include <cooperative_groups.h>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
__device__ __forceinline__ uint4 ld_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__global__ void __launch_bounds__(512, 1) k(const uint4* buf, int shared_lines, int iters, long long* out, uint32_t* sink) {
    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x, row = tid & 127, cgp = tid >> 7;
    const uint4* base = buf + (shared_lines ? 0 : (size_t)blockIdx.x * 64 * 128) + (size_t)(cgp * 16) * 128 + row;
    uint32_t acc = 0;
    long long tot = 0;
    for (int it = 0; it < iters; ++it) {
        grid.sync();
        const long long t0 = clock64();
        uint4 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = ld_v4(base + i * 128);
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += v[i].x;
        __syncthreads();
        tot += clock64() - t0;
    }
    if (tid == 0) out[blockIdx.x] = tot;
    if (acc == 12345) *sink = acc;
}
int main() {
    uint4* buf; long long* d; uint32_t* sink;
    cudaMalloc(&buf, (size_t)148 * 64 * 128 * 16); cudaMemset(buf, 0, (size_t)148 * 64 * 128 * 16); cudaMalloc(&d, 8 * 148); cudaMalloc(&sink, 4);
    for (int sh : {0, 1})
        for (int G : {1, 8, 16, 32, 64, 128}) {
            int iters = 200;
            void* args[] = {&buf, (void*)&sh, (void*)&iters, &d, &sink};
            cudaLaunchCooperativeKernel((const void*)k, dim3(G), dim3(512), args, 0, 0);
            cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, d, 8 * G, cudaMemcpyDeviceToHost);
            long long mx = 0, sum = 0; for (int i = 0; i < G; ++i) { mx = h[i] > mx ? h[i] : mx; sum += h[i]; }
            printf("%s CTAs=%3d: CTA-level pass (128 KB per CTA) %.0f clk mean, %.0f clk slowest CTA\n", sh ? "same lines   " : "private lines", G, (double)sum / G / iters, (double)mx / iters);
            fflush(stdout);
        }
    return 0;
}
