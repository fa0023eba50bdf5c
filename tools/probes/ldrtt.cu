// ldrtt.cu -- round-trip time of L2 loads as the exchange issues them: one load at a time vs a batch of 16 independent
// 16-byte loads per thread, strong (relaxed.gpu) vs weak (.cg), 1 / 32 / 512 threads.  Data is resident in L2.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int STRONG>
__device__ __forceinline__ uint4 ld_v4(const uint4* p) {
    uint4 v;
    if (STRONG) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    else asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
template <int STRONG, int B>
__global__ void k(const uint4* buf, int T, int iters, long long* out, uint32_t* sink) {
    const int tid = threadIdx.x;
    if (tid >= T) return;
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint4 v[B];
        // the address of batch `it` depends on the previous batch's data (always 0): a dependent chain of batches
#pragma unroll
        for (int i = 0; i < B; ++i) v[i] = ld_v4<STRONG>(buf + (size_t)(i * 128 + (acc & 1)) * 8 + tid + (size_t)blockIdx.x * 4096);
#pragma unroll
        for (int i = 0; i < B; ++i) acc += v[i].x;
    }
    const long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 12345) *sink = acc;
}
template <int STRONG, int B>
static void run(const uint4* buf, long long* d, uint32_t* sink) {
    for (int G : {1, 32})
        for (int T : {1, 32, 512}) {
            const int iters = 2000;
            k<STRONG, B><<<G, 512>>>(buf, T, iters, d, sink);
            cudaDeviceSynchronize();
            long long h[32]; cudaMemcpy(h, d, 8 * G, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < G; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("%s batch=%2d CTAs=%2d threads=%3d: %.0f clk per batch\n", STRONG ? "strong" : "weak.cg", B, G, T, (double)mx / iters);
        }
    fflush(stdout);
}
int main() {
    uint4* buf; long long* d; uint32_t* sink;
    cudaMalloc(&buf, 64 << 20); cudaMemset(buf, 0, 64 << 20); cudaMalloc(&d, 8 * 64); cudaMalloc(&sink, 4);
    run<1, 1>(buf, d, sink); run<1, 4>(buf, d, sink); run<1, 16>(buf, d, sink);
    run<0, 1>(buf, d, sink); run<0, 16>(buf, d, sink);
    return 0;
}
