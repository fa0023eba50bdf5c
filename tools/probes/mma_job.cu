// mma_job.cu -- the loop's MMA job in isolation: 32 K steps (K = 512) with A in TMEM columns [0,256), D at column 256,
// B = [8 k-blocks][N rows][128 B] in shared memory; elected-lane issue.  Time per job with a commit + wait after every job
// (latency, what the loop sees) and back to back (throughput), for N = 32 / 64 / 96, and for A read from 4 vs 32 distinct
// column groups.
#include <cstdio>
#include <cstdlib>
#include "../../real-time-voice-cloning_b200/csrc/tc_common.cuh"
using namespace wrnn::tc;
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
                 "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(160, 1) k(int N, int iters, int mode, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 8 * 96 * 128 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tslot, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tslot;
    if (warp == 4) {
        const uint32_t idesc = umma_idesc_f16(128, N);
        const uint64_t bd0 = umma_desc_sw128(smem_u32(smem));
        const uint32_t kb_step = N * 8u;
        uint32_t ph = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
                uint64_t bd = bd0;
                uint32_t a = tmem;
#pragma unroll 1
                for (int kb = 0; kb < 8; ++kb) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        umma_ts(tmem + 256, (mode & 2) ? tmem + kk * 8 : a, bd + 2u * kk, idesc, (kb | kk) != 0 ? 1u : 0u);
                        a += 8u;
                    }
                    bd += kb_step;
                }
                if (mode & 1) umma_commit(&bar);
            }
            __syncwarp();
            if (mode & 1) { while (!mbar_try_wait(&bar, ph)) {} ph ^= 1; tcgen05_fence_after(); }
        }
        if (!(mode & 1)) { if (elect_one()) umma_commit(&bar); __syncwarp(); while (!mbar_try_wait(&bar, 0)) {} }
        const long long t1 = clock64();
        if ((tid & 31) == 0) out[0] = t1 - t0;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    const int smem = 1024 + 8 * 96 * 128;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int mode : {0, 1, 2, 3})
        for (int N : {32, 64, 96}) {
            const int iters = 500;
            k<<<1, 160, smem>>>(N, iters, mode, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("N=%2d %s, A from %s: %s  %.0f clk per 32-step job = %.1f clk per MMA (pipe floor %d)\n", N, (mode & 1) ? "commit+wait per job" : "back to back       ",
                   (mode & 2) ? "4 column groups " : "32 column groups", cudaGetErrorString(e), (double)h / iters, (double)h / iters / 32, N / 2);
        }
    return 0;
}
