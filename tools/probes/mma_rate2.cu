// mma_rate2.cu -- is the ~118-clock cost of a K=16 tcgen05.mma a latency (hidden by independent accumulators or by
// several issuing threads) or an occupancy of the tensor pipe?  One CTA; W issuing warps (lane 0 each), each round-robins
// over NA accumulators; every MMA is M x N x 16 on zero operands in shared memory (SS) or TMEM (TS).
#include <cstdio>
#include <cstdlib>
#include "../../real-time-voice-cloning_b200/csrc/tc_common.cuh"
using namespace wrnn::tc;

template <bool kAcc>
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
                 "r"(tmem_a), "l"(bdesc), "r"(idesc) : "memory");
}

__global__ void __launch_bounds__(256, 1) rate_kernel(int M, int N, int W, int NA, int ts, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                 // 16 KB
    uint8_t* sB = smem + 16384;         // 32 KB
    __shared__ uint64_t bar[8];
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (16384 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tslot, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tslot;
    long long t0 = 0, t1 = 0, t2 = 0;
    if (warp < W && lane == 0) {
        const uint32_t idesc = umma_idesc_f16(M, N);
        const uint64_t ad = umma_desc_sw128(smem_u32(sA)), bd = umma_desc_sw128(smem_u32(sB));
        // accumulators: warp w uses columns [w*NA*N', ...) with N' = max(N,32); total must be <= 512 - 32 (A region for TS at 480..)
        const uint32_t np = N < 32 ? 32 : N;
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tmem + (uint32_t)((warp * NA + (i % NA)) * np) % 448u;
            if (ts) umma_ts<true>(d, tmem + 480 + (i & 3) * 8, umma_desc_advance(bd, (i & 3) * 32), idesc);
            else umma_f16_c<true>(d, umma_desc_advance(ad, (i & 3) * 32), umma_desc_advance(bd, (i & 3) * 32), idesc);
        }
        t1 = clock64();
        umma_commit(&bar[warp]);
        while (!mbar_try_wait(&bar[warp], 0)) {}
        t2 = clock64();
        out[warp * 2] = t1 - t0; out[warp * 2 + 1] = t2 - t0;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    const int smem = 1024 + 16384 + 32768;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 4096;
    for (int ts = 0; ts < 2; ++ts)
        for (int M : {128, 64})
            for (int N : {16, 64, 128, 256})
                for (int W : {1, 2, 4})
                    for (int NA : {1, 2, 4}) {
                        if ((long)W * NA * (N < 32 ? 32 : N) > 448 && !(W == 1 && NA == 1)) continue;
                        cudaMemset(d, 0, 64);
                        rate_kernel<<<1, 256, smem>>>(M, N, W, NA, ts, iters, d);
                        cudaError_t e = cudaDeviceSynchronize();
                        long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
                        long long iss = 0, tot = 0;
                        for (int w = 0; w < W; ++w) { if (h[2 * w] > iss) iss = h[2 * w]; if (h[2 * w + 1] > tot) tot = h[2 * w + 1]; }
                        printf("%s M=%3d N=%3d warps=%d accs/warp=%d: %s  issue %.1f clk/mma/warp, done %.1f clk per mma (all warps: %.1f clk per mma)\n", ts ? "TS" : "SS", M, N,
                               W, NA, cudaGetErrorString(e), (double)iss / iters, (double)tot / iters, (double)tot / (iters * W));
                    }
    return 0;
}
