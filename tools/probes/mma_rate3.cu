// mma_rate3.cu -- the ~118 clocks per tcgen05.mma measured in round 1 are not a tensor-pipe floor: under `if (lane == 0)`
// ptxas cannot prove the operands warp-uniform and wraps every UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY
// waterfall.  This probe issues the same MMA stream (a) under lane == 0, (b) under elect.sync, for several N.
#include <cstdio>
#include <cstdlib>
#include "../../real-time-voice-cloning_b200/csrc/tc_common.cuh"
using namespace wrnn::tc;

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_ts_acc(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
                 "r"(tmem_a), "l"(bdesc), "r"(idesc) : "memory");
}

template <int MODE, int TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (16384 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tslot, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tslot;
    if (warp == 1) {
        const uint32_t idesc = umma_idesc_f16(128, N);
        const uint64_t ad = umma_desc_sw128(smem_u32(smem)), bd = umma_desc_sw128(smem_u32(smem + 16384));
        long long t0 = 0, t1 = 0, t2 = 0;
        if (MODE == 0) {
            if (lane == 0) {
                t0 = clock64();
                for (int i = 0; i < iters; ++i) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (TS) umma_ts_acc(tmem, tmem + 256 + k * 8, umma_desc_advance(bd, k * 32), idesc);
                        else umma_f16_c<true>(tmem, umma_desc_advance(ad, k * 32), umma_desc_advance(bd, k * 32), idesc);
                    }
                }
                t1 = clock64();
                umma_commit(&bar);
                while (!mbar_try_wait(&bar, 0)) {}
                t2 = clock64();
                out[0] = t1 - t0; out[1] = t2 - t0;
            }
        } else {
            // the whole warp walks the loop; one elected lane issues
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (TS) umma_ts_acc(tmem, tmem + 256 + k * 8, umma_desc_advance(bd, k * 32), idesc);
                        else umma_f16_c<true>(tmem, umma_desc_advance(ad, k * 32), umma_desc_advance(bd, k * 32), idesc);
                    }
                }
                __syncwarp();
            }
            t1 = clock64();
            if (elect_one()) umma_commit(&bar);
            __syncwarp();
            while (!mbar_try_wait(&bar, 0)) {}
            t2 = clock64();
            if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int MODE, int TS>
static void run(long long* d) {
    const int smem = 1024 + 16384 + 32768;
    cudaFuncSetAttribute(rate_kernel<MODE, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2048;
    for (int N : {16, 32, 64, 96, 128, 192, 256}) {
        cudaMemset(d, 0, 64);
        rate_kernel<MODE, TS><<<1, 128, smem>>>(N, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%s %s N=%3d: %s issue %.1f clk/mma, done %.1f clk/mma (tensor-pipe floor N/2 = %d)\n", TS ? "TS" : "SS", MODE ? "elect.sync" : "lane==0   ", N,
               cudaGetErrorString(e), (double)h[0] / (4 * iters), (double)h[1] / (4 * iters), N / 2);
    }
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    run<0, 0>(d); run<1, 0>(d); run<0, 1>(d); run<1, 1>(d);
    return 0;
}
