// tc_gemm_test.cu -- self-test of the tensor-core building blocks used by the fp16 loop kernel:
// C[128][N] = A[128][512] * W[N][512]^T with A arriving by TMA (128B swizzle) through a 4-slot mbarrier
// ring, W staged once in shared memory in the canonical K-major SWIZZLE_128B layout, tcgen05.mma
// accumulating in TMEM and tcgen05.ld bringing the tile back.  One CTA; exposed as wrnn_debug_tc_gemm.
#include "engine_internal.h"
#include "tc_common.cuh"

namespace wrnn {
namespace {
using namespace tc;

constexpr int kM = 128, kK = 512, kKB = 64, kNKB = kK / kKB, kSlots = 4;
constexpr long long kDeadline = 400000000LL;

__device__ __forceinline__ bool wait_bar(uint64_t* bar, uint32_t parity) {
    long long t0 = 0;
    int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (((++spins) & 255) == 0) {
            if (t0 == 0) t0 = clock64();
            if (clock64() - t0 > kDeadline) return false;
        }
    }
    return true;
}

__global__ void __launch_bounds__(192, 1) tc_gemm_test_kernel(const __grid_constant__ CUtensorMap tmapA,
                                                              const __half* __restrict__ W, int N, float* __restrict__ C,
                                                              int* __restrict__ status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = base;                                   // kSlots x [128 rows x 128 B]
    uint8_t* sW = sA + kSlots * kM * 128;                 // kNKB x [N rows x 128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + kNKB * N * 128);
    uint64_t* full = bars;                                // [kSlots]
    uint64_t* empty = bars + kSlots;                      // [kSlots]
    uint64_t* accfull = bars + 2 * kSlots;                // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kSlots + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // stage W (generic proxy) in the swizzled layout
    for (int i = tid; i < N * kNKB * 8; i += blockDim.x) {
        const int n = i / (kNKB * 8), rem = i % (kNKB * 8), kb = rem / 8, c = rem % 8;
        const uint4 v = *reinterpret_cast<const uint4*>(W + (size_t)n * kK + kb * kKB + c * 8);
        *reinterpret_cast<uint4*>(sW + kb * N * 128 + sw128_offset(n, c)) = v;
    }
    fence_proxy_async_smem();
    if (tid == 0) {
        for (int i = 0; i < kSlots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(accfull, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 64);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 4) {                       // ---- TMA producer ----
        if (lane == 0) {
            tma_prefetch_desc(&tmapA);
            for (int kb = 0; kb < kNKB; ++kb) {
                const int slot = kb % kSlots, round = kb / kSlots;
                if (round > 0 && !wait_bar(&empty[slot], (round - 1) & 1)) { atomicExch(status, 2); break; }
                mbar_arrive_expect_tx(&full[slot], kM * 128);
                tma_load_2d(sA + slot * kM * 128, &tmapA, &full[slot], kb * kKB, 0);
            }
        }
    } else if (warp == 5) {                // ---- MMA issuer ----
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(kM, N);
            for (int kb = 0; kb < kNKB; ++kb) {
                const int slot = kb % kSlots, round = kb / kSlots;
                if (!wait_bar(&full[slot], round & 1)) { atomicExch(status, 3); break; }
                tcgen05_fence_after();
                const uint32_t a0 = smem_u32(sA + slot * kM * 128), b0 = smem_u32(sW + kb * N * 128);
#pragma unroll
                for (int j = 0; j < kKB / 16; ++j)
                    umma_f16(tmem, umma_desc_sw128(a0 + j * 32), umma_desc_sw128(b0 + j * 32), idesc, (kb | j) ? 1u : 0u);
                umma_commit(&empty[slot]);
            }
            umma_commit(accfull);
        }
    } else {                               // ---- epilogue warps 0..3: TMEM lanes 32*warp .. +31 ----
        const bool ok = wait_bar(accfull, 0);
        tcgen05_fence_after();
        if (!ok) { if (lane == 0) atomicExch(status, 4); }
        else {
            const int row = warp * 32 + lane;
            for (int c0 = 0; c0 < N; c0 += 8) {
                float v[8];
                tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) C[(size_t)row * N + c0 + i] = v[i];
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

// Self-test of the CTA-pair path (cta_group::2): C[256][N] = A[256][512] * W[N][512]^T on one 2-CTA cluster.  CTA r loads A
// rows [128 r, 128 r + 128) (two k-blocks per TMA operation, 3-slot ring, completion on the even CTA's barriers) and holds W
// rows [N/2 r, N/2 r + N/2); the even CTA issues tcgen05.mma.cta_group::2 (M = 256) and multicasts the commits; each CTA
// reads its 128 rows of the accumulator from its own TMEM.
constexpr int kSlots2 = 3, kSlotBytes2 = 2 * kM * 128;
__global__ void __launch_bounds__(192, 1) tc_gemm2_test_kernel(const __grid_constant__ CUtensorMap tmapA, const __half* __restrict__ W, int N,
                                                               float* __restrict__ C, int* __restrict__ status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = base;                                   // kSlots2 x 2 x [128 rows x 128 B]
    uint8_t* sW = sA + kSlots2 * kSlotBytes2;             // kNKB x [N/2 rows x 128 B]
    const int NH = N / 2;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + kNKB * NH * 128);
    uint64_t* full = bars;                                // [kSlots2]  (used in the even CTA)
    uint64_t* empty = bars + kSlots2;                     // [kSlots2]
    uint64_t* accfull = bars + 2 * kSlots2;               // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kSlots2 + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    for (int i = tid; i < NH * kNKB * 8; i += blockDim.x) {
        const int n = i / (kNKB * 8), rem = i % (kNKB * 8), kb = rem / 8, c = rem % 8;
        const uint4 v = *reinterpret_cast<const uint4*>(W + (size_t)(rank * NH + n) * kK + kb * kKB + c * 8);
        *reinterpret_cast<uint4*>(sW + kb * NH * 128 + sw128_offset(n, c)) = v;
    }
    fence_proxy_async_smem();
    if (tid == 0) {
        for (int i = 0; i < kSlots2; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(accfull, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc_pair(tmem_slot, 128);
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr int kOps = kNKB / 2;

    if (warp == 4) {                       // ---- TMA producer (both CTAs) ----
        if (lane == 0) {
            tma_prefetch_desc(&tmapA);
            for (int op = 0; op < kOps; ++op) {
                const int slot = op % kSlots2, round = op / kSlots2;
                if (round > 0 && !wait_bar(&empty[slot], (round - 1) & 1)) { atomicExch(status, 2); break; }
                if (leader) mbar_arrive_expect_tx(&full[slot], 2 * kSlotBytes2);
                tma_load_3d_pair(sA + slot * kSlotBytes2, &tmapA, &full[slot], 0, (int)rank * kM, 2 * op);
            }
        }
    } else if (warp == 5) {                // ---- MMA issuer (even CTA) ----
        if (lane == 0 && leader) {
            const uint32_t idesc = umma_idesc_f16(2 * kM, N);
            bool ok = true;
            for (int op = 0; op < kOps && ok; ++op) {
                const int slot = op % kSlots2, round = op / kSlots2;
                if (!wait_bar(&full[slot], round & 1)) { atomicExch(status, 3); ok = false; break; }
                tcgen05_fence_after();
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const uint64_t ad = umma_desc_sw128(smem_u32(sA + slot * kSlotBytes2 + kk * kM * 128));
                    const uint64_t bd = umma_desc_sw128(smem_u32(sW + (2 * op + kk) * NH * 128));
                    if (op == 0 && kk == 0) umma_f16_pair<false>(tmem, ad, bd, idesc); else umma_f16_pair<true>(tmem, ad, bd, idesc);
                    umma_f16_pair<true>(tmem, umma_desc_advance(ad, 32), umma_desc_advance(bd, 32), idesc);
                    umma_f16_pair<true>(tmem, umma_desc_advance(ad, 64), umma_desc_advance(bd, 64), idesc);
                    umma_f16_pair<true>(tmem, umma_desc_advance(ad, 96), umma_desc_advance(bd, 96), idesc);
                }
                umma_commit_pair(&empty[slot]);
            }
            umma_commit_pair(accfull);
        }
    } else {                               // ---- epilogue warps 0..3 of both CTAs ----
        const bool ok = wait_bar(accfull, 0);
        tcgen05_fence_after();
        if (!ok) { if (lane == 0) atomicExch(status, 4); }
        else {
            const int row = warp * 32 + lane;
            for (int c0 = 0; c0 < N; c0 += 8) {
                float v[8];
                tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) C[(size_t)(rank * kM + row) * N + c0 + i] = v[i];
            }
        }
    }
    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc_pair(tmem, 128);
}
// Microbenchmark: issue rate of tcgen05.mma (M=128, K=16, SS operands, SWIZZLE_128B) from one thread.
//   mode 0: back-to-back MMAs on one accumulator, one commit at the end
//   mode 1: commit to an mbarrier after every 4 MMAs (the loop kernels' per-k-block pattern), no waiting
//   mode 2: as 1, and wait for each commit before issuing the next 4 (fully serialised: MMA latency)
//   mode 5 / 7: rotate over 4 / 2 independent accumulators (N <= 64);  mode 6: M = 64
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int N, int iters, int mode, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = base;                    // 4 x 16 KB (modes 3/4 rotate over them)
    uint8_t* sB = base + 65536;            // N x 128 B (<= 32 KB)
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + 65536 + 32768);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 2);
    for (int i = threadIdx.x; i < (65536 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); tslot[1] = 0; }
    if (threadIdx.x < 32) tmem_alloc(tslot, 256);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tslot;
    if (threadIdx.x == 0) {
        // mode 0..2: descriptors precomputed, accumulate flag constant (the tight issue loop)
        // mode 20: descriptors rebuilt and predicate set up for every MMA (what the first loop kernels did)
        const uint32_t idesc = umma_idesc_f16(128, N), a0 = smem_u32(sA), b0 = smem_u32(sB);
        uint32_t ph = 0;
        const uint64_t ad = umma_desc_sw128(a0), bd = umma_desc_sw128(b0);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (mode == 3 || mode == 4) {          // the loop kernels' per-k-block pattern: fence, 4 MMAs on a fresh tile, commit
                tcgen05_fence_after();
                const uint64_t a2 = umma_desc_sw128(a0 + (i & 3) * 16384);
                umma_f16_c<true>(tmem, a2, bd, idesc);
                umma_f16_c<true>(tmem, umma_desc_advance(a2, 32), umma_desc_advance(bd, 32), idesc);
                umma_f16_c<true>(tmem, umma_desc_advance(a2, 64), umma_desc_advance(bd, 64), idesc);
                umma_f16_c<true>(tmem, umma_desc_advance(a2, 96), umma_desc_advance(bd, 96), idesc);
                umma_commit(&bar[0]);
                continue;
            }
            if (mode == 5 || mode == 7) {          // independent accumulators: is the ~100 clk a dependent-accumulate latency?
                const uint32_t na = mode == 5 ? 4u : 2u;
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j)
                    umma_f16_c<true>(tmem + (j % na) * 64u, umma_desc_advance(ad, 32 * j), umma_desc_advance(bd, 32 * j), idesc);
                continue;
            }
            if (mode == 6) {                       // M = 64
                const uint32_t id64 = umma_idesc_f16(64, N);
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) umma_f16_c<true>(tmem, umma_desc_advance(ad, 32 * j), umma_desc_advance(bd, 32 * j), id64);
                continue;
            }
            if (mode == 20) {
#pragma unroll
                for (int j = 0; j < 4; ++j) umma_f16(tmem, umma_desc_sw128(a0 + j * 32), umma_desc_sw128(b0 + j * 32), idesc, (i | j) ? 1u : 0u);
            } else {
                umma_f16_c<true>(tmem, ad, bd, idesc);
                umma_f16_c<true>(tmem, umma_desc_advance(ad, 32), umma_desc_advance(bd, 32), idesc);
                umma_f16_c<true>(tmem, umma_desc_advance(ad, 64), umma_desc_advance(bd, 64), idesc);
                umma_f16_c<true>(tmem, umma_desc_advance(ad, 96), umma_desc_advance(bd, 96), idesc);
            }
            if (mode >= 1) umma_commit(&bar[0]);
            if (mode == 2) { while (!mbar_try_wait(&bar[0], ph)) {} ph ^= 1; }
        }
        const long long t1 = clock64();
        umma_commit(&bar[1]);
        while (!mbar_try_wait(&bar[1], 0)) {}
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
        *reinterpret_cast<volatile int*>(tslot + 1) = 1;
    } else if (mode == 4 && threadIdx.x >= 32) {
        // contention: three warps stream 16-byte loads/stores over the A tiles while the MMAs run
        uint4 acc = make_uint4(0, 0, 0, 0);
        int it = 0;
        while (*reinterpret_cast<volatile int*>(tslot + 1) == 0 && it < 2000000) {
            const uint4 v = reinterpret_cast<uint4*>(sA)[(threadIdx.x * 7 + it * 96) & 4095];
            acc.x ^= v.x;
            reinterpret_cast<uint4*>(sB + 16384)[(threadIdx.x + it) & 1023] = acc;
            ++it;
        }
        if (acc.x == 12345u) out[3] = it;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}
}  // namespace

cudaError_t run_umma_rate(int N, int iters, int mode, long long* out_dev, cudaStream_t stream) {
    const int smem = 1024 + 65536 + 32768 + 64;
    cudaError_t e = cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    umma_rate_kernel<<<1, 128, smem, stream>>>(N, iters, mode, out_dev);
    return cudaGetLastError();
}

cudaError_t run_tc_gemm_test(const void* A_dev, const void* W_dev, int N, float* C_dev, int* status_dev, cudaStream_t stream) {
    if (N % 16 != 0 || N < 16 || N > 64) return cudaErrorInvalidValue;
    alignas(64) CUtensorMap tmap;
    cudaError_t e = make_tmap_f16_2d(&tmap, A_dev, kM, kK, kM, kKB);
    if (e != cudaSuccess) return e;
    const size_t smem = 1024 + (size_t)kSlots * kM * 128 + (size_t)kNKB * N * 128 + 256;
    e = cudaFuncSetAttribute(tc_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    tc_gemm_test_kernel<<<1, 192, smem, stream>>>(tmap, reinterpret_cast<const __half*>(W_dev), N, C_dev, status_dev);
    return cudaGetLastError();
}

cudaError_t run_tc_gemm2_test(const void* A_dev, const void* W_dev, int N, float* C_dev, int* status_dev, cudaStream_t stream) {
    if (N % 32 != 0 || N < 32 || N > 128) return cudaErrorInvalidValue;
    alignas(64) CUtensorMap tmap;
    cudaError_t e = make_tmap_f16_kblocks(&tmap, A_dev, 2 * kM, kK, kM, 2);
    if (e != cudaSuccess) return e;
    const size_t smem = 1024 + (size_t)kSlots2 * kSlotBytes2 + (size_t)kNKB * (N / 2) * 128 + 256;
    e = cudaFuncSetAttribute(tc_gemm2_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, tc_gemm2_test_kernel, tmap, reinterpret_cast<const __half*>(W_dev), N, C_dev, status_dev);
}

}  // namespace wrnn
