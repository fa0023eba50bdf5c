// ts_probe.cu -- two hardware probes behind the round-2 loop design:
//  (1) tcgen05.mma with the A operand in TMEM (written there by tcgen05.st): layout check against a CPU product and the
//      issue rate per K=16 step for several N, next to the SS form (A from shared memory);
//  (2) the sentinel exchange: two sides of P CTAs ping-pong a [64 chunks][folds] matrix of 16-byte chunks through L2 with
//      no fence, flag or atomic -- a chunk is valid when none of its fp16 halves is 0xFFFF; triple-buffered, the writer
//      resets the buffer after next.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/probes/bin/ts_probe tools/probes/ts_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../real-time-voice-cloning_b200/csrc/tc_common.cuh"

using namespace wrnn::tc;

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
template <bool kAcc>
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc) {
    if (kAcc)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
                     "r"(tmem_a), "l"(bdesc), "r"(idesc)
                     : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
                     "r"(tmem_a), "l"(bdesc), "r"(idesc)
                     : "memory");
}

// A [128][K] fp16 row-major; Wsw = B operand, pre-swizzled [K/64][N][64] (SWIZZLE_128B K-major); D [128][N] fp32
__global__ void __launch_bounds__(160, 1) ts_mma_kernel(const __half* A, const __half* Wsw, int N, int nkb, int iters, int mode, float* D,
                                                        long long* clk) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sB = smem;                                  // nkb x N x 128 B
    uint8_t* sA = smem + (size_t)nkb * N * 128;          // SS mode: one 16 KB k-block of A (timing only, reused for every k-block)
    sA = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sA) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < nkb * N * 8; i += blockDim.x) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(Wsw)[i];
    for (int i = tid; i < 1024; i += blockDim.x) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tslot, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tslot;
    const int K = nkb * 64;
    if (warp < 4) {       // A -> TMEM: lane = row, 32-bit column c holds k = 2c, 2c+1
        const int m = warp * 32 + lane;
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < K / 2; c += 8) {
            uint32_t r[8];
            const uint4 a = *reinterpret_cast<const uint4*>(A + (size_t)m * K + 2 * c);
            const uint4 b = *reinterpret_cast<const uint4*>(A + (size_t)m * K + 2 * c + 8);
            r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
            tmem_st8(tl + c, r);
        }
        tmem_st_wait();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t dcol = tmem + 256;
    if (warp == 4 && lane == 0) {
        const uint32_t idesc = umma_idesc_f16(128, N);
        uint32_t ph = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            for (int k = 0; k < K / 16; ++k) {
                const uint64_t bd = umma_desc_advance(umma_desc_sw128(smem_u32(sB) + (k >> 2) * N * 128), (k & 3) * 32);
                if (mode == 0) {
                    if (k == 0) umma_ts<false>(dcol, tmem + k * 8, bd, idesc); else umma_ts<true>(dcol, tmem + k * 8, bd, idesc);
                } else {
                    const uint64_t ad = umma_desc_advance(umma_desc_sw128(smem_u32(sA)), (k & 3) * 32);
                    if (k == 0) umma_f16_c<false>(dcol, ad, bd, idesc); else umma_f16_c<true>(dcol, ad, bd, idesc);
                }
            }
            if (mode >= 10) {   // per-job commit + wait (the latency of one K-deep job, not the rate)
                umma_commit(&bar);
                while (!mbar_try_wait(&bar, ph)) {}
                ph ^= 1;
            }
        }
        if (mode < 10) { umma_commit(&bar); while (!mbar_try_wait(&bar, 0)) {} }
        const long long t1 = clock64();
        clk[0] = t1 - t0;
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp < 4) {
        const int m = warp * 32 + lane;
        const uint32_t tl = dcol + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < N; c += 8) {
            float v[8];
            tmem_ld8(tl + c, v);
            tmem_ld_wait();
            for (int i = 0; i < 8; ++i) D[(size_t)m * N + c + i] = v[i];
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_v4(uint4* p, uint4 v) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ bool chunk_ready(uint4 v) {       // no half equals 0xFFFF
    return (__vcmpeq2(v.x, 0xFFFFFFFFu) | __vcmpeq2(v.y, 0xFFFFFFFFu) | __vcmpeq2(v.z, 0xFFFFFFFFu) | __vcmpeq2(v.w, 0xFFFFFFFFu)) == 0u;
}

constexpr int kChunks = 64, kRows = 128;
// X, Y: [3][kChunks][kRows] uint4, all 0xFF at launch.  side 0 CTAs write X and read Y, side 1 the reverse.
__global__ void __launch_bounds__(512, 1) xchg_kernel(uint4* X, uint4* Y, int P, int NF, int rounds, long long* clk, int* errors,
                                                       long long limit) {
    const int side = blockIdx.x / P, cta = blockIdx.x % P, tid = threadIdx.x;
    uint4* out = side == 0 ? X : Y;
    const uint4* in = side == 0 ? Y : X;
    const int cpc = kChunks / P;                    // chunks this CTA writes
    const int row = tid & 127, cg = tid >> 7;       // ingest: thread = (row, 16 chunks)
    const size_t bufsz = (size_t)kChunks * kRows;
    uint32_t acc = 0;
    int bad = 0;
    __shared__ int s_abort;
    if (tid == 0) s_abort = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        // ---- ingest the other side's matrix of this round (side 0 starts round 0 without one)
        if (!(side == 0 && r == 0)) {
            const int rr = side == 0 ? r - 1 : r;
            const uint4* buf = in + (size_t)(rr % 3) * bufsz;
            if (row < NF) {
                uint32_t pending = 0xFFFFu;
                uint4 v[16];
                int spins = 0;
                while (pending) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if ((pending >> i) & 1u) v[i] = ld_v4(buf + (size_t)(cg * 16 + i) * kRows + row);
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (((pending >> i) & 1u) && chunk_ready(v[i])) {
                            pending &= ~(1u << i);
                            const uint32_t want = ((uint32_t)rr << 16) | ((uint32_t)(cg * 16 + i) << 8) | (uint32_t)row;
                            if (v[i].x != (want & 0x7FFF7FFFu)) ++bad;
                            acc += v[i].y;
                        }
                    if (pending && ((++spins) & 1023) == 0 && (s_abort || clock64() - t0 > limit)) { s_abort = 1; break; }
                }
            }
            __syncthreads();
        }
        // ---- publish my slice of this round, reset my slice of the buffer after next
        uint4* wb = out + (size_t)(r % 3) * bufsz;
        uint4* rb = out + (size_t)((r + 1) % 3) * bufsz;
        for (int i = tid; i < cpc * NF; i += 512) {
            const int c = cta * cpc + i / NF, f = i % NF;
            const uint32_t tagw = (((uint32_t)r << 16) | ((uint32_t)c << 8) | (uint32_t)f) & 0x7FFF7FFFu;
            st_v4(rb + (size_t)c * kRows + f, make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu));
            st_v4(wb + (size_t)c * kRows + f, make_uint4(tagw, acc & 0x7FFF7FFFu, 0x3C003C00u, 0x3C003C00u));
        }
    }
    const long long t1 = clock64();
    if (tid == 0) clk[blockIdx.x] = t1 - t0;
    if (bad) atomicAdd(errors, bad);
    if (s_abort && tid == 0) atomicAdd(errors + 1, 1);
}

static uint16_t f2h(float f) { __half h = __float2half_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }
static float h2f(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }

int main(int argc, char** argv) {
    int dev = 0; cudaSetDevice(dev);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, clk_khz);
    // ---- (1) TS-mode MMA
    {
        const int nkbs[2] = {8, 4};
        const int Ns[] = {16, 32, 48, 64, 96, 112, 128, 192, 256};
        for (int N : Ns) {
            const int nkb = N > 128 ? nkbs[1] : nkbs[0], K = nkb * 64;
            std::vector<uint16_t> A(128 * K), W((size_t)N * K), Wsw((size_t)N * K);
            srand(7 + N);
            for (auto& x : A) x = f2h((rand() % 2001 - 1000) / 1000.0f);
            for (auto& x : W) x = f2h((rand() % 2001 - 1000) / 1000.0f);
            for (int kb = 0; kb < nkb; ++kb)
                for (int n = 0; n < N; ++n)
                    for (int c = 0; c < 8; ++c)
                        for (int e = 0; e < 8; ++e)
                            Wsw[((size_t)kb * N + n) * 64 + ((c ^ (n & 7)) * 8) + e] = W[(size_t)n * K + kb * 64 + c * 8 + e];
            __half *dA, *dW; float* dD; long long* dclk;
            cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dW, Wsw.size() * 2); cudaMalloc(&dD, 128 * N * 4); cudaMalloc(&dclk, 8);
            cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
            cudaMemcpy(dW, Wsw.data(), Wsw.size() * 2, cudaMemcpyHostToDevice);
            const int smem = 1024 + nkb * N * 128 + 1024 + 16384;
            cudaFuncSetAttribute(ts_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            for (int mode : {0, 1, 10, 11}) {
                const int iters = 200;
                cudaMemset(dD, 0, 128 * N * 4);
                ts_mma_kernel<<<1, 160, smem>>>(dA, dW, N, nkb, iters, mode, dD, dclk);
                cudaError_t e = cudaDeviceSynchronize();
                long long clk = 0; cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost);
                std::vector<float> D(128 * N);
                cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
                double maxerr = 0, maxref = 0;
                if (mode == 0 || mode == 10) {
                    for (int m = 0; m < 128; ++m)
                        for (int n = 0; n < N; ++n) {
                            double s = 0;
                            for (int k = 0; k < K; ++k) s += (double)h2f(A[(size_t)m * K + k]) * h2f(W[(size_t)n * K + k]);
                            maxerr = fmax(maxerr, fabs(s - D[(size_t)m * N + n])); maxref = fmax(maxref, fabs(s));
                        }
                }
                printf("mma N=%3d K=%d mode=%2d (%s%s): %s  %.1f clk per K=16 step, %.0f clk per K-deep job  max err %.3g (ref %.3g)\n", N, K, mode,
                       (mode % 10) == 0 ? "TS" : "SS", mode >= 10 ? ", commit+wait per job" : "", cudaGetErrorString(e),
                       (double)clk / (iters * (K / 16)), (double)clk / iters, maxerr, maxref);
            }
            cudaFree(dA); cudaFree(dW); cudaFree(dD); cudaFree(dclk);
        }
    }
    // ---- (2) sentinel exchange
    {
        const size_t bytes = 3ull * kChunks * kRows * 16;
        uint4 *X, *Y; long long* dclk; int* derr;
        cudaMalloc(&X, bytes); cudaMalloc(&Y, bytes); cudaMalloc(&dclk, 8 * 256); cudaMalloc(&derr, 8);
        for (int P : {1, 4, 16, 32})
            for (int NF : {32, 64, 96, 128}) {
                cudaMemset(X, 0xFF, bytes); cudaMemset(Y, 0xFF, bytes); cudaMemset(derr, 0, 8);
                const int rounds = 2000;
                long long limit = 4000000000LL;
                void* args[] = {&X, &Y, (void*)&P, (void*)&NF, (void*)&rounds, &dclk, &derr, &limit};
                cudaError_t e = cudaLaunchCooperativeKernel((const void*)xchg_kernel, dim3(2 * P), dim3(512), args, 0, 0);
                cudaError_t e2 = cudaDeviceSynchronize();
                long long clk[256]; int err[2];
                cudaMemcpy(clk, dclk, 8 * 2 * P, cudaMemcpyDeviceToHost);
                cudaMemcpy(err, derr, 8, cudaMemcpyDeviceToHost);
                long long mx = 0;
                for (int i = 0; i < 2 * P; ++i) mx = clk[i] > mx ? clk[i] : mx;
                printf("xchg P=%2d folds=%3d (%3d KB per matrix): %s %s  %.0f clk = %.3f us per exchange  errors %d aborts %d\n", P, NF, NF * kChunks * 16 / 1024,
                       cudaGetErrorString(e), cudaGetErrorString(e2), (double)mx / (2.0 * rounds), (double)mx / (2.0 * rounds) / (clk_khz * 1e-3), err[0], err[1]);
            }
    }
    return 0;
}
