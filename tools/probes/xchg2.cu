// xchg2.cu -- how fast can P producer CTAs hand a [64 chunks][folds] fp16 matrix (16-byte chunks) to P consumer CTAs
// through L2 without fences?  Two sides ping-pong; time per one-way exchange for several ingest strategies:
//   mode 0: every thread re-polls all of its 16 chunks until none holds the 0xFFFF sentinel
//   mode 1: thread polls its FIRST chunk only; when it is valid it loads the other 15 once and re-polls only stragglers
//   mode 2: producers write an unordered flag word after their data; one warp polls the P flags, then everybody
//           bulk-loads and validates by sentinel (stragglers re-polled)
//   mode 3: as 1, but the consumer side uses 256 threads with 32 chunks each (fewer, longer request streams)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/probes/bin/xchg2 tools/probes/xchg2.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ld_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_v4(uint4* p, uint4 v) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned ld_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool chunk_ready(uint4 v) {
    return (__vcmpeq2(v.x, 0xFFFFFFFFu) | __vcmpeq2(v.y, 0xFFFFFFFFu) | __vcmpeq2(v.z, 0xFFFFFFFFu) | __vcmpeq2(v.w, 0xFFFFFFFFu)) == 0u;
}

constexpr int kChunks = 64, kRows = 128;

template <int MODE>
__global__ void __launch_bounds__(512, 1) xchg_kernel(uint4* X, uint4* Y, unsigned* flags, int P, int NF, int rounds, long long* clk, int* errors,
                                                       long long limit) {
    const int side = blockIdx.x / P, cta = blockIdx.x % P, tid = threadIdx.x;
    uint4* out = side == 0 ? X : Y;
    const uint4* in = side == 0 ? Y : X;
    unsigned* fout = flags + side * 3 * 64;
    const unsigned* fin = flags + (1 - side) * 3 * 64;
    const int cpc = kChunks / P;
    constexpr int CPT = MODE == 3 ? 32 : 16;                 // chunks per thread
    const int row = tid & 127, cg = tid >> 7;
    const bool ingest = row < NF && cg < kChunks / CPT;
    const size_t bufsz = (size_t)kChunks * kRows;
    uint32_t acc = 0;
    int bad = 0;
    __shared__ int s_abort;
    if (tid == 0) s_abort = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        if (!(side == 0 && r == 0)) {
            const int rr = side == 0 ? r - 1 : r;
            const uint4* buf = in + (size_t)(rr % 3) * bufsz;
            if (MODE == 2) {
                if (tid < P) {
                    int spins = 0;
                    while (ld_u32(fin + (rr % 3) * 64 + tid) != (unsigned)rr + 1u)
                        if (((++spins) & 4095) == 0 && (s_abort || clock64() - t0 > limit)) { s_abort = 1; break; }
                }
                __syncthreads();
            }
            if (ingest) {
                const uint4* base = buf + (size_t)(cg * CPT) * kRows + row;
                uint4 v[CPT];
                int spins = 0;
                if (MODE == 1 || MODE == 3) {
                    while (true) {
                        v[0] = ld_v4(base);
                        if (chunk_ready(v[0])) break;
                        if (((++spins) & 1023) == 0 && (s_abort || clock64() - t0 > limit)) { s_abort = 1; break; }
                    }
                }
                uint32_t pending = (MODE == 1 || MODE == 3) ? (CPT == 32 ? 0xFFFFFFFEu : 0xFFFEu) : (CPT == 32 ? 0xFFFFFFFFu : 0xFFFFu);
                while (pending) {
#pragma unroll
                    for (int i = 0; i < CPT; ++i)
                        if ((pending >> i) & 1u) v[i] = ld_v4(base + (size_t)i * kRows);
#pragma unroll
                    for (int i = 0; i < CPT; ++i)
                        if (((pending >> i) & 1u) && chunk_ready(v[i])) pending &= ~(1u << i);
                    if (pending && ((++spins) & 1023) == 0 && (s_abort || clock64() - t0 > limit)) { s_abort = 1; break; }
                }
#pragma unroll
                for (int i = 0; i < CPT; ++i) {
                    const uint32_t want = (((uint32_t)rr << 16) | ((uint32_t)(cg * CPT + i) << 8) | (uint32_t)row) & 0x7FFF7FFFu;
                    if (v[i].x != want) ++bad;
                    acc += v[i].y;
                }
            }
            __syncthreads();
        }
        uint4* wb = out + (size_t)(r % 3) * bufsz;
        uint4* rb = out + (size_t)((r + 1) % 3) * bufsz;
        for (int i = tid; i < cpc * NF; i += 512) {
            const int c = cta * cpc + i / NF, f = i % NF;
            const uint32_t tagw = (((uint32_t)r << 16) | ((uint32_t)c << 8) | (uint32_t)f) & 0x7FFF7FFFu;
            st_v4(wb + (size_t)c * kRows + f, make_uint4(tagw, acc & 0x7FFF7FFFu, 0x3C003C00u, 0x3C003C00u));
        }
        for (int i = tid; i < cpc * NF; i += 512) {
            const int c = cta * cpc + i / NF, f = i % NF;
            st_v4(rb + (size_t)c * kRows + f, make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu));
        }
        if (MODE == 2) {
            __syncthreads();
            if (tid == 0) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(fout + (r % 3) * 64 + cta), "r"((unsigned)r + 1u) : "memory");
        }
    }
    const long long t1 = clock64();
    if (tid == 0) clk[blockIdx.x] = t1 - t0;
    if (bad) atomicAdd(errors, bad);
    if (s_abort && tid == 0) atomicAdd(errors + 1, 1);
}

template <int MODE>
static void run(uint4* X, uint4* Y, unsigned* flags, long long* dclk, int* derr, int clk_khz) {
    const size_t bytes = 3ull * kChunks * kRows * 16;
    for (int P : {8, 16, 32})
        for (int NF : {32, 64, 72, 107, 128}) {
            cudaMemset(X, 0xFF, bytes); cudaMemset(Y, 0xFF, bytes); cudaMemset(derr, 0, 8); cudaMemset(flags, 0, 4 * 2 * 3 * 64);
            int rounds = 2000;
            long long limit = 4000000000LL;
            void* args[] = {&X, &Y, &flags, (void*)&P, (void*)&NF, (void*)&rounds, &dclk, &derr, &limit};
            cudaError_t e = cudaLaunchCooperativeKernel((const void*)xchg_kernel<MODE>, dim3(2 * P), dim3(512), args, 0, 0);
            cudaError_t e2 = cudaDeviceSynchronize();
            long long clk[256]; int err[2];
            cudaMemcpy(clk, dclk, 8 * 2 * P, cudaMemcpyDeviceToHost);
            cudaMemcpy(err, derr, 8, cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < 2 * P; ++i) mx = clk[i] > mx ? clk[i] : mx;
            printf("mode %d P=%2d folds=%3d (%3d KB): %s %s  %.0f clk = %.3f us per exchange  errors %d aborts %d\n", MODE, P, NF, NF * kChunks * 16 / 1024,
                   cudaGetErrorString(e), cudaGetErrorString(e2), (double)mx / (2.0 * rounds), (double)mx / (2.0 * rounds) / (clk_khz * 1e-3), err[0], err[1]);
        }
}

int main() {
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const size_t bytes = 3ull * kChunks * kRows * 16;
    uint4 *X, *Y; long long* dclk; int* derr; unsigned* flags;
    cudaMalloc(&X, bytes); cudaMalloc(&Y, bytes); cudaMalloc(&dclk, 8 * 256); cudaMalloc(&derr, 8); cudaMalloc(&flags, 4 * 2 * 3 * 64);
    run<0>(X, Y, flags, dclk, derr, clk_khz);
    run<1>(X, Y, flags, dclk, derr, clk_khz);
    run<2>(X, Y, flags, dclk, derr, clk_khz);
    run<3>(X, Y, flags, dclk, derr, clk_khz);
    return 0;
}
