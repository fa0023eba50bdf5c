// xchg5.cu -- the exchange as the loop kernel would do it, no integer divisions anywhere: P CTAs per side, 512 threads,
// thread = (fold row = tid & 127, column group cg = tid >> 7).  Publish: CTA c owns chunks [4c, 4c+4) (P = 16: 8 hidden units
// per thread -> one 16-byte chunk at [4c + cg][row]).  Ingest: thread polls the 16 chunks [16 cg, 16 cg + 16) of its row.
// Sentinel validity, triple buffering, relaxed accesses.  MODE 0: re-poll everything pending; MODE 1: first chunk as canary.
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
__device__ __forceinline__ bool chunk_ready_s(uint4 v) {
    return (__vcmpeq2(v.x, 0xFFFFFFFFu) | __vcmpeq2(v.y, 0xFFFFFFFFu) | __vcmpeq2(v.z, 0xFFFFFFFFu) | __vcmpeq2(v.w, 0xFFFFFFFFu)) == 0u;
}
constexpr int kChunks = 64, kRows = 128, P = 16;
template <int MODE, int NTHR, int RESET, int NACT>
__global__ void __launch_bounds__(NTHR, 1) k(uint4* X, uint4* Y, int NF, int rounds, long long* clk, int* errors, long long limit) {
    const int side = blockIdx.x / P, cta = blockIdx.x % P, tid = threadIdx.x;
    uint4* out = side == 0 ? X : Y;
    const uint4* in = side == 0 ? Y : X;
    constexpr int NCG = NTHR / 128, CPT = kChunks / NCG;     // column groups, chunks per ingest thread
    const int row = tid & 127, cg = tid >> 7;
    const bool live = row < NF;
    const size_t bufsz = (size_t)kChunks * kRows;
    uint32_t acc = 1;
    int bad = 0;
    long long tpoll = 0;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        if (!(side == 0 && r == 0)) {
            const int rr = side == 0 ? r - 1 : r;
            const uint4* base = in + (size_t)(rr % 3) * bufsz + (size_t)(cg * CPT) * kRows + row;
            const long long ta = clock64();
            if (live) {
                uint4 v[CPT];
                uint32_t pending = CPT == 32 ? 0xFFFFFFFFu : ((1u << CPT) - 1u);
                if (cta >= NACT) pending = (cg == 0) ? 1u : 0u;
                int spins = 0;
                if (MODE == 1) {
                    while (true) {
                        v[0] = ld_v4(base);
                        if (RESET ? chunk_ready_s(v[0]) : (v[0].x == (uint32_t)rr)) break;
                        if (((++spins) & 4095) == 0 && clock64() - t0 > limit) { bad += 1000; break; }
                    }
                    pending &= ~1u;
                }
                while (pending) {
#pragma unroll
                    for (int i = 0; i < CPT; ++i)
                        if ((pending >> i) & 1u) v[i] = ld_v4(base + (size_t)i * kRows);
#pragma unroll
                    for (int i = 0; i < CPT; ++i)
                        if (((pending >> i) & 1u) && (RESET ? chunk_ready_s(v[i]) : (v[i].x == (uint32_t)rr))) pending &= ~(1u << i);
                    if (pending && ((++spins) & 4095) == 0 && clock64() - t0 > limit) { bad += 1000; break; }
                }
                if (cta < NACT) {
#pragma unroll
                for (int i = 0; i < CPT; ++i) {
                    if (v[i].x != (uint32_t)rr) ++bad;
                    acc += v[i].y;
                }
                }
            }
            tpoll += clock64() - ta;
            if (bad >= 1000) break;
            __syncthreads();
        }
        // publish: chunk [4 cta + (cg & 3)][row]; with NTHR = 1024 the upper half of the threads only ingests
        if (live && cg < 4) {
            const size_t o = (size_t)(4 * cta + cg) * kRows + row;
            st_v4(out + (size_t)(r % 3) * bufsz + o, make_uint4((uint32_t)r, acc & 0x7FFF7FFFu, 0x3C003C00u, 0x3C003C00u));
            if (RESET) st_v4(out + (size_t)((r + 1) % 3) * bufsz + o, make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu));
        }
    }
    const long long t1 = clock64();
    if (tid == 0) { clk[2 * blockIdx.x] = t1 - t0; clk[2 * blockIdx.x + 1] = tpoll; }
    if (bad) atomicAdd(errors, bad);
}
template <int MODE, int NTHR, int RESET, int NACT>
static void run(uint4* X, uint4* Y, long long* dclk, int* derr, int clk_khz) {
    const size_t bytes = 3ull * kChunks * kRows * 16;
    for (int NF : {16, 72}) {
        cudaMemset(X, 0xFF, bytes); cudaMemset(Y, 0xFF, bytes); cudaMemset(derr, 0, 8);
        int rounds = 4000;
        long long limit = 4000000000LL;
        void* args[] = {&X, &Y, (void*)&NF, (void*)&rounds, &dclk, &derr, &limit};
        cudaError_t e = cudaLaunchCooperativeKernel((const void*)k<MODE, NTHR, RESET, NACT>, dim3(2 * P), dim3(NTHR), args, 0, 0);
        cudaError_t e2 = cudaDeviceSynchronize();
        long long clk[128]; int err[2];
        cudaMemcpy(clk, dclk, 8 * 4 * P, cudaMemcpyDeviceToHost);
        cudaMemcpy(err, derr, 8, cudaMemcpyDeviceToHost);
        long long mx = 0, pl = 0;
        for (int i = 0; i < 2 * P; ++i) { mx = clk[2 * i] > mx ? clk[2 * i] : mx; pl += clk[2 * i + 1]; }
        printf("reset %d active consumers %2d mode %d threads %4d folds=%3d (%3d KB): %s %s  %.0f clk = %.3f us per exchange (thread 0 in ingest: %.0f clk)  errors %d\n", RESET, NACT, MODE, NTHR, NF,
               NF * kChunks * 16 / 1024, cudaGetErrorString(e), cudaGetErrorString(e2), (double)mx / (2.0 * rounds), (double)mx / (2.0 * rounds) / (clk_khz * 1e-3),
               (double)pl / (2.0 * P) / rounds, err[0]);
        fflush(stdout);
    }
}
int main() {
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const size_t bytes = 3ull * kChunks * kRows * 16;
    uint4 *X, *Y; long long* dclk; int* derr;
    cudaMalloc(&X, bytes); cudaMalloc(&Y, bytes); cudaMalloc(&dclk, 8 * 128); cudaMalloc(&derr, 8);
    run<0, 512, 1, 16>(X, Y, dclk, derr, clk_khz);
    run<0, 512, 0, 16>(X, Y, dclk, derr, clk_khz);
    run<0, 512, 1, 4>(X, Y, dclk, derr, clk_khz);
    run<0, 512, 1, 1>(X, Y, dclk, derr, clk_khz);
    run<0, 512, 0, 1>(X, Y, dclk, derr, clk_khz);
    return 0;
}
