// xchg6.cu -- hypothesis: a line that many SMs (on both dies) poll is slow to show a new store; a line with ONE writer and
// ONE polling SM is fast (pairwise one-way 0.4-0.7 us).  Protocol: data chunks as before (sentinel-validated, read ONCE),
// plus a private hint flag per (producer, consumer) pair in its own 128-byte line: the producer stores data, then (no
// fence) the 16 hint flags; warp 0 of the consumer polls its 16 flags (lane = producer), then everybody loads the data in
// one pass and re-polls only chunks that still hold the sentinel.
//   MODE 0: as described.   MODE 1: the producer waits DELAY clocks between data and flags.
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
__device__ __forceinline__ void st_u32(unsigned* p, unsigned v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ bool chunk_ready(uint4 v) {
    return (__vcmpeq2(v.x, 0xFFFFFFFFu) | __vcmpeq2(v.y, 0xFFFFFFFFu) | __vcmpeq2(v.z, 0xFFFFFFFFu) | __vcmpeq2(v.w, 0xFFFFFFFFu)) == 0u;
}
constexpr int kChunks = 64, kRows = 128, P = 16;
// flags: [2 sides][P producers][P consumers][32 words] (one 128-byte line per pair), value = round + 1
template <int DELAY>
__global__ void __launch_bounds__(512, 1) k(uint4* X, uint4* Y, unsigned* flags, int NF, int rounds, long long* clk, int* errors, long long limit) {
    const int side = blockIdx.x / P, cta = blockIdx.x % P, tid = threadIdx.x;
    uint4* out = side == 0 ? X : Y;
    const uint4* in = side == 0 ? Y : X;
    unsigned* fout = flags + (size_t)side * P * P * 32;             // [me as producer][consumer]
    const unsigned* fin = flags + (size_t)(1 - side) * P * P * 32;  // [producer][me as consumer]
    const int row = tid & 127, cg = tid >> 7;
    const bool live = row < NF;
    const size_t bufsz = (size_t)kChunks * kRows;
    uint32_t acc = 1;
    int bad = 0, repolls = 0;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        if (!(side == 0 && r == 0)) {
            const int rr = side == 0 ? r - 1 : r;
            if (tid < P) {
                int spins = 0;
                while (ld_u32(fin + ((size_t)tid * P + cta) * 32) != (unsigned)rr + 1u)
                    if (((++spins) & 4095) == 0 && clock64() - t0 > limit) { bad += 1000; break; }
            }
            __syncthreads();
            const uint4* base = in + (size_t)(rr % 3) * bufsz + (size_t)(cg * 16) * kRows + row;
            if (live) {
                uint4 v[16];
                uint32_t pending = 0xFFFFu;
                int spins = 0;
                while (pending) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if ((pending >> i) & 1u) v[i] = ld_v4(base + (size_t)i * kRows);
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (((pending >> i) & 1u) && chunk_ready(v[i])) pending &= ~(1u << i);
                    if (pending) ++repolls;
                    if (pending && ((++spins) & 4095) == 0 && clock64() - t0 > limit) { bad += 1000; break; }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (v[i].x != (uint32_t)rr) ++bad;
                    acc += v[i].y;
                }
            }
            if (__syncthreads_or(bad >= 1000)) break;
        }
        if (live) {
            const size_t o = (size_t)(4 * cta + cg) * kRows + row;
            st_v4(out + (size_t)(r % 3) * bufsz + o, make_uint4((uint32_t)r, acc & 0x7FFF7FFFu, 0x3C003C00u, 0x3C003C00u));
            st_v4(out + (size_t)((r + 1) % 3) * bufsz + o, make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu));
        }
        __syncthreads();                 // every thread has ISSUED its data stores
        if (DELAY > 0 && tid < P) { const long long td = clock64(); while (clock64() - td < DELAY) {} }
        if (tid < P) st_u32(fout + ((size_t)cta * P + tid) * 32, (unsigned)r + 1u);
    }
    const long long t1 = clock64();
    if (tid == 0) clk[blockIdx.x] = t1 - t0;
    if (bad) atomicAdd(errors, bad);
    if (repolls) atomicAdd(errors + 1, repolls);
}
template <int DELAY>
static void run(uint4* X, uint4* Y, unsigned* flags, long long* dclk, int* derr, int clk_khz) {
    const size_t bytes = 3ull * kChunks * kRows * 16;
    for (int NF : {16, 32, 54, 72, 107, 128}) {
        cudaMemset(X, 0xFF, bytes); cudaMemset(Y, 0xFF, bytes); cudaMemset(derr, 0, 8); cudaMemset(flags, 0, 2 * P * P * 128);
        int rounds = 4000;
        long long limit = 2000000000LL;
        void* args[] = {&X, &Y, &flags, (void*)&NF, (void*)&rounds, &dclk, &derr, &limit};
        cudaError_t e = cudaLaunchCooperativeKernel((const void*)k<DELAY>, dim3(2 * P), dim3(512), args, 0, 0);
        cudaError_t e2 = cudaDeviceSynchronize();
        long long clk[64]; int err[2];
        cudaMemcpy(clk, dclk, 8 * 2 * P, cudaMemcpyDeviceToHost);
        cudaMemcpy(err, derr, 8, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < 2 * P; ++i) mx = clk[i] > mx ? clk[i] : mx;
        printf("hint flags, delay %4d, folds=%3d (%3d KB): %s %s  %.0f clk = %.3f us per exchange  errors %d, straggler re-polls per thread-round %.3f\n", DELAY, NF,
               NF * kChunks * 16 / 1024, cudaGetErrorString(e), cudaGetErrorString(e2), (double)mx / (2.0 * rounds), (double)mx / (2.0 * rounds) / (clk_khz * 1e-3),
               err[0], (double)err[1] / (2.0 * P * 4 * NF * rounds));
        fflush(stdout);
    }
}
int main() {
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const size_t bytes = 3ull * kChunks * kRows * 16;
    uint4 *X, *Y; long long* dclk; int* derr; unsigned* flags;
    cudaMalloc(&X, bytes); cudaMalloc(&Y, bytes); cudaMalloc(&dclk, 8 * 64); cudaMalloc(&derr, 8); cudaMalloc(&flags, 2 * P * P * 128);
    run<0>(X, Y, flags, dclk, derr, clk_khz);
    run<200>(X, Y, flags, dclk, derr, clk_khz);
    run<500>(X, Y, flags, dclk, derr, clk_khz);
    return 0;
}
