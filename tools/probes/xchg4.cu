// xchg3.cu -- anatomy of the L2 exchange latency: ping-pong between two sides of P CTAs; T threads per CTA take part;
// each thread owns CPT 16-byte chunks it writes (own slot) and CPT chunks it reads from the other side (slot of the
// same thread index of CTA (cta + k) % P for chunk k -- so data crosses CTAs all-to-all when P > 1).
// Validity = sentinel (no 0xFFFF half), triple-buffered, no fences.  Reports clocks per one-way exchange.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
template <int K>
__device__ __forceinline__ uint4 ld_v4(const uint4* p) {
    uint4 v;
    if (K == 0) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    if (K == 1) asm volatile("ld.global.cv.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    if (K == 2) asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    if (K == 3) asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
template <int K>
__device__ __forceinline__ void st_v4(uint4* p, uint4 v) {
    if (K == 0) asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    if (K == 1) asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    if (K == 2) asm volatile("st.global.cg.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    if (K == 3) asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ bool chunk_ready(uint4 v) {
    return (__vcmpeq2(v.x, 0xFFFFFFFFu) | __vcmpeq2(v.y, 0xFFFFFFFFu) | __vcmpeq2(v.z, 0xFFFFFFFFu) | __vcmpeq2(v.w, 0xFFFFFFFFu)) == 0u;
}
constexpr int kMaxP = 32, kMaxT = 512, kMaxC = 16;
// buffer: [3][P][CPT][T] uint4
template <int CPT, bool SYNC, int LK, int SK>
__global__ void __launch_bounds__(512, 1) k(uint4* X, uint4* Y, int P, int T, int rounds, long long* clk, int* errors, long long limit) {
    const int side = blockIdx.x / P, cta = blockIdx.x % P, tid = threadIdx.x;
    uint4* out = side == 0 ? X : Y;
    const uint4* in = side == 0 ? Y : X;
    const size_t bufsz = (size_t)P * CPT * T;
    uint32_t acc = 1;
    int bad = 0;
    const long long t0 = clock64();
    if (tid < T) {
        for (int r = 0; r < rounds; ++r) {
            if (!(side == 0 && r == 0)) {
                const int rr = side == 0 ? r - 1 : r;
                const uint4* buf = in + (size_t)(rr % 3) * bufsz;
                uint4 v[CPT];
                uint32_t pending = (1u << CPT) - 1u;
                int spins = 0;
                while (pending) {
#pragma unroll
                    for (int i = 0; i < CPT; ++i)
                        if ((pending >> i) & 1u) v[i] = ld_v4<LK>(buf + ((size_t)((cta + i) % P) * CPT + i) * T + tid);
#pragma unroll
                    for (int i = 0; i < CPT; ++i)
                        if (((pending >> i) & 1u) && chunk_ready(v[i])) pending &= ~(1u << i);
                    if (pending && ((++spins) & 4095) == 0 && clock64() - t0 > limit) { bad += 1000; break; }
                }
#pragma unroll
                for (int i = 0; i < CPT; ++i) {
                    if (v[i].x != (uint32_t)rr) ++bad;
                    acc += v[i].y;
                }
                if (SYNC) asm volatile("bar.sync 1, %0;" ::"r"(T) : "memory");
            }
            uint4* wb = out + (size_t)(r % 3) * bufsz + (size_t)cta * CPT * T + tid;
            uint4* rb = out + (size_t)((r + 1) % 3) * bufsz + (size_t)cta * CPT * T + tid;
#pragma unroll
            for (int i = 0; i < CPT; ++i) st_v4<SK>(wb + (size_t)i * T, make_uint4((uint32_t)r, acc & 0x7FFF7FFFu, 0x3C003C00u, 0x3C003C00u));
#pragma unroll
            for (int i = 0; i < CPT; ++i) st_v4<SK>(rb + (size_t)i * T, make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu));
        }
    }
    const long long t1 = clock64();
    if (tid == 0) clk[blockIdx.x] = t1 - t0;
    if (bad) atomicAdd(errors, bad);
}
template <int CPT, bool SYNC, int LK, int SK>
static void run(uint4* X, uint4* Y, long long* dclk, int* derr, int clk_khz) {
    const size_t bytes = 3ull * kMaxP * kMaxC * kMaxT * 16;
    for (int P : {16})
        for (int T : {1, 32, 512}) {
            if (SYNC && T < 32) continue;
            cudaMemset(X, 0xFF, bytes); cudaMemset(Y, 0xFF, bytes); cudaMemset(derr, 0, 8);
            int rounds = 2000;
            long long limit = 4000000000LL;
            void* args[] = {&X, &Y, (void*)&P, (void*)&T, (void*)&rounds, &dclk, &derr, &limit};
            cudaError_t e = cudaLaunchCooperativeKernel((const void*)k<CPT, SYNC, LK, SK>, dim3(2 * P), dim3(512), args, 0, 0);
            cudaError_t e2 = cudaDeviceSynchronize();
            long long clk[64]; int err[2];
            cudaMemcpy(clk, dclk, 8 * 2 * P, cudaMemcpyDeviceToHost);
            cudaMemcpy(err, derr, 8, cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < 2 * P; ++i) mx = clk[i] > mx ? clk[i] : mx;
            printf("ld%d st%d CPT=%2d sync=%d P=%2d T=%3d (%6.1f KB per CTA): %s %s  %.0f clk = %.3f us per exchange  errors %d\n", LK, SK, CPT, (int)SYNC, P, T, CPT * T * 16 / 1024.0,
                   cudaGetErrorString(e), cudaGetErrorString(e2), (double)mx / (2.0 * rounds), (double)mx / (2.0 * rounds) / (clk_khz * 1e-3), err[0]);
        }
}
int main() {
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const size_t bytes = 3ull * kMaxP * kMaxC * kMaxT * 16;
    uint4 *X, *Y; long long* dclk; int* derr;
    cudaMalloc(&X, bytes); cudaMalloc(&Y, bytes); cudaMalloc(&dclk, 8 * 64); cudaMalloc(&derr, 8);
    run<16, false, 0, 0>(X, Y, dclk, derr, clk_khz);
    run<16, false, 1, 0>(X, Y, dclk, derr, clk_khz);
    run<16, false, 2, 0>(X, Y, dclk, derr, clk_khz);
    run<16, false, 3, 0>(X, Y, dclk, derr, clk_khz);
    run<16, false, 0, 1>(X, Y, dclk, derr, clk_khz);
    run<16, false, 1, 1>(X, Y, dclk, derr, clk_khz);
    run<16, false, 2, 1>(X, Y, dclk, derr, clk_khz);
    run<16, false, 2, 2>(X, Y, dclk, derr, clk_khz);
    run<16, false, 3, 3>(X, Y, dclk, derr, clk_khz);
    run<4, false, 1, 1>(X, Y, dclk, derr, clk_khz);
    run<4, false, 2, 1>(X, Y, dclk, derr, clk_khz);
    return 0;
}
