// xchg7.cu -- timeline of the all-to-all exchange with %globaltimer stamps (common to all SMs): per round and CTA the
// time it published and the time its ingest completed; the host prints, per round, (last publish of the producing side ->
// each consumer's completion).  Also prints the granularity of %globaltimer.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <algorithm>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint4 ld_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_v4(uint4* p, uint4 v) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ bool chunk_ready(uint4 v) {
    return (__vcmpeq2(v.x, 0xFFFFFFFFu) | __vcmpeq2(v.y, 0xFFFFFFFFu) | __vcmpeq2(v.z, 0xFFFFFFFFu) | __vcmpeq2(v.w, 0xFFFFFFFFu)) == 0u;
}
constexpr int kChunks = 64, kRows = 128, P = 16, R0 = 200, NR = 24;
__global__ void __launch_bounds__(512, 1) k(uint4* X, uint4* Y, int NF, int rounds, unsigned long long* stamps, int* errors, long long limit) {
    const int side = blockIdx.x / P, cta = blockIdx.x % P, tid = threadIdx.x;
    uint4* out = side == 0 ? X : Y;
    const uint4* in = side == 0 ? Y : X;
    const int row = tid & 127, cg = tid >> 7;
    const bool live = row < NF;
    const size_t bufsz = (size_t)kChunks * kRows;
    uint32_t acc = 1;
    int bad = 0;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        if (!(side == 0 && r == 0)) {
            const int rr = side == 0 ? r - 1 : r;
            const uint4* base = in + (size_t)(rr % 3) * bufsz + (size_t)(cg * 16) * kRows + row;
            int passes = 0;
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
                    ++passes;
                    if (pending && ((++spins) & 4095) == 0 && clock64() - t0 > limit) { bad += 1000; break; }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) { if (v[i].x != (uint32_t)rr) ++bad; acc += v[i].y; }
            }
            if (tid == 0 && r >= R0 && r < R0 + NR) { stamps[((size_t)(r - R0) * 2 * P + blockIdx.x) * 4 + 0] = gtime(); stamps[((size_t)(r - R0) * 2 * P + blockIdx.x) * 4 + 3] = passes; }
            if (__syncthreads_or(bad >= 1000)) break;
            if (tid == 0 && r >= R0 && r < R0 + NR) stamps[((size_t)(r - R0) * 2 * P + blockIdx.x) * 4 + 1] = gtime();
        }
        if (live) {
            const size_t o = (size_t)(4 * cta + cg) * kRows + row;
            st_v4(out + (size_t)(r % 3) * bufsz + o, make_uint4((uint32_t)r, acc & 0x7FFF7FFFu, 0x3C003C00u, 0x3C003C00u));
            st_v4(out + (size_t)((r + 1) % 3) * bufsz + o, make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu));
        }
        if (tid == 0 && r >= R0 && r < R0 + NR) stamps[((size_t)(r - R0) * 2 * P + blockIdx.x) * 4 + 2] = gtime();
    }
    if (bad) atomicAdd(errors, bad);
}
__global__ void gran(unsigned long long* out) { for (int i = 0; i < 64; ++i) out[i] = gtime(); }
int main() {
    const size_t bytes = 3ull * kChunks * kRows * 16;
    uint4 *X, *Y; unsigned long long* st; int* derr;
    cudaMalloc(&X, bytes); cudaMalloc(&Y, bytes); cudaMalloc(&st, 8 * 4 * 2 * P * NR); cudaMalloc(&derr, 8);
    gran<<<1, 1>>>(st); cudaDeviceSynchronize();
    unsigned long long g[64]; cudaMemcpy(g, st, 8 * 64, cudaMemcpyDeviceToHost);
    printf("globaltimer deltas (ns):"); for (int i = 1; i < 24; ++i) printf(" %llu", g[i] - g[i - 1]); printf("\n");
    for (int NF : {16, 72}) {
        cudaMemset(X, 0xFF, bytes); cudaMemset(Y, 0xFF, bytes); cudaMemset(derr, 0, 8); cudaMemset(st, 0, 8 * 4 * 2 * P * NR);
        int rounds = 400; long long limit = 2000000000LL;
        void* args[] = {&X, &Y, (void*)&NF, (void*)&rounds, &st, &derr, &limit};
        cudaLaunchCooperativeKernel((const void*)k, dim3(2 * P), dim3(512), args, 0, 0);
        cudaDeviceSynchronize();
        static unsigned long long h[4 * 2 * P * NR]; cudaMemcpy(h, st, sizeof(h), cudaMemcpyDeviceToHost);
        printf("folds %d: per round: side-0 publish spread | side-1: (thread0 done - last side-0 publish) min/median/max, passes of thread 0 min/max, sync wait max | same for side 0 reading side 1\n", NF);
        for (int r = 1; r < NR; ++r) {
            for (int s = 0; s < 2; ++s) {           // producers = side s in round r (side 0) / r (side 1 publishes round r after reading)
                // consumers of side-0's round-r data: side 1 in round r; consumers of side-1's round-r data: side 0 in round r+1
                const int rc = s == 0 ? r : r + 1;
                if (rc >= NR) continue;
                unsigned long long pmin = ~0ull, pmax = 0;
                for (int c = 0; c < P; ++c) { unsigned long long t = h[((size_t)r * 2 * P + s * P + c) * 4 + 2]; pmin = std::min(pmin, t); pmax = std::max(pmax, t); }
                long long d[P], sy = 0; unsigned long long pa_min = ~0ull, pa_max = 0;
                for (int c = 0; c < P; ++c) {
                    const size_t o = ((size_t)rc * 2 * P + (1 - s) * P + c) * 4;
                    d[c] = (long long)(h[o + 0] - pmax); sy = std::max<long long>(sy, (long long)(h[o + 1] - h[o + 0]));
                    pa_min = std::min(pa_min, h[o + 3]); pa_max = std::max(pa_max, h[o + 3]);
                }
                std::sort(d, d + P);
                printf("  r%2d side%d: publish spread %5llu ns | seen after last publish %5lld / %5lld / %5lld ns, passes %llu..%llu, sync wait <= %lld ns", r, s, pmax - pmin, d[0], d[P / 2], d[P - 1], pa_min, pa_max, sy);
            }
            printf("\n");
        }
    }
    return 0;
}
