// xchg8.cu -- one producer CTA, C consumer CTAs.  Round: the producer writes L lines (thread i -> 16-byte chunk in line i,
// value = round), each consumer's threads poll "their" lines (consumer thread i polls line i; T = L threads) until all show
// the round, then consumer thread 0 writes an ack word in a private line; the producer polls the C acks (lane = consumer)
// and starts the next round.  time/round = (producer->consumers, L lines, C pollers per line) + (ack back, private lines).
// A second mode measures the ack path alone (L = 0: consumers poll nothing but a private go-flag from the producer).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned ld_u32(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_u32(unsigned* p, unsigned v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// data: [L lines][32 words]; acks: [C][32 words]; go: [C][32 words] (private go flags, mode 1)
__global__ void __launch_bounds__(512, 1) k(unsigned* data, unsigned* acks, unsigned* go, int C, int L, int mode, int rounds, long long* clk, long long limit) {
    const int tid = threadIdx.x;
    const long long t0 = clock64();
    bool dead = false;
    if (blockIdx.x == 0) {
        for (int r = 1; r <= rounds && !dead; ++r) {
            if (mode == 0) { if (tid < L) st_u32(data + (size_t)tid * 32, (unsigned)r); }
            else if (tid < C) st_u32(go + (size_t)tid * 32, (unsigned)r);
            if (tid < C) {
                int spins = 0;
                while (ld_u32(acks + (size_t)tid * 32) != (unsigned)r)
                    if (((++spins) & 4095) == 0 && clock64() - t0 > limit) { dead = true; break; }
            }
            dead = __syncthreads_or(dead);
        }
        if (tid == 0) clk[0] = clock64() - t0;
    } else {
        const int c = blockIdx.x - 1;
        for (int r = 1; r <= rounds && !dead; ++r) {
            int spins = 0;
            if (mode == 0) {
                if (tid < L)
                    while (ld_u32(data + (size_t)tid * 32) != (unsigned)r)
                        if (((++spins) & 4095) == 0 && clock64() - t0 > limit) { dead = true; break; }
            } else if (tid == 0) {
                while (ld_u32(go + (size_t)c * 32) != (unsigned)r)
                    if (((++spins) & 4095) == 0 && clock64() - t0 > limit) { dead = true; break; }
            }
            dead = __syncthreads_or(dead);
            if (tid == 0) st_u32(acks + (size_t)c * 32, (unsigned)r);
        }
    }
}
int main() {
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    unsigned *data, *acks, *go; long long* dclk;
    cudaMalloc(&data, 512 * 128); cudaMalloc(&acks, 64 * 128); cudaMalloc(&go, 64 * 128); cudaMalloc(&dclk, 64);
    for (int mode : {1, 0})
        for (int C : {1, 2, 4, 8, 16, 32})
            for (int L : {1, 16, 64, 256, 512}) {
                if (mode == 1 && L != 1) continue;
                cudaMemset(data, 0, 512 * 128); cudaMemset(acks, 0, 64 * 128); cudaMemset(go, 0, 64 * 128);
                int rounds = 3000; long long limit = 2000000000LL;
                void* args[] = {&data, &acks, &go, (void*)&C, (void*)&L, (void*)&mode, (void*)&rounds, &dclk, &limit};
                cudaError_t e = cudaLaunchCooperativeKernel((const void*)k, dim3(1 + C), dim3(512), args, 0, 0);
                cudaError_t e2 = cudaDeviceSynchronize();
                long long clk; cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost);
                printf("%s consumers=%2d lines=%3d: %s %s  %.0f clk = %.3f us per ROUND TRIP\n", mode ? "private go flags + private acks" : "shared data lines + private acks  ", C, L,
                       cudaGetErrorString(e), cudaGetErrorString(e2), (double)clk / rounds, (double)clk / rounds / (clk_khz * 1e-3));
                fflush(stdout);
            }
    return 0;
}
