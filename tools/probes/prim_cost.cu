// prim_cost.cu -- latency of the primitives on the role-specialised loop's chain, measured with clock64 by lane 0 of warp 0
// while `nw` warps (1, 4 on one scheduler, 16) execute the same sequence: tcgen05.st + wait, tcgen05.ld + wait, the fences,
// mbarrier arrive / try_wait, a single MMA -> commit -> try_wait round trip, and a 16-deep batch of L2 loads.
#include <cstdio>
#include <cstdlib>
#include "../../real-time-voice-cloning_b200/csrc/tc_common.cuh"
using namespace wrnn::tc;
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
                 "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void st16(uint32_t t, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(t), "r"(v) : "memory");
}
__global__ void __launch_bounds__(640, 1) k(int nw, int sel, long long* out, const uint4* g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[20], mb;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 8 * 96 * 128 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if (tid == 0) { for (int i = 0; i < 20; ++i) mbar_init(&bar[i], 1); mbar_init(&mb, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tslot, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tslot;
    // active warps: nw == 4 -> warps 0, 4, 8, 12 (one scheduler, one lane quadrant); else the first nw
    const bool active = nw == 4 ? ((warp & 3) == 0 && warp < 16) : warp < nw;
    const uint32_t tl = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + 16 * (warp >> 2);
    const int iters = 200;
    long long acc = 0;
    uint32_t ph = 0;
    float f[8], s = 0.f;
    if (active) {
        for (int it = 0; it < iters; ++it) {
            __syncwarp();
            const long long t0 = clock64();
            if (sel == 0) { st16(tl, it); asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
            else if (sel == 1) { st16(tl, it); st16(tl + 64, it); st16(tl + 128, it); st16(tl + 192, it); asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
            else if (sel == 2) { tmem_ld8(tl + 256, f); tmem_ld_wait(); s += f[0]; }
            else if (sel == 3) { tmem_ld8(tl + 256, f); s += f[0]; tmem_ld8(tl + 320, f); s += f[1]; tmem_ld8(tl + 384, f); tmem_ld_wait(); s += f[2]; }
            else if (sel == 4) { tcgen05_fence_before(); }
            else if (sel == 5) { tcgen05_fence_after(); }
            else if (sel == 6) { if (lane == 0) mbar_arrive(&bar[warp]); __syncwarp(); while (!mbar_try_wait(&bar[warp], ph)) {} ph ^= 1; }
            else if (sel == 7) {          // one MMA -> commit -> wait (warp 0 only meaningful)
                if (warp == 0) {
                    if (elect_one()) { umma_ts(tmem + 256, tmem, umma_desc_sw128(smem_u32(smem)), umma_idesc_f16(128, 32), 0u); umma_commit(&mb); }
                    __syncwarp();
                    while (!mbar_try_wait(&mb, ph)) {}
                    ph ^= 1; tcgen05_fence_after();
                }
            } else if (sel == 8) {        // 8 MMAs (N = 48) -> commit -> wait
                if (warp == 0) {
                    if (elect_one()) {
                        const uint64_t bd = umma_desc_sw128(smem_u32(smem));
                        for (int q = 0; q < 8; ++q) umma_ts(tmem + 256, tmem + 8 * q, bd + 2u * (q & 3), umma_idesc_f16(128, 48), q ? 1u : 0u);
                        umma_commit(&mb);
                    }
                    __syncwarp();
                    while (!mbar_try_wait(&mb, ph)) {}
                    ph ^= 1; tcgen05_fence_after();
                }
            } else if (sel == 9 || sel == 10) {        // L2 loads: 4 / 16 x 16 bytes per lane, strong
                const int n = sel == 9 ? 4 : 16;
                uint4 w[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) if (i < n) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[i].x), "=r"(w[i].y), "=r"(w[i].z), "=r"(w[i].w) : "l"(g + (size_t)i * 2048 + tid + 64 * (it & 7)) : "memory");
                uint32_t o = 0;
#pragma unroll
                for (int i = 0; i < 16; ++i) if (i < n) o |= w[i].x | w[i].w;
                s += (float)o;
            } else if (sel == 11) {       // st + wait + fence + arrive, the per-quarter tail of an ingest
                st16(tl, it); asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); tcgen05_fence_before(); if (lane == 0) mbar_arrive(&bar[16]);
            }
            __syncwarp();
            acc += clock64() - t0;
        }
    }
    if (tid == 0) out[0] = acc / iters;
    if (s == 12345.f) out[1] = 1;
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    uint4* g; cudaMalloc(&g, (size_t)16 * 2048 * 16 + 1024 * 16 * 16); cudaMemset(g, 0, (size_t)16 * 2048 * 16 + 1024 * 16 * 16);
    const int smem = 8 * 96 * 128;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* names[] = {"tcgen05.st x16 + wait", "4 x tcgen05.st x16 + wait", "tcgen05.ld x8 + wait", "3 x tcgen05.ld x8 + wait", "fence::before_thread_sync", "fence::after_thread_sync",
                           "mbarrier arrive + try_wait", "1 MMA + commit + wait", "8 MMAs N=48 + commit + wait", "4 x 16 B L2 loads", "16 x 16 B L2 loads", "st + wait + fence + arrive"};
    for (int sel = 0; sel < 12; ++sel) {
        printf("%-30s:", names[sel]);
        for (int nw : {1, 4, 16}) {
            k<<<1, 640, smem>>>(nw, sel, d, g);
            cudaError_t e = cudaDeviceSynchronize();
            long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("  %2d warps %5lld clk%s", nw, h, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
        printf("\n");
    }
    return 0;
}
