// loop_tc2.cu -- cluster-local tensor-core sample loop (MOL): the second-generation fp16 loop.
//
// loop_tc.cu exchanges activations through L2 (publish -> fence -> counter -> acquire -> TMA: ~2.5 us per stage).
// Here a 16-CTA thread-block cluster owns <= 32 folds completely, so every exchange is a DSMEM copy into the peers'
// operand tiles plus a remote mbarrier arrive (measured floor of a 16-CTA cluster exchange: 0.6 us), and clusters
// never talk to each other (folds are partitioned across the 8 clusters; no co-residency requirement).
// The price: a cluster's shared memory (16 x 227 KB) cannot hold the fp16 loop weights (6 MB), so weights are not
// resident: every CTA streams its 414 KB of pre-swizzled weight tiles from L2 every step with cp.async.bulk through
// a 6-slot mbarrier ring.  The stream is data-independent, so the producer warp simply runs ahead of the chain.
//
// Orientation is swapped w.r.t. loop_tc.cu: D[128 weight rows (TMEM lanes)][32 folds (columns)] = W_tile * act^T,
// i.e. A = streamed weight tile (K-major, 128B swizzle), B = activation matrix [32 folds][512] (K-major) that the
// peers write straight into.  CTA c owns hidden units 32c..32c+31; tile rows are ordered gate-major so that TMEM
// lane quarter q (= epilogue warp & 3) is gate q of unit `lane`:
//   T0 = [W_ih2a r|z|n, W_fc1a]   T1 = [W_hh1 r|z|n]   T2 = [W_hh2 r|z|n, W_fc1a]   T3 = fc2 (lanes 96..127)
//   T4 = fc3 (all 30 rows, lanes 0..29; every CTA computes it, so the logits never need an exchange)
// Each gate warp evaluates its own gate; r and z reach the n-gate warp through 8 KB of shared memory.
#include "engine_internal.h"
#include "tc_common.cuh"

namespace wrnn {

__device__ long long g_tc2_deadline = 1500000000LL;

namespace {
using namespace tc;

constexpr int CL = 16, U = 32, NF = 32;           // cluster size, units per CTA, folds per cluster (MMA N)
constexpr int NEPI = 16, NT = (NEPI + 2) * 32;
constexpr int kSlots2 = 3, kSlotBytes2 = 32768;    // ring of 32 KB chunks: 14 bulk copies per step (a copy costs ~0.3 us, serialised)
constexpr int kActBytes = 8 * NF * 128;           // one activation matrix [8 k-blocks][32 folds x 128 B] = 32 KB
// shared memory map
constexpr int oActA = 0;
constexpr int oActB = oActA + kActBytes;                 //  32768
constexpr int oActC = oActB + kActBytes;                 //  65536  (f1 gets its own buffer: T1 may still read h1 in actA)
constexpr int oRing = oActC + kActBytes;                 //  98304  (>= 12 KB into the allocation: T3's A tile starts 96 rows early)
constexpr int oR = oRing + kSlots2 * kSlotBytes2;        // 196608  r gate  [32 u][32 f] fp32
constexpr int oZ = oR + U * NF * 4;                      // z gate
constexpr int oStg = oZ + U * NF * 4;                    // fp16 staging [32 f][32 u]
constexpr int oLg = oStg + NF * U * 2;                   // logits [32 f][32] fp32
constexpr int oMol = oLg + NF * 32 * 4;                  // [32 f][4] {score, index}
constexpr int oXs = oMol + NF * 4 * 8;                   // previous sample per fold
constexpr int oCtl = oXs + NF * 4;
constexpr int kSmem2 = oCtl + 512;
// TMEM columns of the five accumulators
constexpr int kAcc0 = 0, kAcc1 = 32, kAcc2 = 64, kAcc3 = 96, kAcc4 = 128, kTmem2 = 256;

struct Ctl2 {
    uint64_t full[kSlots2];
    uint64_t empty[kSlots2];
    uint64_t accfull[5];
    uint64_t actready[4];
    uint32_t tmem;
    int abort_local;
};

__device__ __forceinline__ bool aborted2(Ctl2* c) { return *reinterpret_cast<volatile int*>(&c->abort_local) != 0; }
__device__ __noinline__ bool spin_check2(const Tc2Params& p, Ctl2* c, long long& t0) {
    if (aborted2(c)) return true;
    if (ld_volatile_i32(p.abort_flag) != 0) { *reinterpret_cast<volatile int*>(&c->abort_local) = 1; return true; }
    if (t0 == 0) t0 = clock64();
    if (clock64() - t0 > g_tc2_deadline) {
        *reinterpret_cast<volatile int*>(&c->abort_local) = 1;
        atomicExch(p.abort_flag, 1);
        return true;
    }
    return false;
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
template <bool kCluster>
__device__ __forceinline__ bool wait2(const Tc2Params& p, Ctl2* c, uint64_t* bar, uint32_t parity) {
    long long t0 = 0;
    int spins = 0;
    while (!(kCluster ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity))) {
        if (((++spins) & 63) == 0 && aborted2(c)) return false;
        if ((spins & 4095) == 0 && spin_check2(p, c, t0)) return false;
    }
    return true;
}
// arrive (release, cluster scope) on the same mbarrier of CTA `rank`
__device__ __forceinline__ void remote_arrive(uint64_t* local_bar, uint32_t rank) {
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local_bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void st_remote_v4(uint32_t local_addr, uint32_t rank, uint4 v) {
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(local_addr), "r"(rank));
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NEPI * 32) : "memory"); }
__device__ __forceinline__ float sigmoid_fast2(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast2(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

constexpr int kTr0 = 64, kTrN = 16;
__device__ __forceinline__ void trace2(const Tc2Params& p, int t, int slot) {
    if (p.trace && blockIdx.x == 0 && t >= kTr0 && t < kTr0 + kTrN) p.trace[(t - kTr0) * 32 + slot] = clock64();
}

// rows of the five weight tiles and the byte offset at which a k-block lands inside a ring slot
// The five weight tiles stream as 14 chunks per step.  Per tile: chunks, k-blocks per chunk, bytes per k-block, and the
// byte offset of the A tile's row 0 relative to a k-block's data (T3 = fc2 holds rows 96..127 only: its tile starts
// 96 rows before the data; T4 = fc3 holds rows 0..31, the rows above read whatever follows: unused TMEM lanes).
__device__ __constant__ int kChunks[5] = {4, 4, 4, 1, 1};
__device__ __constant__ int kKbPerChunk[5] = {2, 2, 2, 8, 8};
__device__ __constant__ int kKbBytes[5] = {128 * 128, 96 * 128, 128 * 128, 32 * 128, 32 * 128};
__device__ __constant__ int kRow0[5] = {0, 0, 0, -96 * 128, 0};

}  // namespace

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT, 1) wrnn_loop_tc2_kernel(Tc2Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw2[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw2) + 1023) & ~(uintptr_t)1023);
    Ctl2* ctl = reinterpret_cast<Ctl2*>(smem + oCtl);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int cl = blockIdx.x / CL;
    const int f0 = cl * p.Bc, nf = min(p.Bc, p.B - f0);          // my cluster's folds
    float* sR = reinterpret_cast<float*>(smem + oR);
    float* sZ = reinterpret_cast<float*>(smem + oZ);
    __half* sStg = reinterpret_cast<__half*>(smem + oStg);
    float* sLg = reinterpret_cast<float*>(smem + oLg);
    float2* sMol = reinterpret_cast<float2*>(smem + oMol);
    float* sXs = reinterpret_cast<float*>(smem + oXs);

    for (int i = tid; i < (3 * kActBytes) / 16; i += NT) reinterpret_cast<uint4*>(smem + oActA)[i] = make_uint4(0, 0, 0, 0);
    if (tid < NF) sXs[tid] = 0.f;
    fence_proxy_async_smem();
    if (tid == 0) {
        for (int i = 0; i < kSlots2; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], 1); }
        for (int i = 0; i < 5; ++i) mbar_init(&ctl->accfull[i], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&ctl->actready[i], CL);
        ctl->abort_local = 0;
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&ctl->tmem, kTmem2);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = ctl->tmem;
    // every CTA's barriers must be initialised before any peer arrives on them
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");

    if (warp == NEPI) {
        // =================================== weight-stream producer =========================================
        if (lane == 0) {
            const uint8_t* img = p.wimg + (size_t)crank * p.img_bytes;
            uint32_t q = 0;
            for (int t = 0; t < p.S; ++t) {
                const uint8_t* src = img;
                for (int tile = 0; tile < 5; ++tile) {
                    const uint32_t bytes = (uint32_t)(kKbPerChunk[tile] * kKbBytes[tile]);
                    for (int c = 0; c < kChunks[tile]; ++c, ++q, src += bytes) {
                        const uint32_t slot = q % kSlots2, round = q / kSlots2;
                        bool go = true;
                        if (round > 0) go = wait2<false>(p, ctl, &ctl->empty[slot], (round - 1) & 1);
                        if (go && !aborted2(ctl)) {
                            mbar_arrive_expect_tx(&ctl->full[slot], bytes);
                            bulk_g2s(smem + oRing + slot * kSlotBytes2, src, bytes, &ctl->full[slot]);
                        }
                    }
                }
            }
        }
    } else if (warp == NEPI + 1) {
        // =================================== MMA issuer =====================================================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, NF);
            const uint32_t acc[5] = {kAcc0, kAcc1, kAcc2, kAcc3, kAcc4};
            const int actofs[5] = {oActA, oActA, oActB, oActC, oActB};   // activation matrix read by each tile
            const int ready[5] = {0, -1, 1, 2, 3};               // exchange that must be complete first
            uint32_t q = 0;
            for (int t = 0; t < p.S; ++t) {
                const uint32_t par = (uint32_t)t & 1u;
                for (int tile = 0; tile < 5; ++tile) {
                    bool ok = true;
                    trace2(p, t, 10 + tile);
                    if (ready[tile] >= 0) ok = wait2<true>(p, ctl, &ctl->actready[ready[tile]], par);
                    fence_proxy_async_smem();
                    trace2(p, t, 15 + tile);                     // peers' generic-proxy stores -> my async-proxy reads
                    const uint32_t bbase = smem_u32(smem + actofs[tile]);
                    const uint32_t dcol = tmem + acc[tile];
                    int kb = 0;
                    for (int c = 0; c < kChunks[tile]; ++c, ++q) {
                        const uint32_t slot = q % kSlots2, round = q / kSlots2;
                        ok = wait2<false>(p, ctl, &ctl->full[slot], round & 1) && ok;
                        tcgen05_fence_after();
                        if (ok) {
                            const uint32_t a0 = smem_u32(smem + oRing + slot * kSlotBytes2) + (uint32_t)kRow0[tile];
                            for (int kk = 0; kk < kKbPerChunk[tile]; ++kk, ++kb) {
                                const uint64_t ad = umma_desc_sw128(a0 + kk * kKbBytes[tile]);
                                const uint64_t bd = umma_desc_sw128(bbase + kb * (NF * 128));
                                if (kb == 0) umma_f16_c<false>(dcol, ad, bd, idesc); else umma_f16_c<true>(dcol, ad, bd, idesc);
                                umma_f16_c<true>(dcol, umma_desc_advance(ad, 32), umma_desc_advance(bd, 32), idesc);
                                umma_f16_c<true>(dcol, umma_desc_advance(ad, 64), umma_desc_advance(bd, 64), idesc);
                                umma_f16_c<true>(dcol, umma_desc_advance(ad, 96), umma_desc_advance(bd, 96), idesc);
                            }
                        }
                        umma_commit(&ctl->empty[slot]);
                    }
                    umma_commit(&ctl->accfull[tile]);
                    trace2(p, t, 20 + tile);
                }
            }
        }
    } else {
        // =================================== epilogue warps =================================================
        const int q = warp & 3, fg = warp >> 2, u = lane;        // gate / TMEM lane quarter, fold group, my unit
        const int j = (int)crank * U + u;                        // global hidden unit
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16) + 8 * fg;
        const float v1g = (q < 3) ? p.v1[q * kRnn + j] : 0.f, v2g = (q < 3) ? p.v2[q * kRnn + j] : 0.f;
        const float v3u = p.v3[j], bh1 = p.bhn1[j], bh2 = p.bhn2[j];
        float h1[8], h2[8], p3[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { h1[i] = 0.f; h2[i] = 0.f; p3[i] = 0.f; }
        const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
        const uint32_t actA = smem_u32(smem + oActA), actB = smem_u32(smem + oActB), actC = smem_u32(smem + oActC);
        // DSMEM copy role: 16-byte chunk (fold cf, 8 units cp*8..) to destinations 4*cd .. 4*cd+3
        const int cf = (tid & 127) >> 2, cp = tid & 3, cd = tid >> 7;
        const uint32_t chunk_off = (uint32_t)((crank >> 1) * (NF * 128) + cf * 128 + ((((crank & 1) * 4 + cp) ^ (cf & 7)) << 4));

        auto exchange = [&](uint32_t act_base, int e, uint64_t* guard, uint32_t gpar) {
            epi_bar();                                            // staging tile complete
            const uint4 v = *reinterpret_cast<const uint4*>(sStg + cf * U + cp * 8);
#pragma unroll
            for (int k = 0; k < 4; ++k) st_remote_v4(act_base + chunk_off, (uint32_t)(cd * 4 + k), v);
            if (guard) wait2<false>(p, ctl, guard, gpar);
            epi_bar();                                            // all my stores issued
            if (tid < CL) remote_arrive(&ctl->actready[e], (uint32_t)tid);
        };

        // conditioning {a, b}[8 folds] for my (gate, unit): gates (c1, c2); q == 3: (fc1, fc2).  Fetched one step ahead
        // (HBM latency stays off the chain).
        auto cs_ptr = [&](int t) {
            return reinterpret_cast<const float4*>(p.CS + ((((((size_t)cl * p.S + t) * CL + crank) * 4 + q) * U + u) * NF + 8 * fg) * 2);
        };
        float4 nx[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) nx[i] = __ldcs(cs_ptr(0) + i);
        for (int t = 0; t < p.S; ++t) {
            const uint32_t par = (uint32_t)t & 1u;
            float ca[8], cb[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) { ca[2 * i] = nx[i].x; cb[2 * i] = nx[i].y; ca[2 * i + 1] = nx[i].z; cb[2 * i + 1] = nx[i].w; }
            if (t + 1 < p.S) {
#pragma unroll
                for (int i = 0; i < 4; ++i) nx[i] = __ldcs(cs_ptr(t + 1) + i);
            }
            if (tid == 0) trace2(p, t, 0);
            float x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = sXs[8 * fg + i];
            // ---- A: GRU1 -> h1 -> everyone's actA ------------------------------------------------------------
            float g[8];
            if (t > 0 && q < 3) { tmem_ld8(tl + kAcc1, g); tmem_ld_wait(); }
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) g[i] = 0.f;
            }
            if (q == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) sR[u * NF + 8 * fg + i] = sigmoid_fast2(fmaf(v1g, x[i], ca[i]) + g[i]);
            } else if (q == 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) sZ[u * NF + 8 * fg + i] = sigmoid_fast2(fmaf(v1g, x[i], ca[i]) + g[i]);
            }
            epi_bar();
            if (q == 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float r = sR[u * NF + 8 * fg + i], z = sZ[u * NF + 8 * fg + i];
                    const float n = tanh_fast2(fmaf(v1g, x[i], ca[i]) + r * (g[i] + bh1));
                    h1[i] = (1.0f - z) * n + z * h1[i];
                    sStg[(8 * fg + i) * U + u] = __float2half_rn(h1[i]);
                }
            }
            tcgen05_fence_before();
            exchange(actA, 0, nullptr, 0);
            if (tid == 0) trace2(p, t, 1);
            // ---- B: T0 = [W_ih2a | W_fc1a] h1 ; GRU2 -> h2 -> actB -----------------------------------------------
            float d[8];
            wait2<false>(p, ctl, &ctl->accfull[0], par);
            tcgen05_fence_after();
            if (tid == 0) trace2(p, t, 2);
            tmem_ld8(tl + kAcc0, d);
            if (t > 0 && q < 3) tmem_ld8(tl + kAcc2, g);
            tmem_ld_wait();
            if (q == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) sR[u * NF + 8 * fg + i] = sigmoid_fast2(d[i] + fmaf(v2g, x[i], cb[i]) + g[i]);
            } else if (q == 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) sZ[u * NF + 8 * fg + i] = sigmoid_fast2(d[i] + fmaf(v2g, x[i], cb[i]) + g[i]);
            } else if (q == 3) {
#pragma unroll
                for (int i = 0; i < 8; ++i) p3[i] = d[i];
            }
            epi_bar();
            if (q == 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float r = sR[u * NF + 8 * fg + i], z = sZ[u * NF + 8 * fg + i];
                    const float n = tanh_fast2(d[i] + fmaf(v2g, x[i], cb[i]) + r * (g[i] + bh2));
                    h2[i] = (1.0f - z) * n + z * h2[i];
                    sStg[(8 * fg + i) * U + u] = __float2half_rn(h2[i]);
                }
            }
            tcgen05_fence_before();
            exchange(actB, 1, &ctl->accfull[1], par);
            if (tid == 0) trace2(p, t, 3);             // T1 (reads actA) must be done before f1 may land in actA
            // ---- C: T2 = [W_hh2 | W_fc1a] h2 ; f1 -> actA ---------------------------------------------------------
            wait2<false>(p, ctl, &ctl->accfull[2], par);
            tcgen05_fence_after();
            if (tid == 0) trace2(p, t, 4);
            if (q == 3) {
                tmem_ld8(tl + kAcc2, d);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    sStg[(8 * fg + i) * U + u] = __float2half_rn(fmaxf(p3[i] + d[i] + fmaf(v3u, x[i], ca[i]), 0.f));
            }
            tcgen05_fence_before();
            exchange(actC, 2, nullptr, 0);
            if (tid == 0) trace2(p, t, 5);
            // ---- D: T3 = fc2 (lanes 96..127) ; f2 -> actB ---------------------------------------------------------
            wait2<false>(p, ctl, &ctl->accfull[3], par);
            tcgen05_fence_after();
            if (tid == 0) trace2(p, t, 6);
            if (q == 3) {
                tmem_ld8(tl + kAcc3, d);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) sStg[(8 * fg + i) * U + u] = __float2half_rn(fmaxf(d[i] + cb[i], 0.f));
            }
            tcgen05_fence_before();
            exchange(actB, 3, nullptr, 0);
            if (tid == 0) trace2(p, t, 7);
            // ---- E: T4 = fc3 (every CTA has all 30 outputs) ; mixture-of-logistics draw -------------------------------
            wait2<false>(p, ctl, &ctl->accfull[4], par);
            tcgen05_fence_after();
            if (tid == 0) trace2(p, t, 8);
            if (q == 0) {
                tmem_ld8(tl + kAcc4, d);
                tmem_ld_wait();
                if (u < 30) {
                    const float b = __ldg(p.bfc3 + u);
#pragma unroll
                    for (int i = 0; i < 8; ++i) sLg[(8 * fg + i) * 32 + u] = d[i] + b;
                }
            }
            tcgen05_fence_before();
            epi_bar();
            if (tid < 128) {
                // thread = (fold, Philox block part): mixtures 4*part .. 4*part+3 (vocoder/distribution.py:123-125)
                const int f = tid & 31, part = tid >> 5;
                const FoldDesc fd = p.folds[f0 + min(f, nf - 1)];
                const uint4 r = philox4x32_10(make_uint4((uint32_t)t, (uint32_t)fd.fold, (uint32_t)fd.utt, (uint32_t)min(part, 2)), key);
                float best = -INFINITY;
                int kbest = 0;
                if (part < 3) {
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const int i = part * 4 + w;
                        if (i < 10) {
                            const float um = 1e-5f + u01(word_of(r, w)) * (1.0f - 2e-5f);
                            const float sc = sLg[f * 32 + i] - __logf(-__logf(um));
                            if (sc > best) { best = sc; kbest = i; }
                        }
                    }
                }
                sMol[f * 4 + part] = make_float2(best, __int_as_float(kbest));
                asm volatile("bar.sync 2, 128;" ::: "memory");
                if (part == 2) {
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const float2 c = sMol[f * 4 + k];
                        if (c.x >= best && !(c.x == best && __float_as_int(c.y) > kbest)) { best = c.x; kbest = __float_as_int(c.y); }
                    }
                    const float mean = sLg[f * 32 + 10 + kbest];
                    const float lsc = fmaxf(sLg[f * 32 + 20 + kbest], -32.23619130191664f);
                    const float ul = 1e-5f + u01(r.z) * (1.0f - 2e-5f);
                    float xv = mean + __expf(lsc) * (__logf(ul) - __logf(1.0f - ul));
                    xv = fminf(fmaxf(xv, -1.0f), 1.0f);
                    if (f < nf) {
                        if (crank == 0) {
                            p.samples[(size_t)(f0 + f) * p.S + t] = xv;
                            if (p.logits_out)
                                for (int i = 0; i < 30; ++i) p.logits_out[((size_t)(f0 + f) * p.S + t) * 30 + i] = sLg[f * 32 + i];
                        }
                        sXs[f] = p.forced ? p.forced[(size_t)(f0 + f) * p.S + t] : xv;
                    }
                }
            }
            epi_bar();
            if (tid == 0) trace2(p, t, 9);
            if (blockIdx.x == 0 && tid == 0 && (t % 100) == 0 && p.progress) {
                *reinterpret_cast<volatile int*>(p.progress) = t;
                __threadfence_system();
            }
        }
    }
    // ---- teardown: nobody leaves while peers may still write into its shared memory / barriers ------------------------
    if (aborted2(ctl)) __nanosleep(200000);
    tcgen05_fence_before();
    __syncthreads();
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) tmem_dealloc(tmem, kTmem2);
}

cudaError_t set_tc2_deadline(long long cycles) { return cudaMemcpyToSymbol(g_tc2_deadline, &cycles, sizeof(cycles)); }
size_t loop_tc2_image_bytes() { return (size_t)(128 + 96 + 128 + 32 + 32) * 1024; }

int loop_tc2_max_clusters() {
    static int cached = -1;
    if (cached >= 0) return cached;
    if (cudaFuncSetAttribute(wrnn_loop_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem2 + 1024) != cudaSuccess) return 0;
    if (cudaFuncSetAttribute(wrnn_loop_tc2_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL * 8);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = kSmem2 + 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, wrnn_loop_tc2_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
    cached = n;
    return n;
}

cudaError_t launch_loop_tc2(const Tc2Params& p, int n_clusters, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(wrnn_loop_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem2 + 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(wrnn_loop_tc2_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    wrnn_loop_tc2_kernel<<<n_clusters * CL, NT, kSmem2 + 1024, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace wrnn
