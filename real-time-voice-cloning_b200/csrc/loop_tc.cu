// loop_tc.cu -- the autoregressive sample loop on the 5th-gen tensor cores (fp16 operands, fp32 accumulate,
// fp32 recurrent state): the batched-fold path (BASELINE configs 1, 3, 5).
//
// Same five-exchange step as loop_f32.cu, restructured for tcgen05:
//  * 2 groups x 64 CTAs (one CTA per SM).  A group owns one or two SETS of up to 128 folds (MMA M = 128: one fold
//    per TMEM lane); each of its CTAs owns 8 hidden units of every layer.  With two sets the CTA is software-pipelined:
//    a stage of set 0 is followed by the same stage of set 1, so while one set's activations travel through L2 and
//    its MMAs run, the epilogue warps do the other set's math (the resident weights serve 256 folds per group).  Its weight rows -- [W_ih2a|W_fc1a|W_hh1] (stage B),
//    [W_hh2|W_fc1a] (C), fc2 (D), fc3 (E) -- stay in shared memory as K-major SWIZZLE_128B tiles (fp16,
//    ~144 KB) for the whole sequence and are the B operand of tcgen05.mma (N = 64/32/16/16|32).
//  * activations h1, h2, f1, f2 travel as fp16 rows [fold][512] through L2.  A stage = publish my 8 columns,
//    arrive on the group's counter (release), the producer thread acquires the counter, then TMA-loads the
//    group's [128 x 512] activation matrix in eight [128 x 64] swizzled tiles through a 4-slot mbarrier ring;
//    one thread issues 32 tcgen05.mma per stage into TMEM; 16 epilogue warps (thread = fold x 2 units) read the
//    accumulator with tcgen05.ld and do the GRU / ReLU / sampling math in fp32.
//  * conditioning arrives pre-interpolated per sample (cond.cu: expand_cond), 64 B per thread per step.
//  * "soft abort": a wait that passes its deadline raises a flag; from then on every wait returns at once, so
//    all warps still walk the same barriers and the kernel ends cleanly (never a hung GPU).
#include "engine_internal.h"
#include "sampling.cuh"
#include "tc_common.cuh"

namespace wrnn {

__device__ long long g_tc_deadline = 1500000000LL;

namespace {
using namespace tc;

constexpr int NEPI = 16;                    // epilogue warps
constexpr int NT = (NEPI + 2) * 32;         // + TMA producer warp + MMA warp
constexpr int kSlots = 4, kMaxSlots = 8, kKB = 64, kNKB = kRnn / kKB;   // ring: kSlots x 16 KB, re-cut into up to 8 smaller slots
constexpr int kTileBytes = 128 * 128;       // one [128 rows x 64 fp16] activation tile
constexpr int NB_ = 64, NC_ = 32, ND_ = 16; // MMA N per stage (E: 16 RAW / 32 MOL)
constexpr int kWB = 0, kWC = kWB + NB_ * 128 * kNKB, kWD = kWC + NC_ * 128 * kNKB, kWE = kWD + ND_ * 128 * kNKB;
constexpr int kWBytes = kWE + 32 * 128 * kNKB;          // 147456
constexpr int kRing = kWBytes;                          // 4 x 16 KB
constexpr int kBars = kRing + kSlots * kTileBytes;      // mbarriers + misc
constexpr int kMolScratchBytes = 128 * 4 * 8;           // [128 rows][4] {score, index} for the cooperative MOL draw, per fold set
constexpr int kMolScratch = kBars + 256;
constexpr int kBias = kMolScratch + kTcSets * kMolScratchBytes;   // fc3 bias (MOL)
// tcgen05.mma always reads 128 rows (16 KB) from a slot base; with slots shorter than that the last slot reads up to
// 8 KB past the ring (into the control words: harmless garbage rows), so the allocation must cover ring + 72 KB
constexpr int kSmemBytes = (kBias + 128 > kRing + 73728) ? (kBias + 128) : (kRing + 73728);
// TMEM columns
constexpr int kAccB = 0, kAccC = 64, kAccD = 96, kAccE = 112, kSetCols = 256, kTmemCols = kTcSets * kSetCols;   // per fold set

struct Ctl {
    uint64_t full[kMaxSlots];
    uint64_t empty[kMaxSlots];
    uint64_t accfull[4 * kTcSets];
    uint32_t tmem;
    int abort_local;
};
static_assert(sizeof(Ctl) <= 256, "control block");

// Abort state: a CTA-local flag in shared memory (cheap to poll) mirrors the global flag (polled rarely: a
// global load costs ~0.7 us and must stay off the wait paths).
__device__ __forceinline__ bool aborted_local(Ctl* c) { return *reinterpret_cast<volatile int*>(&c->abort_local) != 0; }
__device__ __forceinline__ bool aborted(const TcParams& p, Ctl* c) { return aborted_local(c); }
__device__ __forceinline__ void raise_abort(const TcParams& p, Ctl* c) {
    *reinterpret_cast<volatile int*>(&c->abort_local) = 1;
    atomicExch(p.abort_flag, 1);
}
// slow path of every spin loop: returns true when the wait must be abandoned
__device__ __noinline__ bool spin_check(const TcParams& p, Ctl* c, long long& t0) {
    if (aborted_local(c)) return true;
    if (ld_volatile_i32(p.abort_flag) != 0) { *reinterpret_cast<volatile int*>(&c->abort_local) = 1; return true; }
    if (t0 == 0) t0 = clock64();
    if (clock64() - t0 > g_tc_deadline) { raise_abort(p, c); return true; }
    return false;
}
// All waits go through these: they return false (and everything keeps moving) once the abort flag is up.
__device__ __forceinline__ bool wait_mbar(const TcParams& p, Ctl* c, uint64_t* bar, uint32_t parity) {
    long long t0 = 0;
    int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (((++spins) & 63) == 0 && aborted_local(c)) return false;
        if ((spins & 4095) == 0 && spin_check(p, c, t0)) return false;
    }
    return true;
}
__device__ __forceinline__ bool wait_counter(const TcParams& p, Ctl* c, const unsigned int* ctr, unsigned int want) {
    long long t0 = 0;
    int spins = 0;
    while (true) {
        unsigned int v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        if (v >= want) return true;
        if (((++spins) & 255) == 0 && spin_check(p, c, t0)) return false;
    }
}
__device__ __forceinline__ bool wait_x(const TcParams& p, Ctl* c, const unsigned long long* w, uint32_t tag, float& v) {
    long long t0 = 0;
    int spins = 0;
    while (true) {
        const unsigned long long q = ll_load(w);
        if (ll_tag(q) == tag) { v = ll_val(q); return true; }
        if (((++spins) & 255) == 0 && spin_check(p, c, t0)) { v = 0.f; return false; }
    }
}

// optional timeline of one CTA (group 0, CTA 0), steps [kTraceStep0, kTraceStep0+kTraceSteps): SM clocks
constexpr int kTraceStep0 = 64, kTraceSteps = 16, kTraceSlots = 32;
__device__ __forceinline__ void trace(const TcParams& p, int t, int slot) {
    if (p.trace && blockIdx.x == 0 && t >= kTraceStep0 && t < kTraceStep0 + kTraceSteps)
        p.trace[(t - kTraceStep0) * kTraceSlots + slot] = clock64();
}

__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

// epilogue threads only: named barrier, then one thread releases the group counter
__device__ __forceinline__ void publish_arrive(unsigned int* ctr) {
    asm volatile("bar.sync 1, %0;" ::"n"(NEPI * 32) : "memory");
    if (threadIdx.x == 0)   // release at gpu scope: cumulative over the stores the barrier above ordered before it
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}

}  // namespace

// per-thread state of one fold set (a CTA serves up to kTcSets sets of <= 128 folds, software-pipelined: while one set's
// activations travel and its MMAs run, the epilogue warps work on the other set)
struct SetState {
    int nrows, fold0;            // live folds of this virtual group, its first fold (launch index)
    bool live;                   // my TMEM lane (row) carries a fold
    uint32_t fold, utt;          // Philox counter words of my fold
    size_t grow;                 // my row in the exchange buffers
    uint32_t tacc;               // TMEM address: my lane quarter, this set's column block
    unsigned int* ctrs;          // H1, H2, F1, F2 arrival counters of this virtual group
    const float4* cs;            // conditioning records of (this virtual group, my row, my unit pair), step 0
    float x, h1[2], h2[2], p3[2];
    float4 ca, cb, cc, cd;
};

__global__ void __launch_bounds__(NT, 1)
wrnn_loop_tc_kernel(const __grid_constant__ CUtensorMap tmH1, const __grid_constant__ CUtensorMap tmH2,
                    const __grid_constant__ CUtensorMap tmF1, const __grid_constant__ CUtensorMap tmF2, TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    Ctl* ctl = reinterpret_cast<Ctl*>(smem + kBars);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x / kTcCtas, cta = blockIdx.x % kTcCtas;       // group, CTA inside the group
    const int nsets = p.nsets;                                            // fold sets per group (1 or 2)
    const bool has_e = (p.mode == 0) || (cta == 0);                        // MOL: only CTA 0 of a group runs fc3
    const int NE = (p.mode == 0) ? 16 : 32;

    // ---- one-time setup ---------------------------------------------------------------------------------------
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.wimg + (size_t)cta * kWBytes);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        for (int i = tid; i < kWBytes / 16; i += NT) dst[i] = src[i];
        fence_proxy_async_smem();
    }
    if (tid == 0) {
        for (int i = 0; i < kMaxSlots; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], 1); }
        for (int i = 0; i < 4 * kTcSets; ++i) mbar_init(&ctl->accfull[i], 1);
        ctl->abort_local = 0;
        mbar_fence_init();
    }
    if (tid < 32) reinterpret_cast<float*>(smem + kBias)[tid] = (p.mode == 1 && tid < 30) ? p.bfc3[tid] : 0.f;
    if (warp == 0) tmem_alloc(&ctl->tmem, kTmemCols);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = ctl->tmem;
    // activation ring: kSlots x 16 KB
    // (measured: packing shorter slots so that the 16 KB MMA read of one slot overlaps the TMA target of the next is
    //  ~20 % slower per step than keeping the slots 16 KB apart, so the stride stays at the full tile)
    constexpr uint32_t slot_bytes = kTileBytes, nslots = kSlots;

    if (warp == NEPI) {
        // =================================== TMA producer ===================================================
        if (lane == 0) {
            const CUtensorMap* maps[4] = {&tmH1, &tmH2, &tmF1, &tmF2};
            for (int i = 0; i < 4; ++i) tma_prefetch_desc(maps[i]);
            uint32_t q = 0;
            const int nph = has_e ? 4 : 3;
            for (int t = 0; t < p.S; ++t) {
                for (int ph = 0; ph < nph; ++ph) {
                    for (int s = 0; s < nsets; ++s) {
                        const int vg = g * nsets + s;
                        const bool ok = wait_counter(p, ctl, p.counters + vg * 4 + ph, (unsigned int)kTcCtas * (unsigned int)(t + 1));
                        fence_proxy_async();
                        if (s == 0) trace(p, t, 12 + ph);
                        for (int kb = 0; kb < kNKB; ++kb, ++q) {
                            const uint32_t slot = q % nslots, round = q / nslots;
                            bool go = ok;
                            if (round > 0) go = wait_mbar(p, ctl, &ctl->empty[slot], (round - 1) & 1) && go;
                            if (go && !aborted(p, ctl)) {
                                mbar_arrive_expect_tx(&ctl->full[slot], (uint32_t)p.tile_bytes);
                                tma_load_2d(smem + kRing + slot * slot_bytes, maps[ph], &ctl->full[slot], kb * kKB, vg * 128);
                            }
                        }
                        if (s == 0) trace(p, t, 16 + ph);
                    }
                }
            }
        }
    } else if (warp == NEPI + 1) {
        // =================================== MMA issuer =====================================================
        if (lane == 0) {
            const uint32_t wofs[4] = {kWB, kWC, kWD, kWE};
            const uint32_t ncol[4] = {NB_, NC_, ND_, (uint32_t)NE};
            const uint32_t acc[4] = {kAccB, kAccC, kAccD, kAccE};
            uint32_t q = 0;
            const int nph = has_e ? 4 : 3;
            for (int t = 0; t < p.S; ++t) {
                for (int ph = 0; ph < nph; ++ph) {
                    const uint32_t idesc = umma_idesc_f16(128, (int)ncol[ph]);
                    for (int s = 0; s < nsets; ++s) {
                        const uint32_t dcol = tmem + (uint32_t)s * kSetCols + acc[ph];
                        for (int kb = 0; kb < kNKB; ++kb, ++q) {
                            const uint32_t slot = q % nslots, round = q / nslots;
                            if (s == 0 && ph == 0 && kb < 4) trace(p, t, 24 + kb);
                            const bool ok = wait_mbar(p, ctl, &ctl->full[slot], round & 1);
                            tcgen05_fence_after();
                            if (s == 0 && ph == 0 && kb < 4) trace(p, t, 20 + kb);
                            if (ok) {
                                const uint64_t ad = umma_desc_sw128(smem_u32(smem + kRing + slot * slot_bytes));
                                const uint64_t bd = umma_desc_sw128(smem_u32(smem + wofs[ph] + kb * ncol[ph] * 128));
                                if (kb == 0) umma_f16_c<false>(dcol, ad, bd, idesc); else umma_f16_c<true>(dcol, ad, bd, idesc);
                                umma_f16_c<true>(dcol, umma_desc_advance(ad, 32), umma_desc_advance(bd, 32), idesc);
                                umma_f16_c<true>(dcol, umma_desc_advance(ad, 64), umma_desc_advance(bd, 64), idesc);
                                umma_f16_c<true>(dcol, umma_desc_advance(ad, 96), umma_desc_advance(bd, 96), idesc);
                            }
                            umma_commit(&ctl->empty[slot]);
                        }
                        umma_commit(&ctl->accfull[s * 4 + ph]);
                    }
                }
            }
        }
    } else {
        // =================================== epilogue warps =================================================
        // thread = (fold row, unit pair up): TMEM lane = row, my units are 8*cta + 2*up + {0,1}
        const int row = (warp & 3) * 32 + lane, up = warp >> 2;
        const int j0 = cta * kTcUnits + 2 * up;                           // first of my two hidden units
        float v1[6], v2[6], v3[2], bh1[2], bh2[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int gt = 0; gt < 3; ++gt) { v1[gt * 2 + u] = p.v1[gt * kRnn + j0 + u]; v2[gt * 2 + u] = p.v2[gt * kRnn + j0 + u]; }
            v3[u] = p.v3[j0 + u]; bh1[u] = p.bhn1[j0 + u]; bh2[u] = p.bhn2[j0 + u];
        }
        const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
        const size_t cs_rec = (size_t)kTcCtas * 4 * 4;                      // float4 per (t,row) record = 1024
        SetState st[kTcSets];
#pragma unroll
        for (int s = 0; s < kTcSets; ++s) {
            SetState& S = st[s];
            const int vg = g * nsets + s;
            S.fold0 = vg * p.Mg;
            S.nrows = (s < nsets) ? max(0, min(p.Mg, p.B - S.fold0)) : 0;
            S.live = row < S.nrows;
            const FoldDesc fd = p.folds[S.live ? S.fold0 + row : 0];
            S.fold = (uint32_t)fd.fold; S.utt = (uint32_t)fd.utt;
            S.grow = (size_t)vg * 128 + row;
            S.tacc = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)s * kSetCols;
            S.ctrs = p.counters + vg * 4;
            S.cs = p.CS + (((size_t)vg * p.S) * p.Mg + row) * cs_rec + ((size_t)cta * 4 + up) * 4;
            S.x = 0.f; S.h1[0] = S.h1[1] = S.h2[0] = S.h2[1] = S.p3[0] = S.p3[1] = 0.f;
            S.ca = S.cb = S.cc = S.cd = make_float4(0.f, 0.f, 0.f, 0.f);
        }

        // ---- A: x_{t-1}, GRU1 for my 2 units, publish h1 ---------------------------------------------------
        auto stageA = [&](SetState& S, const int s, const int t) {
            if (S.live) {                    // conditioning of this step (issued before the wait on x)
                const float4* cs = S.cs + (size_t)t * p.Mg * cs_rec;
                S.ca = __ldcs(cs); S.cb = __ldcs(cs + 1); S.cc = __ldcs(cs + 2); S.cd = __ldcs(cs + 3);
            }
            if (tid == 0) trace(p, t, s == 0 ? 0 : 31);
            S.x = 0.f;
            if (t > 0 && S.live) wait_x(p, ctl, p.bX + S.grow, (uint32_t)t, S.x);
            if (tid == 0 && s == 0) trace(p, t, 1);
            float gh[8];
            if (t > 0) { tmem_ld8(S.tacc + kAccB + 16 * up + 8, gh); tmem_ld_wait(); }
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) gh[i] = 0.f;
            }
            const float c1r[2] = {S.ca.x, S.ca.y}, c1z[2] = {S.ca.z, S.ca.w}, c1n[2] = {S.cb.x, S.cb.y};
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float r = sigmoid_fast(fmaf(v1[0 + u], S.x, c1r[u]) + gh[0 + u]);
                const float z = sigmoid_fast(fmaf(v1[2 + u], S.x, c1z[u]) + gh[2 + u]);
                const float n = tanh_fast(fmaf(v1[4 + u], S.x, c1n[u]) + r * (gh[4 + u] + bh1[u]));
                S.h1[u] = (1.0f - z) * n + z * S.h1[u];
            }
            if (S.live) *reinterpret_cast<__half2*>(p.H1 + S.grow * kRnn + j0) = __floats2half2_rn(S.h1[0], S.h1[1]);
            tcgen05_fence_before();
            publish_arrive(S.ctrs + 0);
            if (tid == 0 && s == 0) trace(p, t, 2);
        };
        // ---- B: [W_ih2a h1 | W_fc1a h1 | gh1'] ; GRU2 ; publish h2 -------------------------------------------
        auto stageB = [&](SetState& S, const int s, const int t) {
            float pb[8], gh[8];
            wait_mbar(p, ctl, &ctl->accfull[s * 4 + 0], (uint32_t)t & 1u);
            tcgen05_fence_after();
            if (tid == 0 && s == 0) trace(p, t, 3);
            tmem_ld8(S.tacc + kAccB + 16 * up, pb);
            if (t > 0) tmem_ld8(S.tacc + kAccC + 8 * up, gh);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) gh[i] = 0.f;
            }
            tmem_ld_wait();
            if (tid == 0 && s == 0) trace(p, t, 28);
            const float c2r[2] = {S.cb.z, S.cb.w}, c2z[2] = {S.cc.x, S.cc.y}, c2n[2] = {S.cc.z, S.cc.w};
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float r = sigmoid_fast(pb[0 + u] + fmaf(v2[0 + u], S.x, c2r[u]) + gh[0 + u]);
                const float z = sigmoid_fast(pb[2 + u] + fmaf(v2[2 + u], S.x, c2z[u]) + gh[2 + u]);
                const float n = tanh_fast(pb[4 + u] + fmaf(v2[4 + u], S.x, c2n[u]) + r * (gh[4 + u] + bh2[u]));
                S.h2[u] = (1.0f - z) * n + z * S.h2[u];
                S.p3[u] = pb[6 + u];
            }
            if (S.live) *reinterpret_cast<__half2*>(p.H2 + S.grow * kRnn + j0) = __floats2half2_rn(S.h2[0], S.h2[1]);
            if (tid == 0 && s == 0) trace(p, t, 29);
            tcgen05_fence_before();
            publish_arrive(S.ctrs + 1);
            if (tid == 0 && s == 0) trace(p, t, 4);
        };
        // ---- C: [gh2' | W_fc1a h2] ; f1 ; publish ------------------------------------------------------------
        auto stageC = [&](SetState& S, const int s, const int t) {
            float pb[8];
            wait_mbar(p, ctl, &ctl->accfull[s * 4 + 1], (uint32_t)t & 1u);
            tcgen05_fence_after();
            if (tid == 0 && s == 0) trace(p, t, 5);
            tmem_ld8(S.tacc + kAccC + 8 * up, pb);
            tmem_ld_wait();
            const float f0 = fmaxf(S.p3[0] + pb[6] + fmaf(v3[0], S.x, S.cd.x), 0.f);
            const float f1 = fmaxf(S.p3[1] + pb[7] + fmaf(v3[1], S.x, S.cd.y), 0.f);
            if (S.live) *reinterpret_cast<__half2*>(p.F1 + S.grow * kRnn + j0) = __floats2half2_rn(f0, f1);
            tcgen05_fence_before();
            publish_arrive(S.ctrs + 2);
            if (tid == 0 && s == 0) trace(p, t, 6);
        };
        // ---- D: fc2 ; publish ----------------------------------------------------------------------------------
        auto stageD = [&](SetState& S, const int s, const int t) {
            wait_mbar(p, ctl, &ctl->accfull[s * 4 + 2], (uint32_t)t & 1u);
            tcgen05_fence_after();
            if (tid == 0 && s == 0) trace(p, t, 7);
            float d[4];
            tmem_ld4(S.tacc + kAccD + 2 * up, d);
            tmem_ld_wait();
            if (S.live) *reinterpret_cast<__half2*>(p.F2 + S.grow * kRnn + j0) = __floats2half2_rn(fmaxf(d[0] + S.cd.z, 0.f), fmaxf(d[1] + S.cd.w, 0.f));
            tcgen05_fence_before();
            publish_arrive(S.ctrs + 3);
            if (tid == 0 && s == 0) trace(p, t, 8);
        };
        // ---- E: fc3 + sampling ---------------------------------------------------------------------------------
        auto stageE = [&](SetState& S, const int s, const int t) {
            const uint32_t par = (uint32_t)t & 1u;
            if (p.mode == 1) {
                // MOL (vocoder/distribution.py:104-140): CTA 0 of the group has all 30 outputs of a fold in one TMEM
                // lane.  The four threads of a fold split the Gumbel draws (thread `up` owns Philox block `up`, i.e.
                // mixtures 4up..4up+3), meet through shared memory, and thread up==2 (which also holds the logistic
                // uniform, block 2 word 2) finishes the draw.
                if (cta == 0) {
                    wait_mbar(p, ctl, &ctl->accfull[s * 4 + 3], par);
                    tcgen05_fence_after();
                    if (tid == 0 && s == 0) trace(p, t, 9);
                    float lg[32];
                    tmem_ld8(S.tacc + kAccE + 0, lg); tmem_ld8(S.tacc + kAccE + 8, lg + 8);
                    tmem_ld8(S.tacc + kAccE + 16, lg + 16); tmem_ld8(S.tacc + kAccE + 24, lg + 24);
                    tmem_ld_wait();
                    const float* sbias = reinterpret_cast<const float*>(smem + kBias);
                    float2* scratch = reinterpret_cast<float2*>(smem + kMolScratch + s * kMolScratchBytes);
                    const uint4 r = philox4x32_10(make_uint4((uint32_t)t, S.fold, S.utt, (uint32_t)(up < 3 ? up : 2)), key);
                    float best = -INFINITY;
                    int kbest = 0;
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const int i = up * 4 + w;
                        if (i < 10) {
                            const float um = 1e-5f + u01(word_of(r, w)) * (1.0f - 2e-5f);
                            float li = 0.f;
#pragma unroll
                            for (int q = 0; q < 10; ++q) if (q == i) li = lg[q] + sbias[q];
                            const float sc = li - __logf(-__logf(um));
                            if (sc > best) { best = sc; kbest = i; }
                        }
                    }
                    scratch[row * 4 + up] = make_float2(best, __int_as_float(kbest));
                    asm volatile("bar.sync 2, %0;" ::"n"(NEPI * 32) : "memory");
                    if (up == 2 && S.live) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) {          // candidates of up = 0, 1 come first (lower indices win ties)
                            const float2 c = scratch[row * 4 + q];
                            if (c.x >= best && !(c.x == best && __float_as_int(c.y) > kbest)) { best = c.x; kbest = __float_as_int(c.y); }
                        }
                        float mean = 0.f, lsc = 0.f;
#pragma unroll
                        for (int i = 0; i < 10; ++i)
                            if (i == kbest) { mean = lg[10 + i] + sbias[10 + i]; lsc = lg[20 + i] + sbias[20 + i]; }
                        lsc = fmaxf(lsc, -32.23619130191664f);
                        const float ul = 1e-5f + u01(r.z) * (1.0f - 2e-5f);
                        float xs = mean + __expf(lsc) * (__logf(ul) - __logf(1.0f - ul));
                        xs = fminf(fmaxf(xs, -1.0f), 1.0f);
                        p.samples[(size_t)(S.fold0 + row) * p.S + t] = xs;
                        const float fed = p.forced ? p.forced[(size_t)(S.fold0 + row) * p.S + t] : xs;
                        ll_store(p.bX + S.grow, fed, (uint32_t)t + 1u);
                        if (p.logits_out)
                            for (int i = 0; i < 30; ++i) p.logits_out[((size_t)(S.fold0 + row) * p.S + t) * 30 + i] = lg[i] + sbias[i];
                    }
                    tcgen05_fence_before();
                    if (tid == 0 && s == 0) trace(p, t, 10);
                }
            } else {
                // RAW: my CTA's classes of every fold -> exchange words; then one warp per assigned fold samples
                wait_mbar(p, ctl, &ctl->accfull[s * 4 + 3], par);
                tcgen05_fence_after();
                float d[4];
                const int cpu = p.C / (kTcCtas * 4);                      // classes per (CTA, up): 2 (C=512) or 4 (C=1024)
                tmem_ld4(S.tacc + kAccE + cpu * up, d);
                tmem_ld_wait();
                if (S.live) {
                    for (int i = 0; i < cpu; ++i) {
                        const int cls = cta * (cpu * 4) + cpu * up + i;
                        const float v = d[i] + p.bfc3[cls];
                        ll_store(p.bLG + S.grow * p.Cpad + cls, v, (uint32_t)t + 1u);
                        if (p.logits_out) p.logits_out[((size_t)(S.fold0 + row) * p.S + t) * p.C + cls] = v;
                    }
                }
                tcgen05_fence_before();
                const int srow = cta + kTcCtas * warp;                    // warps 0,1 sample rows cta, cta+64
                if (warp < 2 && srow < S.nrows && !aborted(p, ctl)) {
                    const FoldDesc sfd = p.folds[S.fold0 + srow];
                    const size_t sgrow = S.grow - row + srow;
                    const unsigned long long* lrow = p.bLG + sgrow * p.Cpad;
                    const uint4 r = philox4x32_10(make_uint4((uint32_t)t, (uint32_t)sfd.fold, (uint32_t)sfd.utt, 0u), key);
                    const float u = u01(r.x);
                    int k = (p.C == 512) ? sample_raw_warp<16>(lrow, (uint32_t)t + 1u, u, p.abort_flag, g_tc_deadline)
                                         : sample_raw_warp<32>(lrow, (uint32_t)t + 1u, u, p.abort_flag, g_tc_deadline);
                    if (k < 0) { raise_abort(p, ctl); k = 0; }
                    if (lane == 0) {
                        const float xs = 2.0f * (float)k / ((float)p.C - 1.0f) - 1.0f;
                        p.samples[(size_t)(S.fold0 + srow) * p.S + t] = xs;
                        const float fed = p.forced ? p.forced[(size_t)(S.fold0 + srow) * p.S + t] : xs;
                        ll_store(p.bX + sgrow, fed, (uint32_t)t + 1u);
                    }
                }
            }
        };

        const bool two = nsets > 1;
        for (int t = 0; t < p.S; ++t) {
            stageA(st[0], 0, t); if (two) stageA(st[1], 1, t);
            stageB(st[0], 0, t); if (two) stageB(st[1], 1, t);
            stageC(st[0], 0, t); if (two) stageC(st[1], 1, t);
            stageD(st[0], 0, t); if (two) stageD(st[1], 1, t);
            stageE(st[0], 0, t); if (two) stageE(st[1], 1, t);
            if (blockIdx.x == 0 && tid == 0 && (t % 100) == 0 && p.progress) {
                *reinterpret_cast<volatile int*>(p.progress) = t;
                __threadfence_system();
            }
        }
    }
    // ---- teardown ---------------------------------------------------------------------------------------------
    if (aborted(p, ctl)) __nanosleep(200000);      // let any TMA still in flight land before the CTA goes away
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

cudaError_t set_tc_deadline(long long cycles) { return cudaMemcpyToSymbol(g_tc_deadline, &cycles, sizeof(cycles)); }
size_t loop_tc_weight_image_bytes() { return kWBytes; }

cudaError_t launch_loop_tc(const TcParams& p, const void* tmaps /* 4 x CUtensorMap */, cudaStream_t stream) {
    cudaError_t err = cudaFuncSetAttribute(wrnn_loop_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes + 1024);
    if (err != cudaSuccess) return err;
    const CUtensorMap* m = reinterpret_cast<const CUtensorMap*>(tmaps);
    TcParams pp = p;
    void* args[] = {(void*)&m[0], (void*)&m[1], (void*)&m[2], (void*)&m[3], &pp};
    return cudaLaunchCooperativeKernel((const void*)wrnn_loop_tc_kernel, dim3(kTcGroups * kTcCtas), dim3(NT), args,
                                       kSmemBytes + 1024, stream);
}

}  // namespace wrnn
