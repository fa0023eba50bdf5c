// loop_tc.cu -- the autoregressive sample loop on the 5th-gen tensor cores (fp16 operands, fp32 accumulate,
// fp32 recurrent state): the batched-fold path (BASELINE configs 1, 3, 5).  DESIGN.md section 4.2 has the full account.
//
// Same five-exchange step as loop_f32.cu, restructured for tcgen05; one persistent launch with four kinds of CTA:
//  * unit-owning CTAs, 2 groups x 64 (one CTA per SM).  A group owns one to four SETS of up to 128 folds (one fold per
//    TMEM lane); each of its CTAs owns 8 hidden units of every layer.  The sets are software-pipelined through the CTA:
//    while one set's activations travel through L2 and its MMAs run, the epilogue warps do another set's math.  The CTA's
//    weight rows -- [W_ih2a|W_fc1a|W_hh1] (stage B), [W_hh2|W_fc1a] (C), fc2 (D) -- stay in shared memory as K-major
//    SWIZZLE_128B tiles (fp16, 112 KB) for the whole sequence and are the B operand of tcgen05.mma (N = 64/32/16).
//    Above 512 folds the CTAs work as PAIRS (2-CTA clusters, tcgen05 cta_group::2): one M=256 MMA covers a set of each
//    CTA against both CTAs' weight rows, which halves the TMA operations, MMAs and L2->SM bytes per fold.
//  * activations h1, h2, f1, f2 travel as fp16 rows [fold][512] through L2.  A stage = publish my columns, arrive on the
//    set's counter (release, by the publisher warp), the producer thread acquires the counter, then TMA-loads the
//    set's [128 x 512] activation matrix in four operations of two swizzled [128 x 64] k-blocks (a 3-D tensor map) through
//    a 3-slot mbarrier ring; one thread issues 32 tcgen05.mma per stage into TMEM; 16 epilogue warps (thread = fold x unit
//    pair) read the accumulator with tcgen05.ld and do the GRU / ReLU math in fp32.
//  * sampler CTAs: fc3 and the draw never sit on a unit-owning CTA.  MOL: one CTA per group (per group and rank with
//    pairs), N=32.  RAW with 512 classes: four CTAs per group with a quarter of fc3 each (as pairs: N=256 per MMA), online
//    softmax per TMEM lane, {max, sum} partials exchanged as tagged words, inverse-CDF scan in the owning quarter.
//    (Other class counts: classes spread over the unit-owning CTAs, raw_stage_e.)
//  * expander CTAs on the remaining SMs produce the per-sample conditioning records (cond_expand.cuh) a few 32-step
//    chunks ahead of the loop, into a ring with produced / consumed counters.
//  * "soft abort": a wait that passes its deadline raises a flag; from then on every wait returns at once, so
//    all warps still walk the same barriers and the kernel ends cleanly (never a hung GPU).
#include <cstdio>
#include <cstdlib>
#include "engine_internal.h"
#include "sampling.cuh"
#include "tc_common.cuh"
#include "cond_expand.cuh"

namespace wrnn {

__device__ long long g_tc_deadline = 1500000000LL;

namespace {
using namespace tc;

constexpr int NEPI = 16;                    // epilogue warps
constexpr int NT = (NEPI + 3) * 32;         // + TMA producer warp + MMA warp + publisher warp
constexpr int kSlots = 4, kMaxSlots = 6, kKB = 64, kNKB = kRnn / kKB;   // ring: 4 x 16 KB (6 where the fc3 rows are not resident)
constexpr int kTileBytes = 128 * 128;       // one [128 rows x 64 fp16] activation tile
constexpr int NB_ = 64, NC_ = 32, ND_ = 16; // MMA N per stage (E: 16 RAW / 32 MOL)
constexpr int kWB = 0, kWC = kWB + NB_ * 128 * kNKB, kWD = kWC + NC_ * 128 * kNKB, kWE = kWD + ND_ * 128 * kNKB;
constexpr int kWBytes = kWE + 32 * 128 * kNKB;          // 147456
constexpr int kRing = kWBytes;                          // 4 x 16 KB
constexpr int kBars = kRing + kSlots * kTileBytes;      // mbarriers + misc
constexpr int kMolScratchBytes = 128 * 4 * 8;           // [128 rows][4] {score, index} for the cooperative MOL draw, per fold set
constexpr int kMolScratch = 0;                           // sampler CTA only: the GRU weight area is unused there
constexpr int kMolRing = kTcSets * kMolScratchBytes;     // MOL sampler CTA: 4 ring slots between the scratch and its fc3 rows at kRing
constexpr int kBias = kBars + 256;                       // fc3 bias (MOL)
constexpr int kConst = kBias + 128;                      // per-unit constants of the CTA's 8 (pairs: 16) units (72 floats each)
constexpr int kXBuf = kConst + 1024;                     // previous samples of the CTA's sets, [kTcSets][128] floats
#ifndef WRNN_X_GATHER
#define WRNN_X_GATHER 0     // 1: one warp polls a set's sample words for the CTA (halves the cross-CTA skew of stage A, tools/tc_skew.py,
#endif                      //    but the step is not shorter: 40.0 vs 39.3 us at 1024 folds, 26.0 vs 25.7 us at 213)
// tcgen05.mma always reads 128 rows (16 KB) from a slot base; with slots shorter than that the last slot reads up to
// 8 KB past the ring (into the control words: harmless garbage rows), so the allocation must cover ring + 72 KB
constexpr int kSmemBytes = (kXBuf + kTcSets * 512 > kRing + 73728) ? (kXBuf + kTcSets * 512) : (kRing + 73728);
// RAW with 512 classes: fc3 + the draw run on kRawQ dedicated sampler CTAs per group, 128 classes (one N=128 MMA tile,
// 128 KB of resident fc3 rows at offset 0) each; their fc3 bias slice sits in the gap below the ring
constexpr int kRawQ = 4, kRawQCols = 128, kRawBias = kRawQCols * 128 * kNKB;
static_assert(kRawBias + kRawQCols * 4 <= kRing, "RAW sampler bias overlaps the ring");
static_assert(kTcSets <= NEPI / 4, "RAW sampler: one group of four epilogue warps per fold set");
static_assert(kMolRing % 1024 == 0 && kMolRing + 4 * kTcKbPerOp * kTileBytes <= kRing && kRing + 32 * 128 * kNKB <= kBars, "MOL sampler layout");
// TMEM columns
constexpr int kAccB = 0, kAccC = 64, kAccD = 96, kAccE = 112, kSetCols = 128, kTmemCols = 512;   // (sampler CTA: fc3 accumulator at column 0 of the set's block)   // per fold set

struct Ctl {
    uint64_t full[kMaxSlots];
    uint64_t empty[kMaxSlots];
    uint64_t accfull[4 * kTcSets];
    uint32_t tmem;
    int abort_local;
    int exp_released;      // expander CTAs: items whose completion barrier the releaser warp has left
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
#ifdef WRNN_XSLEEP
        __nanosleep(WRNN_XSLEEP);          // 64 K threads poll a handful of L2 lines: back off
#endif
        if (((++spins) & 255) == 0 && spin_check(p, c, t0)) { v = 0.f; return false; }
    }
}

// optional timeline of one CTA (group 0, CTA 0), steps [kTraceStep0, kTraceStep0+kTraceSteps): SM clocks
constexpr int kTraceStep0 = 64, kTraceSteps = 16, kTraceSlots = 192;
__device__ __forceinline__ void trace(const TcParams& p, int t, int slot) {
    if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && t >= kTraceStep0 && t < kTraceStep0 + kTraceSteps)
        p.trace[(t - kTraceStep0) * kTraceSlots + slot] = clock64();
}

// cross-CTA skew: every unit-owning CTA stamps %globaltimer (ns, common to all SMs) at a few events of step kTraceStep0 + 8
// into the area behind CTA 0's timeline: [CTA][16]: 4 s + {0 x arrived, 1 A done, 2 B accumulator ready, 3 D done}
constexpr int kXSlots = 16;
__device__ __forceinline__ void xtrace(const TcParams& p, int t, int s, int k) {
    if (p.trace && threadIdx.x == 0 && t == kTraceStep0 + 8 && s < 4) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        p.trace[kTraceSteps * kTraceSlots + blockIdx.x * kXSlots + 4 * s + k] = (long long)ns;
    }
}

// epilogue timeline of fold sets 0 and 1: slot 32 + 16 s + {0 A enter, 1 x, 2 A done, 3 B enter, 4 B acc, 5 B done, 6 C enter, ...}
__device__ __forceinline__ void etrace(const TcParams& p, int t, int s, int k) {
    if (p.trace && threadIdx.x == 0 && s < 2) trace(p, t, 32 + 16 * s + k);
}

__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

// Publishing a stage: the epilogue threads store their activations and ARRIVE on the set's named barrier without
// waiting; the publisher warp completes that barrier and releases the group counter.  The release fence (a round trip
// to L2 for the stores, ~0.7 us) is paid by the publisher, so the epilogue warps are already working on the next set.
// One barrier per fold set is enough: a set's next stage cannot start before its previous publish was consumed.
constexpr int kPubBar0 = 3;
__device__ __forceinline__ void publish_arrive(int s, bool inline_release, unsigned int* ctr) {
    if (inline_release) {              // the epilogue pays for the release itself (shorter chain, busier epilogue)
        asm volatile("bar.sync 1, %0;" ::"n"(NEPI * 32) : "memory");
        if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        return;
    }
    asm volatile("bar.arrive %0, %1;" ::"r"(kPubBar0 + s), "n"((NEPI + 1) * 32) : "memory");
}
__device__ __forceinline__ void publisher_release(int s, unsigned int* ctr) {
    asm volatile("bar.sync %0, %1;" ::"r"(kPubBar0 + s), "n"((NEPI + 1) * 32) : "memory");
    if ((threadIdx.x & 31) == 0)   // release at gpu scope: cumulative over the stores the barrier above ordered before it
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}

// Station E, RAW: my CTA's classes of every fold -> exchange words; then one warp per assigned fold samples.
// Out of line on purpose: its registers (and spills) stay out of the stages that sit on every step's chain.
__device__ __noinline__ void raw_stage_e(const TcParams& p, Ctl* ctl, uint32_t tacc, int cta, int fold0, int nrows, size_t grow, int t,
                                         uint2 key) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = (warp & 3) * 32 + lane, up = warp >> 2;
    const bool live = row < nrows;
    float d[4];
    const int cpu = p.C / (kTcCtas * 4);                      // classes per (CTA, up): 2 (C=512) or 4 (C=1024)
    tmem_ld4(tacc + kAccE + cpu * up, d);
    tmem_ld_wait();
    if (live) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i < cpu) {
                const int cls = cta * (cpu * 4) + cpu * up + i;
                const float v = d[i] + p.bfc3[cls];
                ll_store(p.bLG + grow * p.Cpad + cls, v, (uint32_t)t + 1u);
                if (p.logits_out) p.logits_out[((size_t)(fold0 + row) * p.S + t) * p.C + cls] = v;
            }
        }
    }
    const int srow = cta + kTcCtas * warp;                    // warps 0,1 sample rows cta, cta+64
    if (warp < 2 && srow < nrows && !aborted(p, ctl)) {
        const FoldDesc sfd = p.folds[fold0 + srow];
        const size_t sgrow = grow - row + srow;
        const unsigned long long* lrow = p.bLG + sgrow * p.Cpad;
        const uint4 r = philox4x32_10(make_uint4((uint32_t)t, (uint32_t)sfd.fold, (uint32_t)sfd.utt, 0u), key);
        const float u = u01(r.x);
        int k = (p.C == 512) ? sample_raw_warp<16>(lrow, (uint32_t)t + 1u, u, p.abort_flag, g_tc_deadline)
                             : sample_raw_warp<32>(lrow, (uint32_t)t + 1u, u, p.abort_flag, g_tc_deadline);
        if (k < 0) { raise_abort(p, ctl); k = 0; }
        if (lane == 0) {
            const float xs = 2.0f * (float)k / ((float)p.C - 1.0f) - 1.0f;
            p.samples[(size_t)(fold0 + srow) * p.S + t] = xs;
            const float fed = p.forced ? p.forced[(size_t)(fold0 + srow) * p.S + t] : xs;
            ll_store(p.bX + sgrow, fed, (uint32_t)t + 1u);
        }
    }
}

// Station E on a RAW sampler CTA (512 classes): the CTA holds fc3 rows [128 qd, 128 qd + 128) and has their products for
// every fold of the set in TMEM (lane = fold, 128 columns).  Thread = fold: softmax partials {max, sum} of my quarter go
// to the other three sampler CTAs of the group as tagged words (double-buffered by step parity), every CTA forms the
// same normaliser and threshold u * Z from the four pairs, and the CTA whose quarter contains the threshold scans its
// 128 classes for the first k with cdf[k] >= u (rule: oracle sample_raw; fatchord_version.py:224-230).
__device__ __noinline__ void raw_sampler_e(const TcParams& p, Ctl* ctl, const float* sbias, uint32_t tacc, int qd, int fold0, int nrows,
                                           size_t grow, int row, int t, uint32_t fold, uint32_t utt, uint2 key) {
    const bool live = row < nrows;
    const uint32_t tag = (uint32_t)t + 1u;
    float m = -INFINITY, ssum = 0.f;
#pragma unroll 1
    for (int c = 0; c < kRawQCols; c += 32) {          // online softmax partials, 32 classes at a time
        float l[32];
        tmem_ld8(tacc + c, l); tmem_ld8(tacc + c + 8, l + 8); tmem_ld8(tacc + c + 16, l + 16); tmem_ld8(tacc + c + 24, l + 24);
        tmem_ld_wait();
        float cm = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) { l[i] += sbias[c + i]; cm = fmaxf(cm, l[i]); }
        if (cm > m) { ssum *= __expf(m - cm); m = cm; }
#pragma unroll
        for (int i = 0; i < 32; ++i) ssum += __expf(l[i] - m);
        if (p.logits_out && live) {
            float* lo = p.logits_out + ((size_t)(fold0 + row) * p.S + t) * p.C + qd * kRawQCols + c;
#pragma unroll
            for (int i = 0; i < 32; ++i) lo[i] = l[i];
        }
    }
    unsigned long long* xw = p.bLG + ((size_t)(t & 1) * ((size_t)kTcGroups * kTcSets * 128) + grow) * (2 * kRawQ);
    if (live) { ll_store(xw + 2 * qd, m, tag); ll_store(xw + 2 * qd + 1, ssum, tag); }
    float mq[kRawQ], zq[kRawQ];
    bool ok = true;
#pragma unroll
    for (int q = 0; q < kRawQ; ++q) {
        mq[q] = m; zq[q] = ssum;
        if (q != qd && live && ok) {
            long long t0 = 0;
            int spins = 0;
            while (true) {
                unsigned long long a, b;
                ll_load2(xw + 2 * q, a, b);
                if (ll_tag(a) == tag && ll_tag(b) == tag) { mq[q] = ll_val(a); zq[q] = ll_val(b); break; }
                if (((++spins) & 255) == 0 && spin_check(p, ctl, t0)) { ok = false; break; }
            }
        }
    }
    const float M = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3]));
#pragma unroll
    for (int q = 0; q < kRawQ; ++q) zq[q] *= __expf(mq[q] - M);
    const float c0 = zq[0], c1 = c0 + zq[1], c2 = c1 + zq[2], Z = c2 + zq[3];
    const uint4 r = philox4x32_10(make_uint4((uint32_t)t, fold, utt, 0u), key);
    const float thr = u01(r.x) * Z;
    const int qs = thr <= c0 ? 0 : (thr <= c1 ? 1 : (thr <= c2 ? 2 : 3));
    const bool mine = live && ok && qs == qd;
    if (!__any_sync(0xffffffffu, mine)) return;         // (tcgen05.ld is warp-wide: the whole warp scans or none of it)
    float cum = qs == 0 ? 0.f : (qs == 1 ? c0 : (qs == 2 ? c1 : c2));
    int k = -1;
#pragma unroll 1
    for (int c = 0; c < kRawQCols; c += 32) {
        float l[32];
        tmem_ld8(tacc + c, l); tmem_ld8(tacc + c + 8, l + 8); tmem_ld8(tacc + c + 16, l + 16); tmem_ld8(tacc + c + 24, l + 24);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            cum += __expf(l[i] + sbias[c + i] - M);
            if (k < 0 && cum >= thr) k = c + i;
        }
        if (__all_sync(0xffffffffu, k >= 0 || !mine)) break;
    }
    if (mine) {
        if (k < 0) k = kRawQCols - 1;                   // rounding left the quarter's own sum short of the threshold
        const float xs = 2.0f * (float)(qd * kRawQCols + k) / ((float)p.C - 1.0f) - 1.0f;
        p.samples[(size_t)(fold0 + row) * p.S + t] = xs;
        const float fed = p.forced ? p.forced[(size_t)(fold0 + row) * p.S + t] : xs;
        ll_store(p.bX + grow, fed, tag);
    }
}

}  // namespace

// The schedule.  Per step a fold set passes five stations: A (GRU1 on the new sample), B, C, D (an MMA stage and its
// epilogue each) and E (fc3 + the draw), each about one exchange + one K=512 reduction long.  Time is cut into slots; in
// slot k, set s is at station (k - off(s)) mod 5 of step (k - off(s)) / 5, with off(s) = s * skew: with skew > 0 the sets sit at
// DIFFERENT stations at any moment, so every in-order resource (TMA/MMA queue, epilogue warps, publisher, sampler CTA)
// sees at most one job per set per slot and never queues one set's stage behind the same stage of all the others.
// Every role enumerates the same (slot, set) sequence and picks the stations it serves.
__device__ __forceinline__ bool job_of(int k, int s, int skew, int S, int& t, int& stn) {
    const int idx = k - s * skew;
    if (idx < 0 || idx >= 5 * S) return false;
    t = idx / 5;
    stn = idx - 5 * t;
    return true;
}

// per-thread recurrent state of one fold set (a CTA serves several sets of <= 128 folds, software-pipelined: while one
// set's activations travel and its MMAs run, the epilogue warps work on the other sets).  H = unit blocks of 8 hidden
// units the CTA's accumulator rows cover: 1, or 2 for a CTA pair (its own and its partner's).  The whole conditioning
// record is fetched in stage A (fetching it stage by stage needs fewer registers but measured 5-8 % slower).
template <int H>
struct SetState {
    uint32_t fold, utt;          // Philox counter words of my fold
    float x, h1[2 * H], h2[2 * H], p3[2 * H];
    float2 cr[H]; float4 cz[H]; float4 c34[H];
};

// NSETS = fold sets the epilogue warps of a unit-owning CTA serve.
// PAIR  = the unit-owning CTAs work as CTA pairs (2-CTA clusters, tcgen05 cta_group::2): one MMA of M = 256 covers a set of
//         the even CTA (A rows 0..127) and a set of the odd CTA (rows 128..255) against BOTH CTAs' weight rows (N doubles),
//         so every activation tile a CTA pulls from L2 is used for 16 hidden units instead of 8: half the TMA operations,
//         MMAs and L2->SM bytes per fold.  A group then has 2*NSETS sets; set 2s+r belongs to the rank-r CTAs (32 CTAs
//         publish it and arrive on its counters).  Needs dedicated sampler CTAs (MOL, RAW-512).
template <int NSETS, bool PAIR>
__global__ void __launch_bounds__(NT, 1)
wrnn_loop_tc_kernel(const __grid_constant__ CUtensorMap tmH1, const __grid_constant__ CUtensorMap tmH2,
                    const __grid_constant__ CUtensorMap tmF1, const __grid_constant__ CUtensorMap tmF2, const __grid_constant__ TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    Ctl* ctl = reinterpret_cast<Ctl*>(smem + kBars);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool mol = p.mode == 1;
    const int nmain = kTcGroups * kTcCtas;
    const bool rawq = !mol && p.raw_samplers != 0;                        // RAW, 512 classes: kRawQ sampler CTAs per group
    const int spg = mol ? (PAIR ? 2 : 1) : (rawq ? kRawQ : 0);            // sampler CTAs per group (MOL pairs: one per rank)
    const bool has_samplers = spg != 0;
    const int sidx = (int)blockIdx.x - nmain;                             // blocks past the groups: samplers, then expanders
    const bool sampler = sidx >= 0;                                       // (fc3 + the draw)
    const int g = sampler ? (spg ? sidx / spg : sidx) : (int)blockIdx.x / kTcCtas;   // group
    const int qd = (sampler && rawq) ? sidx % kRawQ : 0;                  // RAW sampler: my quarter of the classes
    const int cta = sampler ? 0 : (int)blockIdx.x % kTcCtas;              // unit-owning CTA inside the group
    const bool expander = sampler && sidx >= kTcGroups * spg;             // blocks past the loop's CTAs expand the conditioning
    const bool idle = expander;
    constexpr int H = PAIR ? 2 : 1;                                       // unit blocks per epilogue thread
    constexpr int GS = PAIR ? 2 * NSETS : NSETS;                          // fold sets per group
    const bool pairu = PAIR && !sampler;                                  // unit-owning CTA, half of a CTA pair
    const bool pairs = PAIR && rawq && sampler && !expander;              // RAW sampler CTA, half of a pair: 256 classes per pair, N = 256
    const bool paired = pairu || pairs;
    const bool ranked = PAIR && !expander;                                // works on the group's sets 2s + rank only (else: on all)
    const int rank = pairu ? (cta & 1) : (ranked ? (sidx & 1) : 0);       // (MOL pairs: sampler CTA r of a group serves the rank-r sets)
    const bool leader = !paired || rank == 0;                             // issues the pair's MMAs, owns the `full` barriers
    const int ctab = pairu ? (cta & ~1) : cta;                            // first unit block of my accumulator columns
    const int nsets = (sampler && !ranked) ? GS : NSETS;                  // sets this CTA works on
    const unsigned int arrivals = PAIR ? kTcCtas / 2 : kTcCtas;           // CTAs that publish a set's activations
    // TMEM columns of a unit-owning CTA, per set: [B | C | D | (E)]; a pair's accumulators are twice as wide
    constexpr uint32_t accB = 0, accC = PAIR ? 128 : 64, accD = PAIR ? 192 : 96, setColsU = PAIR ? 256 : 128;
    const uint32_t set_cols = sampler ? (pairs ? 256u : 128u) : setColsU;
    static_assert(NSETS * (PAIR ? 256 : 128) <= kTmemCols && GS <= kTcSets, "TMEM columns");
#define VG_OF(s) (g * GS + (ranked ? 2 * (s) + rank : (s)))
    const int NE = mol ? 32 : ((rawq && sampler) ? kRawQCols : 16);
    const int ph0 = sampler ? 3 : 0, ph1 = (has_samplers && !sampler) ? 3 : 4;   // stages whose MMAs this CTA runs
    const int skew = (p.flags >> 4) & 7;                                  // stations set s runs behind set s-1 (0: all sets in phase)
    const int nslot_total = 5 * p.S + skew * (nsets - 1);                  // slots of the schedule (job_of)
    const bool deep = has_samplers && !sampler && !(p.flags & 2);
    // ring of 32 KB slots (two k-blocks each): 3 where the fc3 rows are not resident; the MOL sampler CTA holds only 32 KB of
    // weights (kept at kRing) and its draw scratch (16 KB at 0), so its ring is the 4 slots in between: all four operations
    // of a job are in flight at once (the sampler's stage is on every step's chain)
    const bool molsamp = mol && sampler;
    const uint32_t ring0 = molsamp ? (uint32_t)kMolRing : (deep ? kWE : kRing), nslots = molsamp ? 4u : (deep ? 3u : 2u);
    constexpr uint32_t slot_bytes = kTcKbPerOp * kTileBytes;
    const uint32_t kb_bytes = (uint32_t)p.tile_bytes;                         // the second k-block of a slot starts here
    // (measured: packing shorter slots so that the 16 KB MMA read of one slot overlaps the TMA target of the next is
    //  ~20 % slower per step than keeping the slots 16 KB apart, so the stride stays at the full tile)

    // ---- one-time setup ---------------------------------------------------------------------------------------
    if (!idle) {
        const bool rs = rawq && sampler;                // RAW sampler: its quarter of fc3 (N = 128 tiles) at offset 0
        const int w0 = rs ? 0 : (sampler ? kWE : 0), w1 = rs ? kRawBias : ((has_samplers && !sampler) ? kWE : kWBytes);
        const uint4* src = reinterpret_cast<const uint4*>(rs ? p.wimg_s + (size_t)qd * kRawBias : p.wimg + (size_t)cta * kWBytes + w0);
        uint4* dst = reinterpret_cast<uint4*>(smem + (molsamp ? kRing : w0));
        for (int i = tid; i < (w1 - w0) / 16; i += NT) dst[i] = src[i];
        if (rs && tid < 2 * kRawQCols)                   // fc3 bias of my quarter (pairs: of the pair's two quarters)
            reinterpret_cast<float*>(smem + kRawBias)[tid] = pairs ? p.bfc3[(qd & ~1) * kRawQCols + tid] : (tid < kRawQCols ? p.bfc3[qd * kRawQCols + tid] : 0.f);
        fence_proxy_async_smem();
    }
    if (tid == 0) {
        for (int i = 0; i < kMaxSlots; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], 1); }
        for (int i = 0; i < 4 * kTcSets; ++i) mbar_init(&ctl->accfull[i], 1);
        ctl->abort_local = 0;
        ctl->exp_released = 0;
        mbar_fence_init();
    }
    if (tid < 32) reinterpret_cast<float*>(smem + kBias)[tid] = (mol && tid < 30) ? p.bfc3[tid] : 0.f;
    if (tid < 72 * H) {                 // per unit block: [unit pair 4][v1 r,z,n x2 | v2 r,z,n x2 | v3 x2 | b_hn1 x2 | b_hn2 x2]
        const int hb = tid / 72, r72 = tid % 72;
        const int up_ = r72 / 18, i = r72 % 18, u = i & 1, j = (ctab + hb) * kTcUnits + 2 * up_ + u;
        float v;
        if (i < 6) v = p.v1[(i >> 1) * kRnn + j];
        else if (i < 12) v = p.v2[((i - 6) >> 1) * kRnn + j];
        else if (i < 14) v = p.v3[j];
        else if (i < 16) v = p.bhn1[j];
        else v = p.bhn2[j];
        reinterpret_cast<float*>(smem + kConst)[tid] = v;
    }
    if (expander && p.cs_done)                      // interpolation weights [200][kTaps]: broadcast reads from shared memory
        for (int i = tid; i < kHop * kTaps; i += NT) reinterpret_cast<float*>(smem)[i] = p.coef[i];
    if (warp == 0 && !idle) {
        if (paired) tmem_alloc_pair(&ctl->tmem, kTmemCols); else tmem_alloc(&ctl->tmem, kTmemCols);
    }
    tcgen05_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();        // (pair: the partner's barriers are initialised before any TMA completes on them)
    tcgen05_fence_after();
    const uint32_t tmem = ctl->tmem;

    if (expander) {
        // =================================== conditioning expander =============================================
        // work item = (16-step chunk, fold), chunk-major, so chunks complete in the order the loop consumes them.  512 worker
        // threads (thread = hidden unit) produce the item's records and ARRIVE on a named barrier; the releaser warp completes
        // it and bumps the chunk's counter with a release (cumulative over the workers' stores through the barrier), so the
        // workers never wait for their stores to drain.  Two barriers alternate by item; exp_released keeps a worker from
        // arriving at a barrier whose previous phase the releaser has not left yet.
        const int e_idx = sidx - kTcGroups * spg;
        if (p.cs_done && warp <= NEPI) {
            const int nchunks = (p.S + kExpandSteps - 1) / kExpandSteps;
            const long long nitems = (long long)nchunks * p.B;
            const int ring_chunks = p.cs_steps / kExpandSteps;
            const float* coef_s = reinterpret_cast<const float*>(smem);
            volatile int* released = reinterpret_cast<volatile int*>(&ctl->exp_released);
            int waited = -1, seq = 0;
            for (long long it = e_idx; it < nitems; it += p.n_expanders, ++seq) {
                const int c = (int)(it / p.B), b = (int)(it - (long long)c * p.B);
                const int bar_id = 9 + (seq & 1);
                if (warp == NEPI) {                             // ---- releaser
                    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"((NEPI + 1) * 32) : "memory");
                    if (lane == 0) {
                        *released = seq + 1;
                        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.cs_done + c) : "memory");
                    }
                    continue;
                }
                if (c >= ring_chunks && c != waited) {          // CS is a ring: chunk c overwrites chunk c - ring_chunks, which every
                    waited = c;                                 // unit-owning CTA must have finished reading
                    if (tid == 0) wait_counter(p, ctl, p.cs_consumed + (c - ring_chunks), (unsigned int)nmain);
                    asm volatile("bar.sync 8, %0;" ::"n"(NEPI * 32) : "memory");
                }
                const FoldDesc fd = p.folds[b];
                // (tried: prefetch.global.L2 of the next item's table rows -- no gain; 32-step chunks instead of 16: 5 % per step)
                expand_cond_item_regs(p.TA1, p.TA2, p.TQ1, p.TQ2, coef_s, fd, b, c * kExpandSteps, min(p.S, (c + 1) * kExpandSteps), p.cs_steps, p.Mg,
                                      p.CSw, tid);
                if (seq >= 2) {
                    if (lane == 0) {
                        long long t0 = 0;
                        int spins = 0;
                        while (*released < seq - 1)
                            if (((++spins) & 1023) == 0 && spin_check(p, ctl, t0)) break;
                    }
                    __syncwarp();
                }
                asm volatile("bar.arrive %0, %1;" ::"r"(bar_id), "n"((NEPI + 1) * 32) : "memory");
            }
        }
    } else
    if (warp == NEPI) {
        // =================================== TMA producer ===================================================
        // (pair: both CTAs load their own set's rows; the bytes of both complete on the even CTA's `full` barrier)
        // The whole warp walks the schedule and polls; one ELECTED lane issues (tc_common.cuh: elect_one).
        {
            const CUtensorMap* maps[4] = {&tmH1, &tmH2, &tmF1, &tmF2};
            if (lane == 0) for (int i = 0; i < 4; ++i) tma_prefetch_desc(maps[i]);
            uint32_t q = 0;
            for (int k = 0; k < nslot_total; ++k) {
                for (int s = 0; s < nsets; ++s) {
                    int t, stn;
                    if (!job_of(k, s, skew, p.S, t, stn)) continue;
                    const int ph = stn - 1;                       // stations B..E consume the exchange H1, H2, F1, F2
                    if (ph < ph0 || ph >= ph1) continue;
                    const int vg = VG_OF(s);
                    const bool ok = __all_sync(0xffffffffu, wait_counter(p, ctl, p.counters + vg * 4 + ph, arrivals * (unsigned int)(t + 1)));
                    fence_proxy_async();
                    if (s == 0) trace(p, t, 12 + ph);
                    trace(p, t, 64 + (ph * 4 + s) * 4 + 0);          // queue timeline: counter seen
                    const CUtensorMap* map = ph == 0 ? &tmH1 : (ph == 1 ? &tmH2 : (ph == 2 ? &tmF1 : &tmF2));
                    for (int kb = 0; kb < kNKB; kb += kTcKbPerOp, ++q) {
                        const uint32_t slot = q % nslots, round = q / nslots;
                        bool go = ok;
                        if (round > 0) go = __all_sync(0xffffffffu, wait_mbar(p, ctl, &ctl->empty[slot], (round - 1) & 1)) && go;
                        if (ph == 0 && s == 1) trace(p, t, 128 + kb);              // tile timeline of one job: slot free
                        go = go && !aborted(p, ctl);
                        if (__all_sync(0xffffffffu, go) && elect_one()) {
                            if (paired) {
                                if (leader) mbar_arrive_expect_tx(&ctl->full[slot], 2 * kTcKbPerOp * kb_bytes);
                                tma_load_3d_pair(smem + ring0 + slot * slot_bytes, map, &ctl->full[slot], 0, vg * 128, kb);
                            } else {
                                mbar_arrive_expect_tx(&ctl->full[slot], kTcKbPerOp * kb_bytes);
                                tma_load_3d(smem + ring0 + slot * slot_bytes, map, &ctl->full[slot], 0, vg * 128, kb);
                            }
                        }
                        __syncwarp();
                    }
                    if (s == 0) trace(p, t, 16 + ph);
                    trace(p, t, 64 + (ph * 4 + s) * 4 + 1);          // last tile issued
                }
            }
        }
    } else if (warp == NEPI + 1) {
        // =================================== MMA issuer =====================================================
        // whole warp in the loop, one elected lane issues: no uniformity waterfall around the UTCHMMAs (tc_common.cuh: elect_one)
        if (leader) {
            uint32_t q = 0;
            for (int k = 0; k < nslot_total; ++k) {
                for (int s = 0; s < nsets; ++s) {
                    int t, stn;
                    if (!job_of(k, s, skew, p.S, t, stn)) continue;
                    const int ph = stn - 1;
                    if (ph < ph0 || ph >= ph1) continue;
                    const uint32_t wofs_ph = ph == 0 ? (uint32_t)kWB : (ph == 1 ? (uint32_t)kWC : (ph == 2 ? (uint32_t)kWD
                                             : ((rawq && sampler) ? 0u : (molsamp ? (uint32_t)kRing : (uint32_t)kWE))));
                    const uint32_t ncol_ph = ph == 0 ? (uint32_t)NB_ : (ph == 1 ? (uint32_t)NC_ : (ph == 2 ? (uint32_t)ND_ : (uint32_t)NE));
                    const uint32_t acc_ph = ph == 0 ? accB : (ph == 1 ? accC : (ph == 2 ? accD : (sampler ? 0u : (uint32_t)kAccE)));
                    const uint32_t idesc = paired ? umma_idesc_f16(256, 2 * (int)ncol_ph) : umma_idesc_f16(128, (int)ncol_ph);
                    const uint32_t dcol = tmem + (uint32_t)s * set_cols + acc_ph;
                    for (int kb = 0; kb < kNKB; kb += kTcKbPerOp, ++q) {
                        const uint32_t slot = q % nslots, round = q / nslots;
                        if (s == 0 && ph == 0 && kb < 4) trace(p, t, 24 + kb);
                        const bool ok = __all_sync(0xffffffffu, wait_mbar(p, ctl, &ctl->full[slot], round & 1));
                        tcgen05_fence_after();
                        if (s == 0 && ph == 0 && kb < 4) trace(p, t, 20 + kb);
                        if (kb == 0) trace(p, t, 64 + (ph * 4 + s) * 4 + 2);     // first tile landed
                        if (ph == 0 && s == 1) trace(p, t, 136 + kb);              // tile landed (MMA thread saw it)
                        if (elect_one()) {
                            if (ok) {
#pragma unroll
                                for (int kk = 0; kk < kTcKbPerOp; ++kk) {
                                    const uint64_t ad = umma_desc_sw128(smem_u32(smem + ring0 + slot * slot_bytes) + kk * kb_bytes);
                                    const uint64_t bd = umma_desc_sw128(smem_u32(smem + wofs_ph + (kb + kk) * ncol_ph * 128));
                                    if (PAIR && paired) {
                                        if (kb + kk == 0) umma_f16_pair<false>(dcol, ad, bd, idesc); else umma_f16_pair<true>(dcol, ad, bd, idesc);
                                        umma_f16_pair<true>(dcol, umma_desc_advance(ad, 32), umma_desc_advance(bd, 32), idesc);
                                        umma_f16_pair<true>(dcol, umma_desc_advance(ad, 64), umma_desc_advance(bd, 64), idesc);
                                        umma_f16_pair<true>(dcol, umma_desc_advance(ad, 96), umma_desc_advance(bd, 96), idesc);
                                    } else {
                                        if (kb + kk == 0) umma_f16_c<false>(dcol, ad, bd, idesc); else umma_f16_c<true>(dcol, ad, bd, idesc);
                                        umma_f16_c<true>(dcol, umma_desc_advance(ad, 32), umma_desc_advance(bd, 32), idesc);
                                        umma_f16_c<true>(dcol, umma_desc_advance(ad, 64), umma_desc_advance(bd, 64), idesc);
                                        umma_f16_c<true>(dcol, umma_desc_advance(ad, 96), umma_desc_advance(bd, 96), idesc);
                                    }
                                }
                            }
                            if (PAIR && paired) umma_commit_pair(&ctl->empty[slot]); else umma_commit(&ctl->empty[slot]);
                            if (kb + kTcKbPerOp >= kNKB) {
                                if (PAIR && paired) umma_commit_pair(&ctl->accfull[s * 4 + ph]); else umma_commit(&ctl->accfull[s * 4 + ph]);
                            }
                        }
                        __syncwarp();
                        if (ph == 0 && s == 1) trace(p, t, 144 + kb);              // MMAs + commit issued
                    }
                    trace(p, t, 64 + (ph * 4 + s) * 4 + 3);          // all MMAs of the job issued
                }
            }
        }
    } else if (warp == NEPI + 2) {
        // =================================== publisher ======================================================
        // walks the stages in the order the epilogue warps publish them
        if (!sampler && !(p.flags & 1)) {
            for (int k = 0; k < nslot_total; ++k)
                for (int s = 0; s < nsets; ++s) {
                    int t, stn;
                    if (!job_of(k, s, skew, p.S, t, stn) || stn >= 4) continue;
                    publisher_release(s, p.counters + VG_OF(s) * 4 + stn);
                    // stage A is the only reader of the conditioning records: after the last set's, the chunk may be recycled
                    if (p.cs_consumed && stn == 0 && s == nsets - 1 && ((t + 1) % kExpandSteps == 0) && lane == 0)
                        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.cs_consumed + t / kExpandSteps) : "memory");
                }
        }
    } else {
        // =================================== epilogue warps =================================================
        // thread = (fold row, unit pair up): TMEM lane = row; my units are 8*(ctab+h) + 2*up + {0,1} for unit block h < H
        const int row = (warp & 3) * 32 + lane, up = warp >> 2;
        const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
        const bool inl = (p.flags & 1) != 0;
        const size_t cs_rec = (size_t)kTcCtas * 4 * 4;                      // float4 per (t,row) record = 1024
        const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        // per-set quantities derived from the set index
#define SET_VIEW(s)                                                                                              \
        const int vg = VG_OF(s);                                                                                 \
        const int fold0 = vg * p.Mg;                                                                             \
        const int nrows = max(0, min(p.Mg, p.B - fold0));                                                        \
        const bool live = row < nrows;                                                                           \
        const size_t grow = (size_t)vg * 128 + row;                                                              \
        const uint32_t tacc = tlane + (uint32_t)(s) * set_cols;                                                  \
        unsigned int* const ctrs = p.counters + vg * 4;                                                          \
        (void)fold0; (void)nrows; (void)live; (void)grow; (void)tacc; (void)ctrs;

        if (sampler) {
            // ---- E on a sampler CTA: fc3 + the draw for every set of the group ---------------------------------
            uint32_t sfold[GS], sutt[GS];
#pragma unroll
            for (int s = 0; s < GS; ++s) {
                sfold[s] = sutt[s] = 0u;
                if (s >= nsets) continue;
                SET_VIEW(s)
                const FoldDesc fd = p.folds[live ? fold0 + row : 0];
                sfold[s] = (uint32_t)fd.fold; sutt[s] = (uint32_t)fd.utt;
            }
            for (int k = 0; k < nslot_total; ++k) {
#pragma unroll
                for (int s = 0; s < GS; ++s) {
                    int t, stn;
                    if (s >= nsets || !job_of(k, s, skew, p.S, t, stn) || stn != 4) continue;
                    SET_VIEW(s)
                    const uint32_t par = (uint32_t)t & 1u;
                    if (!mol) {
                        // RAW sampler: a group of four warps (thread = fold) per (set, 128 classes): warps 4s..4s+3 own set s; on a
                        // pair the accumulator is 256 classes wide and warp groups 2s, 2s+1 own its two halves (= quarters qd&~1, +1)
                        const int wg = warp >> 2, ch = pairs ? (wg & 1) : 0;
                        if (s != (pairs ? (wg >> 1) : wg)) continue;
                        wait_mbar(p, ctl, &ctl->accfull[s * 4 + 3], par);
                        tcgen05_fence_after();
                        raw_sampler_e(p, ctl, reinterpret_cast<const float*>(smem + kRawBias) + ch * kRawQCols, tacc + ch * kRawQCols,
                                      pairs ? (qd & ~1) + ch : qd, fold0, nrows, grow, row, t, sfold[s], sutt[s], key);
                        tcgen05_fence_before();
                        continue;
                    }
                    // MOL (vocoder/distribution.py:104-140): all 30 outputs of a fold sit in one TMEM lane.  The four
                    // threads of a fold split the Gumbel draws (thread `up` owns Philox block `up`, i.e. mixtures
                    // 4up..4up+3), meet through shared memory, and thread up==2 (which also holds the logistic uniform,
                    // block 2 word 2) finishes the draw.
                    // The noise depends on (step, fold) only.  With ONE set per sampler the step is a pure latency chain and the noise
                    // is drawn while the fc3 MMAs are still running (25.6 vs 26.1 us per step at 213 folds); with several sets per
                    // sampler the same change measured 4 % slower (41.1 vs 39.3 us at 1024 folds), so those keep the plain order.
                    if constexpr (GS == 1) {
                        const uint4 r = philox4x32_10(make_uint4((uint32_t)t, sfold[s], sutt[s], (uint32_t)(up < 3 ? up : 2)), key);
                        float gum[4];
#pragma unroll
                        for (int w = 0; w < 4; ++w) gum[w] = -__logf(-__logf(1e-5f + u01(word_of(r, w)) * (1.0f - 2e-5f)));
                        const float ul = 1e-5f + u01(r.z) * (1.0f - 2e-5f);
                        const float lnoise = __logf(ul) - __logf(1.0f - ul);
                        wait_mbar(p, ctl, &ctl->accfull[s * 4 + 3], par);
                        tcgen05_fence_after();
                        if (tid == 0 && s == 0) trace(p, t, 9);
                        float lg[32];
                        tmem_ld8(tacc + 0, lg); tmem_ld8(tacc + 8, lg + 8);
                        tmem_ld8(tacc + 16, lg + 16); tmem_ld8(tacc + 24, lg + 24);
                        tmem_ld_wait();
                        const float* sbias = reinterpret_cast<const float*>(smem + kBias);
                        float2* scratch = reinterpret_cast<float2*>(smem + kMolScratch + s * kMolScratchBytes);
                        float best = -INFINITY;
                        int kbest = 0;
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            const int i = up * 4 + w;
                            if (i < 10) {
                                float li = 0.f;
#pragma unroll
                                for (int q = 0; q < 10; ++q) if (q == i) li = lg[q] + sbias[q];
                                const float sc = li + gum[w];
                                if (sc > best) { best = sc; kbest = i; }
                            }
                        }
                        scratch[row * 4 + up] = make_float2(best, __int_as_float(kbest));
                        asm volatile("bar.sync 2, %0;" ::"n"(NEPI * 32) : "memory");
                        if (up == 2 && live) {
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
                            float xs = mean + __expf(lsc) * lnoise;
                            xs = fminf(fmaxf(xs, -1.0f), 1.0f);
                            p.samples[(size_t)(fold0 + row) * p.S + t] = xs;
                            const float fed = p.forced ? p.forced[(size_t)(fold0 + row) * p.S + t] : xs;
                            ll_store(p.bX + grow, fed, (uint32_t)t + 1u);
                            if (p.logits_out)
                                for (int i = 0; i < 30; ++i) p.logits_out[((size_t)(fold0 + row) * p.S + t) * 30 + i] = lg[i] + sbias[i];
                        }
                    } else {
                        wait_mbar(p, ctl, &ctl->accfull[s * 4 + 3], par);
                        tcgen05_fence_after();
                        if (tid == 0 && s == 0) trace(p, t, 9);
                        float lg[32];
                        tmem_ld8(tacc + 0, lg); tmem_ld8(tacc + 8, lg + 8);
                        tmem_ld8(tacc + 16, lg + 16); tmem_ld8(tacc + 24, lg + 24);
                        tmem_ld_wait();
                        const float* sbias = reinterpret_cast<const float*>(smem + kBias);
                        float2* scratch = reinterpret_cast<float2*>(smem + kMolScratch + s * kMolScratchBytes);
                        const uint4 r = philox4x32_10(make_uint4((uint32_t)t, sfold[s], sutt[s], (uint32_t)(up < 3 ? up : 2)), key);
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
                        if (up == 2 && live) {
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
                            p.samples[(size_t)(fold0 + row) * p.S + t] = xs;
                            const float fed = p.forced ? p.forced[(size_t)(fold0 + row) * p.S + t] : xs;
                            ll_store(p.bX + grow, fed, (uint32_t)t + 1u);
                            if (p.logits_out)
                                for (int i = 0; i < 30; ++i) p.logits_out[((size_t)(fold0 + row) * p.S + t) * 30 + i] = lg[i] + sbias[i];
                        }
                    }
                    tcgen05_fence_before();
                    if (tid == 0 && s == 0) trace(p, t, 10);
                }
            }
        } else {
        // per-unit constants of my units: shared memory (every lane of a warp reads the same word: a broadcast), not registers
        const float* kc0 = reinterpret_cast<const float*>(smem + kConst) + up * 18;
#define KC(h) (kc0 + (h) * 72)
        SetState<H> st[NSETS];
#pragma unroll
        for (int s = 0; s < NSETS; ++s) {
            SET_VIEW(s)
            const FoldDesc fd = p.folds[live ? fold0 + row : 0];
            st[s].fold = (uint32_t)fd.fold; st[s].utt = (uint32_t)fd.utt;
            st[s].x = 0.f;
#pragma unroll
            for (int i = 0; i < 2 * H; ++i) st[s].h1[i] = st[s].h2[i] = st[s].p3[i] = 0.f;
        }
        // conditioning record of unit block h, my unit pair: 4 float4 at
#define CS_OF(h) (p.CS + (((size_t)vg * p.cs_steps + (t % p.cs_steps)) * p.Mg + row) * cs_rec + ((size_t)(ctab + (h)) * 4 + up) * 4)
#define J0_OF(h) ((ctab + (h)) * kTcUnits + 2 * up)

        // ---- A: x_{t-1}, GRU1 for my units, publish h1 -------------------------------------------------------
        auto stageA = [&](SetState<H>& S, const int s, const int t) {
            SET_VIEW(s)
            float4 ca[H];
            float2 cn[H];
            if (p.cs_done && (t % kExpandSteps) == 0 && live)     // the expansion runs on the spare SMs, a few chunks ahead of us
                wait_counter(p, ctl, p.cs_done + t / kExpandSteps, (unsigned int)p.B);
#pragma unroll
            for (int h = 0; h < H; ++h) {
                ca[h] = make_float4(0.f, 0.f, 0.f, 0.f); cn[h] = make_float2(0.f, 0.f);
                if (live) {                                         // issued before the wait on x
                    const float4* cs = CS_OF(h);
                    ca[h] = __ldcs(cs); cn[h] = __ldcs(reinterpret_cast<const float2*>(cs + 1));
                    S.cr[h] = __ldcs(reinterpret_cast<const float2*>(cs + 1) + 1); S.cz[h] = __ldcs(cs + 2); S.c34[h] = __ldcs(cs + 3);
                }
            }
            if (tid == 0) trace(p, t, s == 0 ? 0 : 31);
            etrace(p, t, s, 0);
            S.x = 0.f;
#if WRNN_X_GATHER
            // The previous sample of every fold comes back as a tagged word.  ONE warp polls the set's 128 words (lane = four
            // rows, two 16-byte loads) and hands them to the other fifteen through shared memory and a named barrier: with all
            // 512 threads of 64 CTAs spinning on the same eight L2 lines the last CTA saw x 3-7 us after the first
            // (tools/tc_skew.py), and every stage B waits for the last CTA's stage A.
            if (t > 0) {
                float* xb = reinterpret_cast<float*>(smem + kXBuf) + s * 128;
                if (warp == 0) {
                    const unsigned long long* w = p.bX + (size_t)vg * 128 + 4 * lane;
                    uint32_t pending = 0;
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        if (4 * lane + 2 * i < nrows) pending |= 1u << i;
                    long long t0 = 0;
                    int spins = 0;
                    while (pending) {
#pragma unroll
                        for (int i = 0; i < 2; ++i)
                            if ((pending >> i) & 1u) {
                                unsigned long long a, b;
                                ll_load2(w + 2 * i, a, b);
                                const bool need_b = 4 * lane + 2 * i + 1 < nrows;
                                if (ll_tag(a) == (uint32_t)t && (!need_b || ll_tag(b) == (uint32_t)t)) {
                                    xb[4 * lane + 2 * i] = ll_val(a);
                                    xb[4 * lane + 2 * i + 1] = need_b ? ll_val(b) : 0.f;
                                    pending &= ~(1u << i);
                                }
                            }
                        if (pending && ((++spins) & 255) == 0 && spin_check(p, ctl, t0)) break;
                    }
                }
                asm volatile("bar.sync 7, %0;" ::"n"(NEPI * 32) : "memory");
                if (live) S.x = xb[row];
            }
#else
            if (t > 0 && live) wait_x(p, ctl, p.bX + grow, (uint32_t)t, S.x);
#endif
            if (tid == 0 && s == 0) trace(p, t, 1);
            etrace(p, t, s, 1);
            xtrace(p, t, s, 0);
            float gh[H][8];
            if (t > 0) {
#pragma unroll
                for (int h = 0; h < H; ++h) tmem_ld8(tacc + accB + h * NB_ + 16 * up + 8, gh[h]);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int h = 0; h < H; ++h)
#pragma unroll
                    for (int i = 0; i < 8; ++i) gh[h][i] = 0.f;
            }
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float* v1 = KC(h);
                const float* bh1 = KC(h) + 14;
                const float c1r[2] = {ca[h].x, ca[h].y}, c1z[2] = {ca[h].z, ca[h].w}, c1n[2] = {cn[h].x, cn[h].y};
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const float r = sigmoid_fast(fmaf(v1[0 + u], S.x, c1r[u]) + gh[h][0 + u]);
                    const float z = sigmoid_fast(fmaf(v1[2 + u], S.x, c1z[u]) + gh[h][2 + u]);
                    const float n = tanh_fast(fmaf(v1[4 + u], S.x, c1n[u]) + r * (gh[h][4 + u] + bh1[u]));
                    S.h1[2 * h + u] = (1.0f - z) * n + z * S.h1[2 * h + u];
                }
                if (live) *reinterpret_cast<__half2*>(p.H1 + grow * kRnn + J0_OF(h)) = __floats2half2_rn(S.h1[2 * h], S.h1[2 * h + 1]);
            }
            tcgen05_fence_before();
            publish_arrive(s, inl, ctrs + 0);
            if (tid == 0 && s == 0) trace(p, t, 2);
            etrace(p, t, s, 2);
            xtrace(p, t, s, 1);
        };
        // ---- B: [W_ih2a h1 | W_fc1a h1 | gh1'] ; GRU2 ; publish h2 -------------------------------------------
        auto stageB = [&](SetState<H>& S, const int s, const int t) {
            SET_VIEW(s)
            float pb[H][8], gh[H][8];
            etrace(p, t, s, 3);
            wait_mbar(p, ctl, &ctl->accfull[s * 4 + 0], (uint32_t)t & 1u);
            tcgen05_fence_after();
            if (tid == 0 && s == 0) trace(p, t, 3);
            etrace(p, t, s, 4);
            xtrace(p, t, s, 2);
#pragma unroll
            for (int h = 0; h < H; ++h) {
                tmem_ld8(tacc + accB + h * NB_ + 16 * up, pb[h]);
                if (t > 0) tmem_ld8(tacc + accC + h * NC_ + 8 * up, gh[h]);
                else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) gh[h][i] = 0.f;
                }
            }
            tmem_ld_wait();
            if (tid == 0 && s == 0) trace(p, t, 28);
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float* v2 = KC(h) + 6;
                const float* bh2 = KC(h) + 16;
                float2 cr = make_float2(0.f, 0.f);
                float4 cz = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live) { cr = S.cr[h]; cz = S.cz[h]; }
                const float c2r[2] = {cr.x, cr.y}, c2z[2] = {cz.x, cz.y}, c2n[2] = {cz.z, cz.w};
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const float r = sigmoid_fast(pb[h][0 + u] + fmaf(v2[0 + u], S.x, c2r[u]) + gh[h][0 + u]);
                    const float z = sigmoid_fast(pb[h][2 + u] + fmaf(v2[2 + u], S.x, c2z[u]) + gh[h][2 + u]);
                    const float n = tanh_fast(pb[h][4 + u] + fmaf(v2[4 + u], S.x, c2n[u]) + r * (gh[h][4 + u] + bh2[u]));
                    S.h2[2 * h + u] = (1.0f - z) * n + z * S.h2[2 * h + u];
                    S.p3[2 * h + u] = pb[h][6 + u];
                }
                if (live) *reinterpret_cast<__half2*>(p.H2 + grow * kRnn + J0_OF(h)) = __floats2half2_rn(S.h2[2 * h], S.h2[2 * h + 1]);
            }
            if (tid == 0 && s == 0) trace(p, t, 29);
            tcgen05_fence_before();
            publish_arrive(s, inl, ctrs + 1);
            if (tid == 0 && s == 0) trace(p, t, 4);
            etrace(p, t, s, 5);
        };
        // ---- C: [gh2' | W_fc1a h2] ; f1 ; publish ------------------------------------------------------------
        auto stageC = [&](SetState<H>& S, const int s, const int t) {
            SET_VIEW(s)
            float pb[H][8];
            etrace(p, t, s, 6);
            wait_mbar(p, ctl, &ctl->accfull[s * 4 + 1], (uint32_t)t & 1u);
            tcgen05_fence_after();
            if (tid == 0 && s == 0) trace(p, t, 5);
            etrace(p, t, s, 7);
#pragma unroll
            for (int h = 0; h < H; ++h) tmem_ld8(tacc + accC + h * NC_ + 8 * up, pb[h]);
            tmem_ld_wait();
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float* v3 = KC(h) + 12;
                float2 c3 = make_float2(0.f, 0.f);
                if (live) c3 = make_float2(S.c34[h].x, S.c34[h].y);
                const float f0 = fmaxf(S.p3[2 * h] + pb[h][6] + fmaf(v3[0], S.x, c3.x), 0.f);
                const float f1 = fmaxf(S.p3[2 * h + 1] + pb[h][7] + fmaf(v3[1], S.x, c3.y), 0.f);
                if (live) *reinterpret_cast<__half2*>(p.F1 + grow * kRnn + J0_OF(h)) = __floats2half2_rn(f0, f1);
            }
            tcgen05_fence_before();
            publish_arrive(s, inl, ctrs + 2);
            if (tid == 0 && s == 0) trace(p, t, 6);
            etrace(p, t, s, 8);
        };
        // ---- D: fc2 ; publish ----------------------------------------------------------------------------------
        auto stageD = [&](SetState<H>& S, const int s, const int t) {
            SET_VIEW(s)
            etrace(p, t, s, 9);
            wait_mbar(p, ctl, &ctl->accfull[s * 4 + 2], (uint32_t)t & 1u);
            tcgen05_fence_after();
            if (tid == 0 && s == 0) trace(p, t, 7);
            etrace(p, t, s, 10);
            float d[H][4];
#pragma unroll
            for (int h = 0; h < H; ++h) tmem_ld4(tacc + accD + h * ND_ + 2 * up, d[h]);
            tmem_ld_wait();
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float2 c4 = make_float2(0.f, 0.f);
                if (live) c4 = make_float2(S.c34[h].z, S.c34[h].w);
                if (live) *reinterpret_cast<__half2*>(p.F2 + grow * kRnn + J0_OF(h)) = __floats2half2_rn(fmaxf(d[h][0] + c4.x, 0.f), fmaxf(d[h][1] + c4.y, 0.f));
            }
            tcgen05_fence_before();
            publish_arrive(s, inl, ctrs + 3);
            if (tid == 0 && s == 0) trace(p, t, 8);
            etrace(p, t, s, 11);
            xtrace(p, t, s, 3);
        };

        for (int k = 0; k < nslot_total; ++k) {
#pragma unroll
            for (int s = 0; s < NSETS; ++s) {
                int t, stn;
                if (!job_of(k, s, skew, p.S, t, stn)) continue;
                if (stn == 0) stageA(st[s], s, t);
                else if (stn == 1) stageB(st[s], s, t);
                else if (stn == 2) stageC(st[s], s, t);
                else if (stn == 3) stageD(st[s], s, t);
                else if (!PAIR && !has_samplers) {            // RAW without sampler CTAs: my classes of fc3, then the draw
                    SET_VIEW(s)
                    wait_mbar(p, ctl, &ctl->accfull[s * 4 + 3], (uint32_t)t & 1u);
                    tcgen05_fence_after();
                    raw_stage_e(p, ctl, tacc, cta, fold0, nrows, grow, t, key);
                    tcgen05_fence_before();
                }
            }
            if (blockIdx.x == 0 && tid == 0 && (k % 500) == 0 && p.progress) {
                *reinterpret_cast<volatile int*>(p.progress) = k / 5;
                __threadfence_system();
            }
        }
#undef KC
#undef CS_OF
#undef J0_OF
        }
#undef SET_VIEW
    }
#undef VG_OF
    // ---- teardown ---------------------------------------------------------------------------------------------
    if (aborted(p, ctl)) __nanosleep(200000);      // let any TMA still in flight land before the CTA goes away
    tcgen05_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 0 && !idle) {
        if (paired) tmem_dealloc_pair(tmem, kTmemCols); else tmem_dealloc(tmem, kTmemCols);
    }
}

cudaError_t set_tc_deadline(long long cycles) { return cudaMemcpyToSymbol(g_tc_deadline, &cycles, sizeof(cycles)); }
size_t loop_tc_weight_image_bytes() { return kWBytes; }
size_t loop_tc_raw_sampler_image_bytes() { return kRawBias; }
int loop_tc_raw_sampler_ctas() { return kRawQ; }
int loop_tc_sampler_ctas(int mode, int raw_samplers, int pair) { return mode == 1 ? kTcGroups * (pair ? 2 : 1) : (raw_samplers ? kTcGroups * kRawQ : 0); }

// grid: the unit-owning CTAs of both groups, the sampler CTAs, the expander CTAs; cooperative launch because all CTAs spin
// on each other and must be co-resident.  Pair mode: the same grid as 2-CTA clusters (unit CTAs 2i, 2i+1 form a pair).
template <int NSETS, bool PAIR>
static cudaError_t launch_loop_tc_n(const TcParams& p, const CUtensorMap* m, cudaStream_t stream) {
    cudaError_t err = cudaFuncSetAttribute(wrnn_loop_tc_kernel<NSETS, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes + 1024);
    if (err != cudaSuccess) return err;
    TcParams pp = p;
    void* args[] = {(void*)&m[0], (void*)&m[1], (void*)&m[2], (void*)&m[3], &pp};
    const int grid = kTcGroups * kTcCtas + loop_tc_sampler_ctas(p.mode, p.raw_samplers, p.pair) + (p.cs_done ? p.n_expanders : 0);
    if (PAIR) {
        if (grid & 1) return cudaErrorInvalidValue;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = kSmemBytes + 1024; cfg.stream = stream;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeCooperative;
        at[1].val.cooperative = 1;
        cfg.attrs = at;
        cfg.numAttrs = (getenv("WRNN_TC_COOP") && atoi(getenv("WRNN_TC_COOP")) == 0) ? 1 : 2;
        err = cudaLaunchKernelExC(&cfg, (const void*)wrnn_loop_tc_kernel<NSETS, PAIR>, args);
        if (err != cudaSuccess && cfg.numAttrs == 2) {      // (a profiler may refuse cooperative cluster launches: the grid fits the
            cudaGetLastError();                             //  GPU one CTA per SM, so a plain cluster launch is co-resident as well)
            cfg.numAttrs = 1;
            err = cudaLaunchKernelExC(&cfg, (const void*)wrnn_loop_tc_kernel<NSETS, PAIR>, args);
        }
        return err;
    }
    if (getenv("WRNN_TC_COOP") && atoi(getenv("WRNN_TC_COOP")) == 0) {
        wrnn_loop_tc_kernel<NSETS, PAIR><<<grid, NT, kSmemBytes + 1024, stream>>>(m[0], m[1], m[2], m[3], pp);
        return cudaGetLastError();
    }
    return cudaLaunchCooperativeKernel((const void*)wrnn_loop_tc_kernel<NSETS, PAIR>, dim3(grid), dim3(NT), args, kSmemBytes + 1024, stream);
}

// can `grid` CTAs of the pair kernel be co-resident as 2-CTA clusters on this device? (every TPC must be whole and free)
template <int NSETS>
static bool pair_fits_n(int grid) {
    if (cudaFuncSetAttribute(wrnn_loop_tc_kernel<NSETS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes + 1024) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = kSmemBytes + 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, wrnn_loop_tc_kernel<NSETS, true>, &cfg) != cudaSuccess) { cudaGetLastError(); return false; }
    return 2 * n >= grid;
}
bool loop_tc_pair_fits(int nsets, int grid) { return (grid & 1) == 0 && (nsets <= 2 ? pair_fits_n<1>(grid) : pair_fits_n<2>(grid)); }

// p.nsets = fold sets per group; p.pair: CTA pairs (nsets must be 2 or 4 then)
cudaError_t launch_loop_tc(const TcParams& p, const void* tmaps /* 4 x CUtensorMap */, cudaStream_t stream) {
    const CUtensorMap* m = reinterpret_cast<const CUtensorMap*>(tmaps);
    if (p.pair) {
        if (p.nsets == 2) return launch_loop_tc_n<1, true>(p, m, stream);
        if (p.nsets == 4) return launch_loop_tc_n<2, true>(p, m, stream);
        return cudaErrorInvalidValue;
    }
    switch (p.nsets) {
        case 1: return launch_loop_tc_n<1, false>(p, m, stream);
        case 2: return launch_loop_tc_n<2, false>(p, m, stream);
        case 3: return launch_loop_tc_n<3, false>(p, m, stream);
        case 4: return launch_loop_tc_n<4, false>(p, m, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace wrnn
