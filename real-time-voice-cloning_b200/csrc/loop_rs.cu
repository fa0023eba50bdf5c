// loop_rs.cu -- the latency-bound regime of the autoregressive sample loop (<= 128 folds per CTA group, MOL): a
// ROLE-SPECIALISED tensor-core loop.  DESIGN.md section 4.5 has the measurements behind every choice below.
//
// loop_tc.cu gives every CTA a unit slice of EVERY layer, so every CTA pulls every activation matrix (4 x [folds x 512]
// per step) through a counter + TMA pipeline: ~4.5 us per stage.  Here a CTA owns ONE layer's rows:
//   T1 x16  GRU1, 32 hidden units each: W_hh1 rows (96) + a full copy of fc3 (MOL: 32 rows)
//   T2 x16  GRU2, 32 hidden units each: W_ih2[:, :512] rows (96) + W_hh2 rows (96)
//   T3 x8   fc1, 64 units each          T4 x8   fc2, 64 units each
// (48 CTAs = one group; G groups split the folds).  Consequences:
//  * an activation matrix is read by 8-16 CTAs instead of 64, and only ONE on-path matrix per CTA and step;
//  * fc3 + the mixture draw are replicated on every T1 CTA (same inputs, same Philox counters -> identical samples), so the
//    sample never travels on the critical path: a step has FOUR exchanges (h1, h2|s2, f1, f2) instead of five;
//  * the operand with M = 128 rows is the ACTIVATION matrix (TMEM lane = fold), written straight into TENSOR MEMORY by the
//    threads that receive it (tcgen05.st) and consumed from there (tcgen05.mma, A in TMEM): no shared-memory staging, no
//    TMA, no proxy fence; at the tensor pipe's N/2 clocks per K=16 step (tools/probes/mma_rate3.cu);
//  * the exchange has no fence, counter or flag: activations travel as 16-byte chunks [chunk][fold] (8 fp16 of one fold)
//    whose validity is a GENERATION BIT carried in every half -- bit 14 for |v| < 2 (GRU states, their sum), the sign bit
//    for ReLU outputs -- so the payload keeps all its precision, a torn chunk can never validate, and nothing is ever
//    reset (tools/probes/xchg9.cu: sentinel resets cost ~1 us per exchange).  Double-buffered by step parity; generation
//    = (step / 2) & 1.  Receivers poll one canary chunk per producer, then load everything once and re-poll stragglers;
//  * the recurrent products W_hh1 h1(t), W_hh2 h2(t) for step t+1 run off the critical path on the same A buffer.
// Conditioning: per-sample records [group][step][fold][8][512] fp32 (c1 r,z,n | c2 r,z,n | c3 | c4), read by the owning
// thread only, prefetched before the step's wait.
// Every wait has a deadline ("soft abort", as loop_tc.cu): a bad build ends with an error code, never a hung GPU.
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include "engine_internal.h"
#include "tc_common.cuh"

namespace wrnn {

__device__ long long g_rs_deadline = 1500000000LL;

namespace {
using namespace tc;

constexpr unsigned FULL = 0xffffffffu;
constexpr int NW = 16;                        // ingest / epilogue warps: warp w = (lane quadrant q = w & 3, column slice cs = w >> 2)
constexpr int NT = (NW + 4) * 32;             // + a service warpgroup: the MMA warp and three idle warps (setmaxnreg works on whole warpgroups)
// 16 x 112 + 4 x 32 = 20 x 96, the launch allocation.  The chain warps take everything the MMA warp can spare: against 104 / 64 the
// kernel's spills fall from 264 to 220 bytes (MOL instantiation) and the step by 0.5 % (MOL, 213 folds) to 1.5 % (RAW); 19 MOL folds
// lose 0.5 % (profiles/r2_probes/ab_regs.txt).  -DWRNN_RS_REGS_EPI=104 -DWRNN_RS_REGS_SVC=64 restores the earlier split.
#ifndef WRNN_RS_REGS_EPI
#define WRNN_RS_REGS_EPI 112
#define WRNN_RS_REGS_SVC 32
#endif
constexpr int kRegsEpi = WRNN_RS_REGS_EPI, kRegsSvc = WRNN_RS_REGS_SVC;  // (the SM's spare registers are NOT available to setmaxnreg.inc: measured, it blocks for ever)
constexpr int kChunks = kRnn / 8;             // 64 chunks of 8 fp16 per activation row
constexpr size_t kMatChunks = (size_t)kChunks * 128;     // one buffer of one exchange matrix: [chunk][fold] x 16 bytes = 128 KB
enum { MH1 = 0, MH2, MS2, MF1, MF2, kMats };
constexpr uint32_t kTagE = 0x40004000u;       // generation bit of matrices with |v| < 2: bit 14 of every half
constexpr uint32_t kTagS = 0x80008000u;       // of non-negative matrices (ReLU outputs): the sign bit
// TMEM columns
constexpr uint32_t kColA = 0;                 // activation operand, 128 lanes x 256 columns (512 fp16 per lane)
constexpr uint32_t kColD0 = 256;              // on-path accumulator (T1: W_hh1 h1 [96] lives here too, see below)
constexpr uint32_t kColD1 = 352;              // second accumulator (T2: W_hh2 h2)
constexpr uint32_t kColD1T1 = 384;            // T1's second accumulator (fc3): its D0 is 128 columns wide with the inline conditioning
constexpr uint32_t kColX = 448;               // inline conditioning: 80 fp16 of upsampled mel per fold = 40 columns, K steps 0..4
// shared memory: weight tiles (K-major SWIZZLE_128B, [k-block 8][rows N][128 B]) then constants and the control block
constexpr int kW0 = 0;                                   // first tile: T1 W_hh1 (96 rows), T2 W_ih2a (96), T3 fc1a (64), T4 fc2 (64)
constexpr int kW1 = 96 * 128 * 8;                        // second tile: T1 fc3 (32 rows), T2 W_hh2 (96)
constexpr int kNoiseOfs = kW1 + 32 * 128 * 8;            // T1 only: the step's mixture noise, [12][512 threads] floats (24 KB behind the fc3 tile)
constexpr int kWEnd = kW1 + 96 * 128 * 8;                // 196608
constexpr int kFU = kRnn / kRsT3;                        // units per FC-role CTA: 64 (8 CTAs per role) or 32 (16)
static_assert(kRsT3 == kRsT4 && (kFU == 128 || kFU == 64 || kFU == 32), "thread <-> unit map of the FC roles");
// inline conditioning (kInl): the MEL share of every record is one more K = 80 slab of the role's on-path product: tiles
// [k-block 2][rows][128 B] of W_q = (W_ih1 | W_ih2a | fc1a) . I[:, mel] next to the role's other tiles
constexpr int kXW1 = kNoiseOfs + 12 * 512 * 4;           // T1: 128 rows (r, z, 32 zero rows, n) behind the noise buffer
constexpr int kXW2 = kWEnd;                              // T2: 96 rows behind its two tiles
constexpr int kXW3 = kFU * 128 * 8;                      // T3: kFU rows behind fc1a
constexpr int kXWEnd = kWEnd + 96 * 128 * 2;
static_assert(kXW1 + 128 * 128 * 2 <= kWEnd && (kXW1 & 1023) == 0 && kXW3 + kFU * 128 * 2 <= kWEnd, "inline-conditioning tiles");
constexpr int kConstOfs = kXWEnd;                        // per-unit constants, <= 5 x 64 floats
constexpr int kCtlOfs = kConstOfs + 2048;
constexpr int kSmemBytes = kCtlOfs + 256;
static_assert(kNoiseOfs + 12 * 512 * 4 <= kWEnd, "T1 noise buffer");

struct Ctl {
    uint64_t abar[4];      // K quarter kq of the A operand is in TMEM (all 16 warps arrive)
    uint64_t dbar[2];      // accumulator complete (tcgen05.commit): [0] on-path job, [1] recurrent (off-path) job
    uint64_t ebar;         // all 16 epilogue warps have read the recurrent accumulator of the previous step
    uint64_t xbar;         // inline conditioning: the step's upsampled-mel operand is in TMEM (all 16 warps arrive)
    uint32_t tmem;
    int abort_local;
};
// barriers are addressed by their 32-bit shared-memory address (ctl_s + offset): the generic -> shared conversion of a pointer
// costs ~10 uniform instructions wherever the compiler rematerialises it, and every instruction of the chain counts
constexpr uint32_t kBarA = 0, kBarD = 32, kBarE = 48, kBarX = 56;

__device__ __forceinline__ bool aborted_local(Ctl* c) { return *reinterpret_cast<volatile int*>(&c->abort_local) != 0; }
// warp-uniform view of the abort flag (lane 0's): the step loops end together for all lanes of a warp
__device__ __forceinline__ bool warp_aborted(Ctl* c) { return __shfl_sync(FULL, aborted_local(c) ? 1 : 0, 0) != 0; }
// slow path of every wait (each 256th or 1024th try): abort flags and the deadline.  t0 (the clock at the first slow check) is
// passed BY VALUE: by reference it lived in local memory and every wait paid a store for it.
__device__ __noinline__ bool spin_check(const RsParams& p, Ctl* c, long long t0) {
    if (aborted_local(c)) return true;
    if (ld_volatile_i32(p.abort_flag) != 0) { *reinterpret_cast<volatile int*>(&c->abort_local) = 1; return true; }
    if (clock64() - t0 > g_rs_deadline) {
        *reinterpret_cast<volatile int*>(&c->abort_local) = 1;
        atomicExch(p.abort_flag, 1);
        return true;
    }
    return false;
}
#define RS_SPIN_CHECK(mask) ((((++spins) & (mask)) == 0) && ((t0 = (t0 == 0 ? clock64() : t0)), spin_check(p, ctl, t0)))

__device__ __forceinline__ bool mbar_try_wait_s(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void umma_commit_s(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// kSleep: the 16 ingest / epilogue warps wait for the accumulator while the MMA warp issues: every instruction they spend
// polling is an issue slot the MMA warp does not get (measured: +30 clocks per MMA), so they back off between tries
template <bool kSleep = false>
__device__ __forceinline__ bool wait_mbar(const RsParams& p, Ctl* ctl, uint32_t bar, uint32_t parity) {
    long long t0 = 0;
    int spins = 0;
    while (!mbar_try_wait_s(bar, parity)) {
        if (kSleep) __nanosleep(64);
        if (RS_SPIN_CHECK(1023)) return false;
    }
    return true;
}

// Instrumentation exists only in the kTrace instantiation of the kernel (WRNN_RS_TRACE / WRNN_RS_DEBUG select it at launch):
// in the production instantiation every hook below compiles to nothing -- a disabled hook still cost 6-8 issue slots.
// dbg: checkpoints into mapped host memory (WRNN_RS_DEBUG=1): [CTA][32 warps] last (step << 8 | code) of lane 0
template <bool kTrace>
__device__ __forceinline__ void dbg(const RsParams& p, int t, int code) {
    if constexpr (kTrace) {
        if (p.dbg && (threadIdx.x & 31) == 0) *reinterpret_cast<volatile int*>(p.dbg + blockIdx.x * 32 + (threadIdx.x >> 5)) = (t << 8) | code;
    }
}
// timeline (WRNN_RS_TRACE=path): %globaltimer (ns, common to all SMs) of steps [kTraceStep0, +kTraceSteps) per CTA, written by
// lane 0 of warp 0 (events 0..8) and of the MMA warp (9..12): [CTA][step][48]
constexpr int kTraceStep0 = 96, kTraceSteps = 8;
template <bool kTrace>
__device__ __forceinline__ void trace(const RsParams& p, int t, int ev) {
    if constexpr (kTrace) {
        if (p.trace && (threadIdx.x == 0 || threadIdx.x == NW * 32) && t >= kTraceStep0 && t < kTraceStep0 + kTraceSteps) {
            unsigned long long ns;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
            p.trace[((size_t)blockIdx.x * kTraceSteps + (t - kTraceStep0)) * 48 + ev] = ns;
        }
    }
}
// SM-clock stamps of warp 0 / lane 0 (slots 32 + k): the globaltimer ticks too coarsely (32-256 ns) for the inside of an epilogue
template <bool kTrace>
__device__ __forceinline__ void ctrace(const RsParams& p, int t, int k) {
    if constexpr (kTrace) {
        if (p.trace && threadIdx.x == 0 && t >= kTraceStep0 && t < kTraceStep0 + kTraceSteps)
            p.trace[((size_t)blockIdx.x * kTraceSteps + (t - kTraceStep0)) * 48 + 32 + k] = (unsigned long long)clock64();
    }
}
// per-K-quarter events of the first ingest of a step (warp 0, lane 0): slot 16 + kq = quarter kq of my rows is in TMEM
template <bool kTrace>
__device__ __forceinline__ void trace_kq(const RsParams& p, int t, int ev0, int kq) {
    if constexpr (kTrace) {
        if (p.trace && ev0 == 1 && threadIdx.x == 0 && t >= kTraceStep0 && t < kTraceStep0 + kTraceSteps) {
            unsigned long long ns;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
            p.trace[((size_t)blockIdx.x * kTraceSteps + (t - kTraceStep0)) * 48 + 16 + kq] = ns;
        }
    }
}

__device__ __forceinline__ uint4 ld_chunk(const uint4* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_chunk(uint4* p, uint4 v) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// generation bits of a chunk / of four chunks: nonzero = not (all) of this generation yet.  gen1 is warp-uniform, so each side
// is a tree of three-input logic operations (8 for 16 words) instead of one XOR-OR per word
__device__ __forceinline__ uint32_t bad1(const uint4& v, uint32_t tb, bool gen1) {
    return gen1 ? (~(v.x & v.y & v.z & v.w) & tb) : ((v.x | v.y | v.z | v.w) & tb);
}
__device__ __forceinline__ uint32_t bad4(const uint4* v, uint32_t tb, bool gen1) {
    if (gen1) {
        const uint32_t a = (v[0].x & v[0].y & v[0].z) & (v[0].w & v[1].x & v[1].y) & (v[1].z & v[1].w & v[2].x) & (v[2].y & v[2].z & v[2].w) &
                           (v[3].x & v[3].y & v[3].z) & v[3].w;
        return ~a & tb;
    }
    const uint32_t o = (v[0].x | v[0].y | v[0].z) | (v[0].w | v[1].x | v[1].y) | (v[1].z | v[1].w | v[2].x) | (v[2].y | v[2].z | v[2].w) |
                       (v[3].x | v[3].y | v[3].z) | v[3].w;
    return o & tb;
}
// K parts of an ingest / MMA job, in k-blocks of 64 fp16 (8 chunks; 4 K = 16 steps): the matrix arrives and is multiplied part by
// part; two parts are requested up front, part k + 2 when part k has arrived.  Four even quarters.  -DWRNN_RS_UNEVEN_PARTS makes the
// last part the smallest ({2, 2, 3, 1} k-blocks: fewer MMAs exposed behind the last byte) -- measured SLOWER, 13.8 vs 12.9 us per
// 213-fold step: a T2 job (N = 96, 55 clocks per K step = the tensor pipe's peak rate at M = 128) is already as long as the
// ingest of <= 107 folds, so a larger third part only queues MMAs, and 14 chunks in flight instead of 12 spill another 100 bytes.
#ifndef WRNN_RS_UNEVEN_PARTS
#define RS_PART_KB(kq) 2
#define RS_PART_KB0(kq) (2 * (kq))
#else
#define RS_PART_KB(kq) ((kq) == 2 ? 3 : ((kq) == 3 ? 1 : 2))
#define RS_PART_KB0(kq) ((kq) == 3 ? 7 : 2 * (kq))
#endif
template <int N>
__device__ __forceinline__ uint32_t badN(const uint4* v, uint32_t tb, bool gen1) {
    if constexpr (N == 4) return bad4(v, tb, gen1);
    if (gen1) {
        uint32_t a = 0xFFFFFFFFu;
#pragma unroll
        for (int i = 0; i < N; ++i) a &= (v[i].x & v[i].y) & (v[i].z & v[i].w);
        return ~a & tb;
    }
    uint32_t o = 0u;
#pragma unroll
    for (int i = 0; i < N; ++i) o |= (v[i].x | v[i].y) | (v[i].z | v[i].w);
    return o & tb;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint4& a, const uint4& b, const uint4& c, const uint4& d) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w),
        "r"(d.x), "r"(d.y), "r"(d.z), "r"(d.w)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint4& a, const uint4& b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(a.x), "r"(a.y), "r"(a.z),
                 "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T, A = 128 lanes x 8 columns (16 fp16 per lane)
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// exp / reciprocal as ONE special-function instruction each (ex2.approx.ftz / rcp.approx.ftz): __expf and __fdividef wrap
// the same MUFU operations in three more instructions of denormal handling.
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// per-unit constants live in shared memory: read them with ld.shared at a 32-bit address
__device__ __forceinline__ float4 lds4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void lds8(uint32_t a, float* out) {
    const float4 x = lds4(a), y = lds4(a + 16);
    out[0] = x.x; out[1] = x.y; out[2] = x.z; out[3] = x.w; out[4] = y.x; out[5] = y.y; out[6] = y.z; out[7] = y.w;
}
__device__ __forceinline__ float lds1(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts1(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }

// The GRU cell of 8 units: pr / pz = complete pre-activations of the reset / update gates, pn = the input side of the candidate,
// bn = W_hn h + b_hn (what the reset gate multiplies).  The epilogues are bound by the special-function unit (16 results per
// clock and SM: 6 per unit in the textbook form = 1536 clocks per step for 512 threads x 8 units), so reciprocals are SHARED:
// 1/a and 1/b from one rcp(a b) -- the reset gates of a pair of units, and the update gate with the candidate's tanh of one
// unit: 4.5 special-function results per unit.  The exponents are capped at 2^30 so that no product overflows: a sigmoid then
// saturates at 9.3e-10 instead of 0, far below the fp32 resolution of the state it multiplies.
constexpr float kL2E = 1.4426950408889634f;
__device__ __forceinline__ void gru8(const float* pr, const float* pz, const float* pn, const float* bn, float* h) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        const float d0 = 1.0f + ex2_ftz(fminf(-kL2E * pr[i], 30.f)), d1 = 1.0f + ex2_ftz(fminf(-kL2E * pr[i + 1], 30.f));
        const float rr = rcp_ftz(d0 * d1);
        const float r[2] = {d1 * rr, d0 * rr};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float dz = 1.0f + ex2_ftz(fminf(-kL2E * pz[i + j], 30.f));
            const float dn = 1.0f + ex2_ftz(fminf((2.0f * kL2E) * fmaf(r[j], bn[i + j], pn[i + j]), 30.f));
            const float rc = rcp_ftz(dz * dn);
            const float z = dn * rc, n = fmaf(-2.0f * dz, rc, 1.0f);
            h[i + j] = fmaf(z, h[i + j] - n, n);
        }
    }
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack2(uint32_t w) {
    const __half2 h = *reinterpret_cast<const __half2*>(&w);
    return __half22float2(h);
}

// Exchange-side view of a thread: fold lane and the step-independent addresses.
struct Lane {
    int q, cs, lane, row;         // lane quadrant, column slice, lane, fold row inside the group
    int row_ld;                   // the row this lane READS: its own, or the group's last fold for a padding lane of a live warp
                                  // (same bytes as a neighbour: free, and the warp needs no divergent paths)
    bool live, wlive;             // my row is a fold / my warp has at least one fold
    int nlive_threads;            // threads of the CTA's live ingest warps (the named barrier behind the canary poll)
    uint32_t tlane;               // TMEM address of my lane quadrant, column 0
};

// Receive one activation matrix (this step's buffer) into the A operand in TMEM, K QUARTER BY K QUARTER: warp (q, cs) takes
// the four chunks 16 kq + 4 cs + {0..3} of every quarter kq of its 32 folds, so a quarter is complete -- and its eight MMAs can
// run -- while the other quarters are still travelling or being checked.
// `extra` (optional): one more chunk of my row, returned to the caller (T2: my own units of h1).  Warp 0 first polls canary
// chunks (one per lane and producer warp), then every lane requests two quarters; quarter kq + 2 is requested when quarter kq
// has arrived (64 KB per CTA in flight: more than the L2 latency x bandwidth product), so the requests of ALL warps for the
// early quarters are ahead of anybody's late ones and the quarters complete in K order across the CTA.  Per quarter: one
// logic tree over the generation bits, re-poll until they match, strip them, one 16-column store into TMEM whose completion is
// awaited behind the checks of the next quarter, arrive.
// Lean on purpose: the SM issues ~1 instruction per clock and scheduler, 4 of these warps share a scheduler, so every
// instruction here costs ~4 clocks of the step.
// (inlined on purpose: as a real call the ABI spills around it cost more than the code size saves -- 25.8 vs 19.2 us per step)
struct IngestOut { uint4 extra; float x; };
template <bool kTrace>
__device__ __forceinline__ IngestOut ingest(const RsParams& p, Ctl* ctl, uint32_t ctl_s, const Lane& L, const uint4* mat, uint32_t tb, bool gen1,
                                         int extra_chunk, int dbg_t, int ev0, const unsigned long long* xw, uint32_t xtag) {
    IngestOut o;
    o.extra = make_uint4(0u, 0u, 0u, 0u);
    o.x = 0.f;
    dbg<kTrace>(p, dbg_t, 0x10);
    if (!L.wlive) {            // a quadrant of padding rows keeps whatever it holds: its accumulator rows are never read
        if (L.lane == 0) {
#pragma unroll
            for (int kq = 0; kq < 4; ++kq) mbar_arrive_s(ctl_s + kBarA + 8u * kq);
        }
        return o;
    }
    {   // phase 1, ONE warp per CTA: lane l polls the canary chunk 2 l of the group's first fold (32 different producer warps:
        // every CTA of the producing role) until the first of them shows this step's generation (p.canary_all: until all
        // do); the other live warps sleep on a named barrier.  (All 16 warps polling cost 16 KB of L2 traffic per CTA and
        // round trip, with ~100 CTAs waiting at any time: 13.1 -> 12.75 us per step at 213 folds.)
        if (L.q == 0 && L.cs == 0) {
            const uint4* cp = mat + (size_t)(2 * L.lane) * 128;
            long long t0 = 0;
            int spins = 0;
            bool ok = false;
            while (true) {
                if (!ok) ok = bad1(ld_chunk(cp), tb, gen1) == 0u;
                if (p.canary_all ? __all_sync(FULL, ok) : __any_sync(FULL, ok)) break;
                if (((++spins) & 255) == 0) {               // (spins is warp-uniform)
                    if (t0 == 0) t0 = clock64();
                    if (__any_sync(FULL, spin_check(p, ctl, t0))) break;
                }
            }
        }
        asm volatile("bar.sync 2, %0;" ::"r"(L.nlive_threads) : "memory");
    }
    dbg<kTrace>(p, dbg_t, 0x11);
    trace<kTrace>(p, dbg_t, ev0);
    if (ev0 == 1) ctrace<kTrace>(p, dbg_t, 8);
    // my chunks of part kq: 8 kb0 + (2 kb) cs + i, i < 2 kb (kb = k-blocks of the part, kb0 = its first); kept in v[2 kb0 ..]
    const uint4* base = mat + (size_t)(L.cs * 4) * 128 + L.row_ld;
    uint4 v[16];
    auto request = [&](auto KQ) {
        constexpr int kq = decltype(KQ)::value, n = 2 * RS_PART_KB(kq), c0 = 8 * RS_PART_KB0(kq);
        const uint4* b = n == 4 ? base : base + (n - 4) * L.cs * 128;            // (one base pointer; a 4-chunk part is constant offsets from it)
#pragma unroll
        for (int i = 0; i < n; ++i) v[2 * RS_PART_KB0(kq) + i] = ld_chunk(b + (c0 + i) * 128);
    };
    request(std::integral_constant<int, 0>{});
    request(std::integral_constant<int, 1>{});
    const uint4* xp = mat + (size_t)(extra_chunk >= 0 ? extra_chunk : 0) * 128 + L.row_ld;
    if (extra_chunk >= 0) o.extra = ld_chunk(xp);
    unsigned long long xword = 0ull;
    if (xw) xword = ll_load(xw);            // the sample word was published before this matrix: its load rides along
    int passes = 1;
    auto part = [&](auto KQ) {
        constexpr int kq = decltype(KQ)::value, n = 2 * RS_PART_KB(kq), v0 = 2 * RS_PART_KB0(kq);
        uint32_t bad = badN<n>(&v[v0], tb, gen1);
        asm volatile("" ::"r"(bad) : "memory");          // (pins the next requests behind the arrival of this part: volatile asm keeps its order)
        if constexpr (kq < 2) request(std::integral_constant<int, kq + 2>{});
        if (bad != 0u) {                                   // rare: a producer of this part is late
            long long t0 = 0;
            int spins = 0;
            do {
                ++passes;
                request(KQ);
                bad = badN<n>(&v[v0], tb, gen1);
            } while (bad != 0u && !RS_SPIN_CHECK(255));
        }
        if (gen1) {      // generation 1: the bit is set in every half; take it out (generation 0 needs nothing)
#pragma unroll
            for (int i = 0; i < n; ++i) { v[v0 + i].x ^= tb; v[v0 + i].y ^= tb; v[v0 + i].z ^= tb; v[v0 + i].w ^= tb; }
        }
        __syncwarp();
        if constexpr (kq > 0) {    // the previous part's store has had the checks above to complete
            tmem_st_wait();
            tcgen05_fence_before();
            if (L.lane == 0) mbar_arrive_s(ctl_s + kBarA + 8u * (kq - 1));
            trace_kq<kTrace>(p, dbg_t, ev0, kq - 1);
        }
        const uint32_t col = L.tlane + kColA + 4u * (uint32_t)(8 * RS_PART_KB0(kq) + n * L.cs);       // a chunk = 4 columns
        if constexpr (n == 4) tmem_st16(col, v[v0], v[v0 + 1], v[v0 + 2], v[v0 + 3]);
        else {                     // (8-column stores: a 6-chunk part starts at a multiple of 8 columns, not of 16)
#pragma unroll
            for (int i = 0; i < n; i += 2) tmem_st8(col + 4u * i, v[v0 + i], v[v0 + i + 1]);
        }
    };
    part(std::integral_constant<int, 0>{});
    part(std::integral_constant<int, 1>{});
    part(std::integral_constant<int, 2>{});
    part(std::integral_constant<int, 3>{});
    tmem_st_wait();
    tcgen05_fence_before();
    if (L.lane == 0) mbar_arrive_s(ctl_s + kBarA + 24u);
    trace_kq<kTrace>(p, dbg_t, ev0, 3);
    dbg<kTrace>(p, dbg_t, 0x12);
    trace<kTrace>(p, dbg_t, ev0 + 2);
    if (ev0 == 1) ctrace<kTrace>(p, dbg_t, 11);
    if constexpr (kTrace) {
        if (p.trace && threadIdx.x == 0 && dbg_t >= kTraceStep0 && dbg_t < kTraceStep0 + kTraceSteps)
            p.trace[((size_t)blockIdx.x * kTraceSteps + (dbg_t - kTraceStep0)) * 48 + (ev0 == 1 ? 13 : 14)] = (unsigned long long)passes;
    }
    if (extra_chunk >= 0) {
        long long t0 = 0;
        int spins = 0;
        while (bad1(o.extra, tb, gen1) != 0u) {
            o.extra = ld_chunk(xp);
            if (RS_SPIN_CHECK(255)) break;
        }
        if (gen1) { o.extra.x ^= tb; o.extra.y ^= tb; o.extra.z ^= tb; o.extra.w ^= tb; }
    }
    if (xw) {
        long long t0 = 0;
        int spins = 0;
        while (ll_tag(xword) != xtag) {
            xword = ll_load(xw);
            if (RS_SPIN_CHECK(255)) break;
        }
        o.x = ll_val(xword);
    }
    return o;
}

// Publish 8 values of my fold as one chunk (generation bit in every half).
__device__ __forceinline__ void publish8(uint4* mat, int chunk, int row, const float* v, uint32_t tb, uint32_t want) {
    uint4 w;
    w.x = (pack2(v[0], v[1]) & ~tb) | want; w.y = (pack2(v[2], v[3]) & ~tb) | want;
    w.z = (pack2(v[4], v[5]) & ~tb) | want; w.w = (pack2(v[6], v[7]) & ~tb) | want;
    st_chunk(mat + (size_t)chunk * 128 + row, w);
}

// The MOL draw of one fold from its 30 outputs (vocoder/distribution.py:104-140; same arithmetic as loop_tc.cu) in two halves.
// mol_noise: the noise depends on (step, fold) only and is drawn while the step's activations are still travelling; it waits
// in shared memory ([12][512 threads]: 10 Gumbel terms with the mixture-logit biases already added, the logistic noise) so
// that it does not hold 11 registers across the ingest.
__device__ __forceinline__ void mol_noise(uint32_t nz_s, uint32_t sbias_s, uint32_t t, uint32_t fold, uint32_t utt, uint2 key) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const uint4 r = philox4x32_10(make_uint4(t, fold, utt, (uint32_t)b), key);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int i = 4 * b + w;
            if (i < 10) sts1(nz_s + 2048u * i, lds1(sbias_s + 4u * i) - __logf(-__logf(1e-5f + u01(word_of(r, w)) * (1.0f - 2e-5f))));
        }
        if (b == 2) {
            const float ul = 1e-5f + u01(r.z) * (1.0f - 2e-5f);
            sts1(nz_s + 2048u * 10, __logf(ul) - __logf(1.0f - ul));
        }
    }
}
__device__ __forceinline__ float mol_draw(const float* lg, uint32_t nz_s, uint32_t sbias_s) {
    // argmax over the 10 perturbed mixture logits as a TREE (depth 4 instead of a chain of 9 dependent compare-selects: the
    // draw sits on the critical path of every step); ties keep the lower index, as the sequential scan does
    float sc[10], mu[10], ls[10];
    int kx[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) { sc[i] = lg[i] + lds1(nz_s + 2048u * i); mu[i] = lg[10 + i]; ls[i] = lg[20 + i]; kx[i] = i; }
    auto pick = [&](int l, int r) {
        const bool up = sc[r] > sc[l];
        sc[l] = up ? sc[r] : sc[l]; mu[l] = up ? mu[r] : mu[l]; ls[l] = up ? ls[r] : ls[l]; kx[l] = up ? kx[r] : kx[l];
    };
#ifdef RS_DRAW_CHAIN
    for (int i = 1; i < 10; ++i) pick(0, i);
#else
    pick(0, 1); pick(2, 3); pick(4, 5); pick(6, 7); pick(8, 9);
    pick(0, 2); pick(4, 6);
    pick(0, 4);
    pick(0, 8);
#endif
    const float mean = mu[0] + lds1(sbias_s + 40u + 4u * kx[0]);
    const float lsc = fmaxf(ls[0] + lds1(sbias_s + 80u + 4u * kx[0]), -32.23619130191664f);
    const float xs = fmaf(ex2_ftz(kL2E * lsc), lds1(nz_s + 2048u * 10), mean);
    return fminf(fmaxf(xs, -1.0f), 1.0f);
}

// One MMA job: D[128 folds x N] = A (TMEM, 512 fp16 per lane) x W^T (shared memory tile [k-block][N][64]); issued K quarter by K
// quarter as the ingest warps deliver them.  Whole warp in the loop, one elected lane issues (tc_common.cuh: elect_one); a K
// step is one add on the descriptor and one UTCHMMA.
// Inline conditioning: the five K = 16 steps of W_q . m (upsampled mel, TMEM columns kColX..) START the accumulator, so they
// are issued as soon as the operand is there and the accumulator is free -- long before the awaited matrix arrives.
__device__ __forceinline__ void mma_ext(uint32_t xw_smem, uint32_t NX, uint32_t d, uint32_t tmem_x) {
    if (elect_one()) {
        const uint32_t idx = umma_idesc_f16(128, (int)NX);
        const uint64_t bx = umma_desc_sw128(xw_smem);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ts(d, tmem_x + 8u * k, bx + 2u * k, idx, k != 0 ? 1u : 0u);
        umma_ts(d, tmem_x + 32u, bx + NX * 8u, idx, 1u);           // mel 64..79: first K step of the second k-block
    }
    __syncwarp();
}
template <bool kTrace>
__device__ __forceinline__ void mma_job(const RsParams& p, Ctl* ctl, uint32_t ctl_s, uint32_t w_smem, uint32_t N, uint32_t d, uint32_t a0, uint32_t a_par,
                                     uint32_t done_bar, bool wait_e, uint32_t e_par, int tt, int ev, uint32_t xw_smem = 0u, uint32_t NX = 0u,
                                     uint32_t x_par = 0u) {
    const uint32_t idesc = umma_idesc_f16(128, (int)N);
    const uint64_t bd0 = umma_desc_sw128(w_smem);
    const uint32_t kb_step = N * 8u;                      // one k-block of the tile, in descriptor units of 16 bytes
    const bool ext = xw_smem != 0u;
    if (ext) {
        if (wait_e) wait_mbar(p, ctl, ctl_s + kBarE, e_par);
        wait_mbar(p, ctl, ctl_s + kBarX, x_par);
        tcgen05_fence_after();
        mma_ext(xw_smem, NX, d, a0 + kColX);
    }
#pragma unroll 1
    for (int kq = 0; kq < 4; ++kq) {
        wait_mbar(p, ctl, ctl_s + kBarA + 8u * kq, a_par);
        if (kq == 0 && wait_e && !ext) wait_mbar(p, ctl, ctl_s + kBarE, e_par);
        tcgen05_fence_after();
        if (kq == 0) trace<kTrace>(p, tt, ev);
        if (elect_one()) {
            const uint32_t kb0 = (uint32_t)RS_PART_KB0(kq);
            const int nkb = RS_PART_KB(kq);
            uint64_t bd = bd0 + (uint64_t)(kb0 * kb_step);
            uint32_t a = a0 + 32u * kb0;                   // a k-block = 64 fp16 = 32 columns of the A operand
#pragma unroll 1
            for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    umma_ts(d, a, bd + 2u * k, idesc, (ext || (kq | kb | k) != 0) ? 1u : 0u);
                    a += 8u;
                }
                bd += kb_step;
            }
            if (kq == 3) umma_commit_s(done_bar);
        }
        __syncwarp();
    }
    trace<kTrace>(p, tt, ev + 1);
}

// Per-sample conditioning records of the role-specialised loop: CS[group][step % cs_steps][fold][8][512] fp32
// (c1 r,z,n | c2 r,z,n | c3 | c4), same interpolation and FMA order as cond_expand.cuh.  One call = steps [t0, t1) of one
// fold by 512 threads (thread j = hidden unit j); every store is a coalesced 128 bytes per warp.
__device__ __forceinline__ void expand_item_rs(const float4* __restrict__ TA1, const float4* __restrict__ TA2, const float4* __restrict__ TQ1,
                                               const float4* __restrict__ TQ2, const float* __restrict__ coef, const FoldDesc& fd, int g, int row,
                                               int t0, int t1, int cs_steps, int Ng, float* __restrict__ CS, int j) {
    float ta[8], tq[kTaps][7];
#pragma unroll
    for (int i = 0; i < 8; ++i) ta[i] = 0.f;
#pragma unroll
    for (int d = 0; d < kTaps; ++d)
#pragma unroll
        for (int i = 0; i < 7; ++i) tq[d][i] = 0.f;
    int key = -1;
    for (int t = t0; t < t1; ++t) {
        const int n = fd.n0 + t;
        const bool valid = n < fd.N;
        const int q0 = valid ? n / kHop : 0;
        const int want = valid ? (fd.tq_row0 + q0) : -2 - fd.ta_row0;
        if (want != key) {
            key = want;
            const size_t ra = (size_t)(fd.ta_row0 + (valid ? q0 : fd.T)) * kRnn + j;
            const float4 x1 = __ldg(TA1 + ra), x2 = __ldg(TA2 + ra);
            ta[0] = x1.x; ta[1] = x1.y; ta[2] = x1.z; ta[3] = x1.w; ta[4] = x2.x; ta[5] = x2.y; ta[6] = x2.z; ta[7] = x2.w;
            if (valid) {
#pragma unroll
                for (int d = 0; d < kTaps; ++d) {
                    const size_t rq = (size_t)(fd.tq_row0 + q0 + d) * kRnn + j;
                    const float4 q1 = __ldg(TQ1 + rq), q2 = __ldg(TQ2 + rq);
                    tq[d][0] = q1.x; tq[d][1] = q1.y; tq[d][2] = q1.z; tq[d][3] = q1.w; tq[d][4] = q2.x; tq[d][5] = q2.y; tq[d][6] = q2.z;
                }
            }
        }
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = ta[i];
        if (valid) {
            const float* cf = coef + (n - q0 * kHop) * kTaps;
#pragma unroll
            for (int d = 0; d < kTaps; ++d) {
                const float w = cf[d];
                if (w != 0.f) {
#pragma unroll
                    for (int i = 0; i < 7; ++i) a[i] = fmaf(w, tq[d][i], a[i]);
                }
            }
        }
        float* out = CS + (((size_t)g * cs_steps + (t % cs_steps)) * Ng + row) * 4096 + j;
        // plain write-back stores: the ring is meant to LIVE in L2 (a slot is rewritten every cs_steps steps); the streaming
        // hint (st.cs, evict-first) of the first version sent half of it to DRAM (ncu: 9.6 GB written per 60 s utterance)
        out[0 * 512] = a[0]; out[1 * 512] = a[1]; out[2 * 512] = a[2];      // c1 r, z, n
        out[3 * 512] = a[4]; out[4 * 512] = a[5]; out[5 * 512] = a[6];      // c2 r, z, n
        out[6 * 512] = a[3]; out[7 * 512] = a[7];                            // c3 (fc1), c4 (fc2)
    }
}

// The RAW sampler role (T5).  Thread (q, cs): fold 32 q + lane, classes kCls cs .. kCls cs + kCls - 1 of this CTA's 4 kCls.
// The four slices of a fold meet in shared memory ({max, sum} per slice), the CTAs of the group exchange {max, sum} of their
// classes as tagged words (double-buffered by step parity), every CTA forms the same normaliser Z and threshold u Z, and exactly
// one thread finds the class: the sample goes out as the tagged word T1, T2 and T3 wait for.  The station is bound by the
// special-function unit, so every class is exponentiated ONCE (against its slice's maximum; rescaled by one factor per slice
// afterwards) and the uniform is drawn before the wait.
template <bool kTrace, int kCls>
__device__ __forceinline__ void sampler_role(const RsParams& p, Ctl* ctl, uint32_t ctl_s, uint8_t* smem, uint32_t cst_s, const Lane& L, uint4* X, int g,
                                             int cta, int S, const FoldDesc& fd, size_t srow, unsigned long long* xw, uint2 key) {
#define MAT(m, t) (X + ((size_t)(m) * kRsBufs + ((t) & (kRsBufs - 1))) * kMatChunks)
#define GEN(t) ((((t) / kRsBufs) & 1) != 0)
    constexpr int kQC = 4 * kCls;                                        // classes of this CTA
    const int nq = p.n_samplers;
    float* pm = reinterpret_cast<float*>(smem + kNoiseOfs);          // [4 slices][128 folds] max, then sums
    float* ps = pm + 4 * 128;
    const uint32_t bias_s = cst_s + 4u * kCls * L.cs;
    for (int t = 0; t < S && !warp_aborted(ctl); ++t) {
        trace<kTrace>(p, t, 0);
        const uint32_t tag = (uint32_t)t + 1u;
        const float u = u01(philox4x32_10(make_uint4((uint32_t)t, (uint32_t)fd.fold, (uint32_t)fd.utt, 0u), key).x);      // (before the wait)
        ingest<kTrace>(p, ctl, ctl_s, L, MAT(MF2, t), kTagS, GEN(t), -1, t, 1, nullptr, 0u);
        wait_mbar<true>(p, ctl, ctl_s + kBarD, (uint32_t)t & 1u);
        tcgen05_fence_after();
        trace<kTrace>(p, t, 4);
        ctrace<kTrace>(p, t, 0);
        if (L.wlive) {
            float l[kCls];
#pragma unroll
            for (int i = 0; i < kCls; i += 8) tmem_ld8(L.tlane + kColD0 + kCls * L.cs + i, l + i);
            tmem_ld_wait();
            tcgen05_fence_before();
            ctrace<kTrace>(p, t, 1);
            float m = -INFINITY;
#pragma unroll
            for (int i = 0; i < kCls; i += 4) {
                const float4 b4 = lds4(bias_s + 4u * i);
                l[i] += b4.x; l[i + 1] += b4.y; l[i + 2] += b4.z; l[i + 3] += b4.w;
                m = fmaxf(fmaxf(m, fmaxf(l[i], l[i + 1])), fmaxf(l[i + 2], l[i + 3]));
            }
            if (p.logits_out && L.live) {
                float* lo = p.logits_out + (srow + t) * p.C + kQC * cta + kCls * L.cs;
#pragma unroll
                for (int i = 0; i < kCls; ++i) lo[i] = l[i];
            }
            float ssum = 0.f;
#pragma unroll
            for (int i = 0; i < kCls; ++i) { l[i] = ex2_ftz(kL2E * (l[i] - m)); ssum += l[i]; }      // l[] = exp(logit - slice max) from here on
            pm[L.cs * 128 + L.row] = m;
            ps[L.cs * 128 + L.row] = ssum;
            ctrace<kTrace>(p, t, 2);
            asm volatile("bar.sync %0, 128;" ::"r"(3 + L.q) : "memory");       // the four slice warps of my lane quadrant
            ctrace<kTrace>(p, t, 3);
            float sm[4], ss[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) { sm[c] = pm[c * 128 + L.row]; ss[c] = ps[c * 128 + L.row]; }
            const float mc = fmaxf(fmaxf(sm[0], sm[1]), fmaxf(sm[2], sm[3]));
            float sc = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) { ss[c] *= ex2_ftz(kL2E * (sm[c] - mc)); sc += ss[c]; }       // slice masses against the CTA's maximum
            // ONE warp per lane quadrant (slice 0) talks to the other CTAs: it publishes this CTA's {max, sum} (one 16-byte word with
            // the step tag in both upper halves), gathers the others', forms the global maximum M, the normaliser Z, the threshold
            // u Z, the owner CTA and -- if that is this CTA -- the owner slice, and leaves the verdict in shared memory; the other
            // three slice warps sleep on the quadrant's barrier meanwhile (16 warps polling the same 16 KB of words from 16 CTAs
            // delayed the words themselves: 2-3 us per step, measured).
            float* vd = ps + 4 * 128;                                    // verdict [4][128 folds]: owner slice (or -1), its base mass, M, threshold
            if (L.cs == 0) {
                uint4* pw = reinterpret_cast<uint4*>(p.bP) + (((size_t)g * 2 + (t & 1)) * 128 + L.row_ld) * kRsMaxSamplers;
                if (L.live) st_chunk(pw + cta, make_uint4(__float_as_uint(mc), tag, __float_as_uint(sc), tag));
                ctrace<kTrace>(p, t, 4);
                float mq[kRsMaxSamplers], zq[kRsMaxSamplers];
                uint32_t pend = 0u;
#pragma unroll
                for (int q2 = 0; q2 < kRsMaxSamplers; ++q2) {
                    mq[q2] = -INFINITY; zq[q2] = 0.f;
                    if (q2 == cta) { mq[q2] = mc; zq[q2] = sc; }
                    else if (q2 < nq) pend |= 1u << q2;
                }
                {   // every pass requests ALL the missing partials at once (a pass is one L2 round trip; one after the other cost a
                    // round trip EACH, because the first requests always leave before the other CTAs have published)
                    long long t0 = 0;
                    int spins = 0;
                    while (pend) {
                        uint4 w[kRsMaxSamplers];
#pragma unroll
                        for (int q2 = 0; q2 < kRsMaxSamplers; ++q2)
                            if ((pend >> q2) & 1u) w[q2] = ld_chunk(pw + q2);
#pragma unroll
                        for (int q2 = 0; q2 < kRsMaxSamplers; ++q2)
                            if (((pend >> q2) & 1u) && w[q2].y == tag && w[q2].w == tag) {
                                mq[q2] = __uint_as_float(w[q2].x); zq[q2] = __uint_as_float(w[q2].z);
                                pend &= ~(1u << q2);
                            }
                        if (pend && RS_SPIN_CHECK(255)) break;
                    }
                }
                ctrace<kTrace>(p, t, 5);
                float M = mq[0];
#pragma unroll
                for (int q2 = 1; q2 < kRsMaxSamplers; ++q2) M = fmaxf(M, mq[q2]);
                // masses against the global maximum (same order and arithmetic in every CTA); owner CTA = the first whose cumulative
                // mass reaches the threshold (the last one if rounding left the total short)
                float before = 0.f, Z = 0.f;
#pragma unroll
                for (int q2 = 0; q2 < kRsMaxSamplers; ++q2) {
                    if (q2 < nq) {
                        zq[q2] *= ex2_ftz(kL2E * (mq[q2] - M));
                        if (q2 == cta) before = Z;
                        Z += zq[q2];
                    }
                }
                const float thr = u * Z;
                float cum = 0.f;
                int qs = nq - 1;
                bool got = false;
#pragma unroll
                for (int q2 = 0; q2 < kRsMaxSamplers; ++q2) {
                    if (q2 < nq) {
                        cum += zq[q2];
                        if (!got && thr <= cum) { qs = q2; got = true; }
                    }
                }
                int cs_own = -1;
                float base_own = 0.f;
                if (qs == cta) {        // owner slice inside the CTA: the first whose cumulative mass reaches the threshold (the last if none)
                    const float fc = ex2_ftz(kL2E * (mc - M));
                    float base = before;
                    bool found = false;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float e = ss[c] * fc;
                        if (!found && (thr <= base + e || c == 3)) { cs_own = c; base_own = base; found = true; }
                        base += e;
                    }
                }
                vd[0 * 128 + L.row] = __int_as_float(cs_own); vd[1 * 128 + L.row] = base_own; vd[2 * 128 + L.row] = M; vd[3 * 128 + L.row] = thr;
            }
            ctrace<kTrace>(p, t, 6);
            asm volatile("bar.sync %0, 128;" ::"r"(3 + L.q) : "memory");
            if (__float_as_int(vd[0 * 128 + L.row]) == L.cs && L.live) {
                const float base_own = vd[1 * 128 + L.row], M = vd[2 * 128 + L.row], thr = vd[3 * 128 + L.row];
                const float fs = ex2_ftz(kL2E * (m - M));
                float acc = base_own;
                int k = kCls - 1;
                bool hit = false;
#pragma unroll
                for (int i = 0; i < kCls; ++i) {
                    acc = fmaf(l[i], fs, acc);
                    if (!hit && acc >= thr) { k = i; hit = true; }
                }
                const float xs = 2.0f * (float)(kQC * cta + kCls * L.cs + k) / ((float)p.C - 1.0f) - 1.0f;      // fatchord_version.py:228
                p.samples[srow + t] = xs;
                ll_store(xw, p.forced ? p.forced[srow + t] : xs, tag);
            }
        } else {
            tcgen05_fence_before();
        }
        trace<kTrace>(p, t, 5);
        ctrace<kTrace>(p, t, 7);
    }
#undef MAT
#undef GEN
}

}  // namespace

template <bool kTrace, bool kInl>
__global__ void __launch_bounds__(NT, 1) wrnn_loop_rs_kernel(const __grid_constant__ RsParams p) {
    // (no static shared memory in this kernel: the dynamic window starts 1024-byte aligned, which SWIZZLE_128B tiles need;
    //  checked below instead of rounded up -- the round-up was recomputed at every use of a shared address)
    extern __shared__ __align__(1024) uint8_t smem[];
    Ctl* ctl = reinterpret_cast<Ctl*>(smem + kCtlOfs);
    float* cst = reinterpret_cast<float*>(smem + kConstOfs);
    const uint32_t smem_s = smem_u32(smem), ctl_s = smem_s + kCtlOfs, cst_s = smem_s + kConstOfs;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // Which SMs a group lives on is the block scheduler's choice (bid -> SM differs between GPUs and grid sizes) and the step time
    // follows it (measured: up to 2 us of a 15 us RAW step between two B200s running the same binary).  With p.place the logical
    // CTA index is taken from the PHYSICAL SM instead: rank of my %smid among the SMs of this grid (one CTA per SM: the grid is
    // padded to the SM count), rotated by p.rot -- the engine can then choose the layout (engine.cu: rs_calibrate).
    // (compiled in with -DWRNN_RS_PLACEMENT only -- tools/mkvariant.py: a logical index that is not blockIdx.x costs the product
    //  kernel ~100 bytes of spills, because everything derived from it can no longer be rematerialised from the special register)
#ifndef WRNN_RS_PLACEMENT
    const int bid = (int)blockIdx.x;
#else
    int bid = (int)blockIdx.x;
    if (p.place) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.place + blockIdx.x), "r"(smid + 1u) : "memory");
        int below = 0;
        if (tid < (int)gridDim.x) {
            uint32_t v = 0;
            long long t0 = clock64();
            while (true) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.place + tid) : "memory");
                if (v != 0u || clock64() - t0 > 2000000000ll) break;
            }
            below = (v != 0u && v - 1u < smid) ? 1 : 0;
        }
        bid = (__syncthreads_count(below) + p.rot) % (int)gridDim.x;
    }
#endif
    const int g = bid / p.ctas, rc = bid % p.ctas;
    const int role = rc < kRsT1 ? 0 : (rc < kRsT1 + kRsT2 ? 1 : (rc < kRsT1 + kRsT2 + kRsT3 ? 2 : (rc < kRsCtas ? 3 : 4)));      // 4: RAW sampler (T5)
    const int cta = role == 0 ? rc : (role == 1 ? rc - kRsT1 : (role == 2 ? rc - kRsT1 - kRsT2 : (role == 3 ? rc - kRsT1 - kRsT2 - kRsT3 : rc - kRsCtas)));
    const bool raw = p.mode == 0;                                   // RAW: fc3 + the draw live on the sampler CTAs, T1 only waits for the sample
    const int fold0 = g * p.Ng, nrows = max(0, min(p.Ng, p.B - fold0));
    const int S = p.S;
    const bool expander = bid >= p.G * p.ctas;         // CTAs past the groups produce the conditioning records
    const unsigned int consumers = (unsigned int)(p.G * kRsCtas * NW);   // warps that read every record chunk (the samplers read none)
    if ((smem_s & 1023u) != 0u) {          // never on this toolchain; a misaligned tile would compute garbage silently
        if (tid == 0) atomicExch(p.abort_flag, 1);
        return;
    }

    if (expander && kInl) return;          // (inline conditioning: CTAs past the groups only pad the grid -- see launch_loop_rs)
    if (expander) {
        // =================================== conditioning expander ==========================================================
        // work item = (chunk of kRsChunk steps, fold), chunk-major, so chunks complete in the order the loop consumes them;
        // CS is a ring of cs_steps steps (L2-resident by construction: engine.cu sizes it), refilled behind the loop:
        // cs_done[c] counts the folds of chunk c that are written, cs_consumed[c] the consumer warps that have read it.
        if (tid == 0) ctl->abort_local = 0;
        for (int i = tid; i < kHop * kTaps; i += NT) reinterpret_cast<float*>(smem)[i] = p.coef[i];
        __syncthreads();
        if (warp < NW) {
            const float* coef_s = reinterpret_cast<const float*>(smem);
            const int e_idx = bid - p.G * p.ctas;
            const int nchunks = (S + kRsChunk - 1) / kRsChunk, ring_chunks = p.cs_steps / kRsChunk;
            const long long nitems = (long long)nchunks * p.B;
            int waited = -1;
            for (long long it = e_idx; it < nitems && !aborted_local(ctl); it += p.n_expanders) {
                const int c = (int)(it / p.B), b = (int)(it - (long long)c * p.B);
                if (c >= ring_chunks && c != waited) {          // chunk c overwrites chunk c - ring_chunks: every consumer must be past it
                    waited = c;
                    if (tid == 0) {
                        long long t0 = 0;
                        int spins = 0;
                        while (true) {
                            unsigned int v;
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.cs_consumed + (c - ring_chunks)) : "memory");
                            if (v >= consumers) break;
                            __nanosleep(200);
                            if (RS_SPIN_CHECK(63)) break;
                        }
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
                }
                expand_item_rs(p.TA1, p.TA2, p.TQ1, p.TQ2, coef_s, p.folds[b], b / p.Ng, b % p.Ng, c * kRsChunk, min(S, (c + 1) * kRsChunk),
                               p.cs_steps, p.Ng, p.CSw, tid);
                asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
                if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.cs_done + c) : "memory");
            }
        }
        return;
    }

    // ---- one-time setup: weight tiles, constants, barriers, TMEM ------------------------------------------------------
    {
        const unsigned char* img = role == 4 ? p.w5 + (size_t)cta * ((size_t)p.qcols * 128 * 8)
                                 : role == 0 ? p.w1 + (size_t)cta * (kW1 + 32 * 128 * 8)
                                 : role == 1 ? p.w2 + (size_t)cta * kWEnd
                                 : role == 2 ? p.w3 + (size_t)cta * (kFU * 128 * 8) : p.w4 + (size_t)cta * (kFU * 128 * 8);
        const int bytes = role == 4 ? p.qcols * 128 * 8 : role == 0 ? kW1 + 32 * 128 * 8 : (role == 1 ? kWEnd : kFU * 128 * 8);
        const uint4* src = reinterpret_cast<const uint4*>(img);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        for (int i = tid; i < bytes / 16; i += NT) dst[i] = src[i];
        if (kInl && role <= 2) {      // the role's W_q tile (mel share of the conditioning)
            const int xrows = role == 0 ? 128 : (role == 1 ? 96 : kFU);
            const uint4* xs = reinterpret_cast<const uint4*>((role == 0 ? p.wx1 : role == 1 ? p.wx2 : p.wx3) + (size_t)cta * (xrows * 256));
            uint4* xd = reinterpret_cast<uint4*>(smem + (role == 0 ? kXW1 : role == 1 ? kXW2 : kXW3));
            for (int i = tid; i < xrows * 16; i += NT) xd[i] = xs[i];
        }
        fence_proxy_async_smem();
    }
    if (role == 0) {            // [v1 r,z,n | b_hn1] x 32 units, fc3 bias (32)
        if (tid < 128) { const int a = tid >> 5, u = tid & 31, j = 32 * cta + u; cst[tid] = a < 3 ? p.v1[a * kRnn + j] : p.bhn1[j]; }
        else if (tid < 160) cst[tid] = (tid - 128) < 30 ? p.bfc3[tid - 128] : 0.f;
    } else if (role == 1) {     // [v2 r,z,n | b_hn2] x 32 units
        if (tid < 128) { const int a = tid >> 5, u = tid & 31, j = 32 * cta + u; cst[tid] = a < 3 ? p.v2[a * kRnn + j] : p.bhn2[j]; }
    } else if (role == 2) {     // v3 x 64 units
        if (tid < kFU) cst[tid] = p.v3[kFU * cta + tid];
    } else if (role == 4) {     // fc3 bias of my 128 classes
        if (tid < p.qcols) cst[tid] = p.bfc3[p.qcols * cta + tid];
    }
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&ctl->abar[i], NW);
        mbar_init(&ctl->dbar[0], 1); mbar_init(&ctl->dbar[1], 1);
        mbar_init(&ctl->ebar, NW);
        mbar_init(&ctl->xbar, NW);
        ctl->abort_local = 0;
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&ctl->tmem, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = ctl->tmem;
    dbg<kTrace>(p, 0, 1);
    // (register hand-over, first statement of each branch below: the service warpgroup gives its registers to the
    //  16 ingest / epilogue warps)

    uint4* const X = p.X + (size_t)g * kMats * kRsBufs * kMatChunks;
#define MAT(m, t) (X + ((size_t)(m) * kRsBufs + ((t) & (kRsBufs - 1))) * kMatChunks)
#define GEN(t) ((((t) / kRsBufs) & 1) != 0)
    static_assert(kRsBufs == 2, "MAT / GEN assume two buffers");

    if (warp >= NW) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsSvc));     // (one instruction for the whole warpgroup)
        dbg<kTrace>(p, 0, 2);
        if (warp == NW) {
        // =================================== MMA issuer ====================================================================
        uint32_t n_ingest = 0;
        // (xofs != 0: the job's accumulator starts with the inline-conditioning slab W_q m, operand generation x_par)
        auto job = [&](uint32_t N, uint32_t dcol, uint32_t wofs, uint32_t done_bar, bool wait_e, uint32_t e_par, int tt, int ev, uint32_t xofs = 0u,
                       uint32_t NX = 0u, uint32_t x_par = 0u) {
            mma_job<kTrace>(p, ctl, ctl_s, smem_s + wofs, N, tmem + dcol, tmem + kColA, n_ingest & 1u, done_bar, wait_e, e_par, tt, ev,
                            (kInl && xofs) ? smem_s + xofs : 0u, NX, x_par);
            ++n_ingest;
        };
        if (role == 0) {
            if (kInl) {      // "job -1": D0 = W_q1 m(0) (no recurrent share before the first step); completes phase 0 of dbar[1]
                wait_mbar(p, ctl, ctl_s + kBarX, 0u);
                tcgen05_fence_after();
                mma_ext(smem_s + kXW1, 128u, tmem + kColD0, tmem + kColX);
                if (elect_one()) umma_commit_s(ctl_s + kBarD + 8u);
                __syncwarp();
            }
            for (int t = 0; t <= S && !warp_aborted(ctl); ++t) {
                if (t > 0 && !raw) job(32u, kColD1T1, kW1, ctl_s + kBarD, false, 0, t, 9);               // fc3 f2(t-1)
                if (t < S) job(96u, kColD0, kW0, ctl_s + kBarD + 8u, true, (uint32_t)t & 1u, t, 11, kXW1, 128u, (uint32_t)(t + 1) & 1u);    // W_hh1 h1(t) [+ W_q1 m(t+1)]
            }
        } else if (role == 1) {
            for (int t = 0; t < S && !warp_aborted(ctl); ++t) {
                job(96u, kColD0, kW0, ctl_s + kBarD, false, 0, t, 9, kXW2, 96u, (uint32_t)t & 1u);      // W_ih2a h1(t) [+ W_q2 m(t)]
                job(96u, kColD1, kW1, ctl_s + kBarD + 8u, true, (uint32_t)t & 1u, t, 11);               // W_hh2 h2(t)
            }
        } else if (role == 4) {
            for (int t = 0; t < S && !warp_aborted(ctl); ++t) job((uint32_t)p.qcols, kColD0, kW0, ctl_s + kBarD, false, 0, t, 9);   // fc3 slice f2(t)
        } else if (role == 2) {
            for (int t = 0; t < S && !warp_aborted(ctl); ++t) job((uint32_t)kFU, kColD0, kW0, ctl_s + kBarD, false, 0, t, 9, kXW3, (uint32_t)kFU, (uint32_t)t & 1u);   // fc1a s2(t) [+ W_q3 m(t)]
        } else {
            for (int t = 0; t < S && !warp_aborted(ctl); ++t) job((uint32_t)kFU, kColD0, kW0, ctl_s + kBarD, false, 0, t, 9);          // fc2 f1(t)
        }
        }
    } else {
        // =================================== ingest + epilogue warps ========================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpi));
        dbg<kTrace>(p, 0, 3);
        Lane L;
        L.q = warp & 3; L.cs = warp >> 2; L.lane = lane; L.row = 32 * L.q + lane; L.live = L.row < nrows;
        L.wlive = 32 * L.q < nrows;               // a warp whose 32 rows are all padding skips loads and arithmetic (frees issue slots)
        L.row_ld = L.live ? L.row : max(nrows - 1, 0);
        L.nlive_threads = 4 * 32 * ((nrows + 31) / 32);
        L.tlane = tmem + ((uint32_t)(32 * L.q) << 16);
        const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
        const FoldDesc fd = p.folds[fold0 + L.row_ld];
        const size_t srow = (size_t)(fold0 + L.row_ld) * S;                  // my fold's row of samples / forced
        unsigned long long* const xw = p.bX + (size_t)g * 128 + L.row_ld;
        const float* const csrow = p.CS + ((size_t)g * p.cs_steps * p.Ng + L.row_ld) * 4096;   // + (t % cs_steps) * Ng * 4096
        const size_t cs_step = (size_t)p.Ng * 4096;
        // the expanders run a few chunks ahead: before the first step of a chunk, lane 0 acquires its counter
        auto cs_wait = [&](int t) {
            if (p.cs_done && (t & (kRsChunk - 1)) == 0) {
                if (lane == 0) {
                    long long t0 = 0;
                    int spins = 0;
                    while (true) {
                        unsigned int v;
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.cs_done + t / kRsChunk) : "memory");
                        if (v >= (unsigned int)p.B) break;
                        if (RS_SPIN_CHECK(63)) break;
                    }
                }
                __syncwarp();
            }
        };
        // ... and after the last step of a chunk (its records are in registers and used) the warp hands the ring slot back
        auto cs_release = [&](int t) {
            if (p.cs_done && ((t & (kRsChunk - 1)) == kRsChunk - 1 || t == S - 1)) {
                __syncwarp();
                if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.cs_consumed + t / kRsChunk) : "memory");
            }
        };
        // three gate planes of my 8 units from a conditioning record: a[0..7] r, a[8..15] z, a[16..23] n
        auto load_rec24 = [&](const float* c, float* a) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4 lo = __ldcg(reinterpret_cast<const float4*>(c + k * 512)), hi = __ldcg(reinterpret_cast<const float4*>(c + k * 512) + 1);
                a[8 * k + 0] = lo.x; a[8 * k + 1] = lo.y; a[8 * k + 2] = lo.z; a[8 * k + 3] = lo.w;
                a[8 * k + 4] = hi.x; a[8 * k + 5] = hi.y; a[8 * k + 6] = hi.z; a[8 * k + 7] = hi.w;
            }
        };
        // ---- inline conditioning (kInl): no per-sample records, no expanders.  A record is the per-FRAME row FR[frame] (aux share +
        // biases, fp32, one 16 KB row per frame: it changes every 200 steps) and the mel share is W_q . m(n) inside the role's MMA,
        // m(n) = the upsampled mel of sample n (80 fp16, table M16).  (c_frame, c_phase) = position of sample n0 + t.
        int c_frame = fd.n0 / kHop, c_phase = fd.n0 - c_frame * kHop;
        auto rec_row = [&]() { return p.FR + (size_t)(fd.ta_row0 + min(c_frame, fd.T)) * 4096; };     // row T: bias only (fold tail padding, Q9)
        auto m_row = [&](int frame, int phase) { return frame < fd.T ? (long long)(fd.tq_row0 + frame) * kHop + phase : p.m16_zero; };
        auto advance = [&]() { if (++c_phase == kHop) { c_phase = 0; ++c_frame; } };
        // my warp's K step(s) of row `mrow` -> TMEM columns kColX + 8 k (warp cs: k = cs; cs 0 also k = 4); every warp arrives
        auto ext_store = [&](long long mrow) {
            if (L.wlive) {
                const uint4* src = reinterpret_cast<const uint4*>(p.M16 + mrow * kFeat);
                const uint4 v0 = __ldg(src + 2 * L.cs), v1 = __ldg(src + 2 * L.cs + 1);
                uint4 v2 = make_uint4(0u, 0u, 0u, 0u), v3 = v2;
                if (L.cs == 0) { v2 = __ldg(src + 8); v3 = __ldg(src + 9); }
                tmem_st8(L.tlane + kColX + 8u * (uint32_t)L.cs, v0, v1);
                if (L.cs == 0) tmem_st8(L.tlane + kColX + 32u, v2, v3);
                tmem_st_wait();
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_s(ctl_s + kBarX);
        };
        if (role == 0) {
            // ---- T1: fc3 + draw of step t-1, GRU1 of step t, then the recurrent product for step t+1 -----------------------
            float h1[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) h1[i] = 0.f;
            const uint32_t v1_s = cst_s + 32u * L.cs;            // [a][32 units]: + 128 a bytes
            const uint32_t sbias_s = cst_s + 512u;
            const uint32_t nz_s = smem_s + kNoiseOfs + 4u * tid;
            if (kInl) ext_store(m_row(c_frame, c_phase));        // m(0) for "job -1"
            for (int t = 0; t <= S && !warp_aborted(ctl); ++t) {
                trace<kTrace>(p, t, 0);
                // everything that does not need f2(t-1) happens BEFORE the wait for it: the conditioning record, the mixture
                // noise, and the recurrent accumulator W_hh1 h1(t-1) (complete since the off-path job of the previous step)
                // folded into the record: a[] = (c_r + gh_r, c_z + gh_z, c_n), bn[] = gh_n + b_hn
                float a[24], bn[8];
                if (!kInl && t < S) cs_wait(t);
                if (t < S && L.wlive) load_rec24((kInl ? rec_row() : csrow + (size_t)(t % p.cs_steps) * cs_step) + 32 * cta + 8 * L.cs, a);
                if (raw && t == S) break;           // (RAW: the sampler CTAs draw the last sample; nothing is left to do here)
                if (t > 0 && L.wlive && !raw) mol_noise(nz_s, sbias_s, (uint32_t)(t - 1), (uint32_t)fd.fold, (uint32_t)fd.utt, key);
                if (kInl || t > 0) {      // (kInl: phase t of dbar[1] is job t-1, phase 0 the conditioning-only "job -1")
                    wait_mbar(p, ctl, ctl_s + kBarD + 8u, (uint32_t)(kInl ? t : t - 1) & 1u);      // recurrent job of step t-1: D0 complete, A free
                    tcgen05_fence_after();
                }
                if (t < S) {
                    if (L.wlive) {
                        lds8(v1_s + 384u, bn);
                        if (kInl || t > 0) {
                            float gh[24];
                            tmem_ld8(L.tlane + kColD0 + 0 + 8 * L.cs, gh); tmem_ld8(L.tlane + kColD0 + 32 + 8 * L.cs, gh + 8);
                            tmem_ld8(L.tlane + kColD0 + 64 + 8 * L.cs, gh + 16);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 8; ++i) { a[i] += gh[i]; a[8 + i] += gh[8 + i]; bn[i] += gh[16 + i]; }
                            if (kInl) {       // columns 96..127: the mel share of the candidate's input side (kept apart from W_hn h)
                                tmem_ld8(L.tlane + kColD0 + 96 + 8 * L.cs, gh);
                                tmem_ld_wait();
#pragma unroll
                                for (int i = 0; i < 8; ++i) a[16 + i] += gh[i];
                            }
                        }
                    }
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_s(ctl_s + kBarE);
                }
                float x = 0.f;
                if (t > 0 && raw) {
                    // RAW: the sample of step t-1 comes from the group's sampler CTAs as a tagged word (no matrix on this station's path)
                    if (L.wlive) {
                        long long t0 = 0;
                        int spins = 0;
                        unsigned long long w = ll_load(xw);
                        while (ll_tag(w) != (uint32_t)t) {
                            w = ll_load(xw);
                            if (RS_SPIN_CHECK(255)) break;
                        }
                        x = ll_val(w);
                    }
                    trace<kTrace>(p, t, 4);
                } else if (t > 0) {
                    ingest<kTrace>(p, ctl, ctl_s, L, MAT(MF2, t - 1), kTagS, GEN(t - 1), -1, t, 1, nullptr, 0u);
                    ctrace<kTrace>(p, t, 0);
                    wait_mbar<true>(p, ctl, ctl_s + kBarD, (uint32_t)(t - 1) & 1u);
                    tcgen05_fence_after();
                    trace<kTrace>(p, t, 4);
                    ctrace<kTrace>(p, t, 1);
                    if (L.wlive) {
                        float lg[32];
                        tmem_ld8(L.tlane + kColD1T1 + 0, lg); tmem_ld8(L.tlane + kColD1T1 + 8, lg + 8);
                        tmem_ld8(L.tlane + kColD1T1 + 16, lg + 16); tmem_ld8(L.tlane + kColD1T1 + 24, lg + 24);
                        tmem_ld_wait();
                        ctrace<kTrace>(p, t, 2);
                        const float xs = mol_draw(lg, nz_s, sbias_s);
                        x = xs;
                        if (p.forced) x = p.forced[srow + t - 1];
                        if (cta == 0 && L.cs == 0 && L.live) {
                            p.samples[srow + t - 1] = xs;
                            ll_store(xw, x, (uint32_t)t);
                            if (p.logits_out)
                                for (int i = 0; i < 30; ++i) p.logits_out[(srow + t - 1) * 30 + i] = lg[i] + lds1(sbias_s + 4u * i);
                        }
                    }
                    tcgen05_fence_before();
                }
                if (t == S) break;
                ctrace<kTrace>(p, t, 3);
                if (L.wlive) {
                    float kr[8];
                    lds8(v1_s, kr);
#pragma unroll
                    for (int i = 0; i < 8; ++i) a[i] = fmaf(kr[i], x, a[i]);
                    lds8(v1_s + 128u, kr);
#pragma unroll
                    for (int i = 0; i < 8; ++i) a[8 + i] = fmaf(kr[i], x, a[8 + i]);
                    lds8(v1_s + 256u, kr);
#pragma unroll
                    for (int i = 0; i < 8; ++i) a[16 + i] = fmaf(kr[i], x, a[16 + i]);
                    gru8(a, a + 8, a + 16, bn, h1);
                    ctrace<kTrace>(p, t, 5);
                    if (L.live) publish8(MAT(MH1, t), 4 * cta + L.cs, L.row, h1, kTagE, GEN(t) ? kTagE : 0u);
                }
                trace<kTrace>(p, t, 5);
                ctrace<kTrace>(p, t, 6);
                if (kInl) {       // m(t+1) for the job below (the previous reader of these columns, job t-1, completed before this step)
                    advance();
                    if (L.wlive && L.cs == 1) {      // and the row of step t + 9 on its way into L2 (T1 is the first reader of every row)
                        const __half* pf = p.M16 + m_row(c_frame, c_phase) * kFeat + 8 * kFeat;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
                    }
                    ext_store(m_row(c_frame, c_phase));
                }
                // recurrent product for step t+1: the full h1(t) -> A buffer (the fc3 job has completed: dbar[0] was waited).
                // Off the critical path: wait until the T2 CTAs have read the same lines for the on-path product.
                __nanosleep(p.offpath_delay_ns);
                ingest<kTrace>(p, ctl, ctl_s, L, MAT(MH1, t), kTagE, GEN(t), -1, t, 6, nullptr, 0u);
                if (!kInl) cs_release(t);
                if (cta == 0 && g == 0 && tid == 0 && (t % 100) == 0 && p.progress) {
                    *reinterpret_cast<volatile int*>(p.progress) = t;
                    __threadfence_system();
                }
            }
        } else if (role == 1) {
            // ---- T2: GRU2 of step t from h1(t); publishes h2 and s2 = h1 + h2; then the recurrent product for step t+1 -------
            float h2[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) h2[i] = 0.f;
            const uint32_t v2_s = cst_s + 32u * L.cs;
            for (int t = 0; t < S && !warp_aborted(ctl); ++t) {
                trace<kTrace>(p, t, 0);
                // before the wait for h1(t): the record and the recurrent accumulator W_hh2 h2(t-1), folded together (see T1)
                float a[24], bn[8];
                if (kInl) ext_store(m_row(c_frame, c_phase));     // m(t): the on-path job starts with W_q2 m(t)
                else cs_wait(t);
                if (L.wlive) load_rec24((kInl ? rec_row() : csrow + (size_t)(t % p.cs_steps) * cs_step) + 3 * 512 + 32 * cta + 8 * L.cs, a);
                if (t > 0) {
                    wait_mbar(p, ctl, ctl_s + kBarD + 8u, (uint32_t)(t - 1) & 1u);
                    tcgen05_fence_after();
                }
                if (L.wlive) {
                    lds8(v2_s + 384u, bn);
                    if (t > 0) {
                        float gh[24];
                        tmem_ld8(L.tlane + kColD1 + 0 + 8 * L.cs, gh); tmem_ld8(L.tlane + kColD1 + 32 + 8 * L.cs, gh + 8);
                        tmem_ld8(L.tlane + kColD1 + 64 + 8 * L.cs, gh + 16);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) { a[i] += gh[i]; a[8 + i] += gh[8 + i]; bn[i] += gh[16 + i]; }
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_s(ctl_s + kBarE);
                const IngestOut io = ingest<kTrace>(p, ctl, ctl_s, L, MAT(MH1, t), kTagE, GEN(t), 4 * cta + L.cs, t, 1, t > 0 ? xw : nullptr, (uint32_t)t);
                const float x = io.x;
                dbg<kTrace>(p, t, 0x30);
                ctrace<kTrace>(p, t, 0);
                if (L.wlive) {              // the sample's rank-1 term while the last MMAs run
                    float kr[8];
                    lds8(v2_s, kr);
#pragma unroll
                    for (int i = 0; i < 8; ++i) a[i] = fmaf(kr[i], x, a[i]);
                    lds8(v2_s + 128u, kr);
#pragma unroll
                    for (int i = 0; i < 8; ++i) a[8 + i] = fmaf(kr[i], x, a[8 + i]);
                    lds8(v2_s + 256u, kr);
#pragma unroll
                    for (int i = 0; i < 8; ++i) a[16 + i] = fmaf(kr[i], x, a[16 + i]);
                }
                wait_mbar<true>(p, ctl, ctl_s + kBarD, (uint32_t)t & 1u);
                tcgen05_fence_after();
                dbg<kTrace>(p, t, 0x31);
                trace<kTrace>(p, t, 4);
                ctrace<kTrace>(p, t, 1);
                if (L.wlive) {
                    float pb[24];
                    tmem_ld8(L.tlane + kColD0 + 0 + 8 * L.cs, pb); tmem_ld8(L.tlane + kColD0 + 32 + 8 * L.cs, pb + 8);
                    tmem_ld8(L.tlane + kColD0 + 64 + 8 * L.cs, pb + 16);
                    tmem_ld_wait();
                    tcgen05_fence_before();
                    ctrace<kTrace>(p, t, 4);
#pragma unroll
                    for (int i = 0; i < 24; ++i) a[i] += pb[i];
                    gru8(a, a + 8, a + 16, bn, h2);
                    const float2 e0 = unpack2(io.extra.x), e1 = unpack2(io.extra.y), e2 = unpack2(io.extra.z), e3 = unpack2(io.extra.w);
                    const float h1o[8] = {e0.x, e0.y, e1.x, e1.y, e2.x, e2.y, e3.x, e3.y};
                    float s2[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) s2[i] = fminf(fmaxf(h1o[i] + h2[i], -1.9990234375f), 1.9990234375f);
                    ctrace<kTrace>(p, t, 5);
                    if (L.live) {
                        publish8(MAT(MS2, t), 4 * cta + L.cs, L.row, s2, kTagE, GEN(t) ? kTagE : 0u);
                        publish8(MAT(MH2, t), 4 * cta + L.cs, L.row, h2, kTagE, GEN(t) ? kTagE : 0u);
                    }
                }
                trace<kTrace>(p, t, 5);
                ctrace<kTrace>(p, t, 6);
                ingest<kTrace>(p, ctl, ctl_s, L, MAT(MH2, t), kTagE, GEN(t), -1, t, 6, nullptr, 0u);
                if (kInl) advance(); else cs_release(t);
            }
        } else if (role == 4) {
            // ---- T5 (RAW): my slice of fc3 (p.qcols = 64 or 128 classes) on f2(t), the soft-max partials, and -- in the CTA whose
            // classes contain the threshold -- the inverse-CDF draw (rule: oracle sample_raw; fatchord_version.py:224-230: first k
            // with cdf[k] >= u, one Philox uniform per (step, fold)).
            if (p.qcols == 64) sampler_role<kTrace, 16>(p, ctl, ctl_s, smem, cst_s, L, X, g, cta, S, fd, srow, xw, key);
            else sampler_role<kTrace, 32>(p, ctl, ctl_s, smem, cst_s, L, X, g, cta, S, fd, srow, xw, key);
        } else {
            // ---- T3 / T4: fc1 on s2(t) (+ the sample's rank-1 term) / fc2 on f1(t); ReLU; publish ------------------------------
            const bool fc1 = role == 2;
            constexpr int kPU = kFU / 4;                       // units per thread: 16 (8 CTAs per FC role) or 8 (16 CTAs)
            const uint32_t v3_s = cst_s + 4u * kPU * L.cs;
            for (int t = 0; t < S && !warp_aborted(ctl); ++t) {
                trace<kTrace>(p, t, 0);
                float cc[kPU];
                if (kInl) { if (fc1) ext_store(m_row(c_frame, c_phase)); }     // m(t): fc1's job starts with W_q3 m(t)
                else cs_wait(t);
                if (L.wlive) {
                    const float4* c = reinterpret_cast<const float4*>((kInl ? rec_row() : csrow + (size_t)(t % p.cs_steps) * cs_step) + (fc1 ? 6 : 7) * 512 + kFU * cta + kPU * L.cs);
#pragma unroll
                    for (int i = 0; i < kPU / 4; ++i) { const float4 q4 = __ldcg(c + i); cc[4 * i] = q4.x; cc[4 * i + 1] = q4.y; cc[4 * i + 2] = q4.z; cc[4 * i + 3] = q4.w; }
                }
                const float x = ingest<kTrace>(p, ctl, ctl_s, L, MAT(fc1 ? MS2 : MF1, t), fc1 ? kTagE : kTagS, GEN(t), -1, t, 1,
                                               (fc1 && t > 0) ? xw : nullptr, (uint32_t)t).x;
                if (fc1 && L.wlive) {
#pragma unroll
                    for (int i = 0; i < kPU / 4; ++i) {
                        const float4 v = lds4(v3_s + 16u * i);
                        cc[4 * i] = fmaf(v.x, x, cc[4 * i]); cc[4 * i + 1] = fmaf(v.y, x, cc[4 * i + 1]);
                        cc[4 * i + 2] = fmaf(v.z, x, cc[4 * i + 2]); cc[4 * i + 3] = fmaf(v.w, x, cc[4 * i + 3]);
                    }
                }
                wait_mbar<true>(p, ctl, ctl_s + kBarD, (uint32_t)t & 1u);
                tcgen05_fence_after();
                trace<kTrace>(p, t, 4);
                if (L.wlive) {
                    float d[kPU];
#pragma unroll
                    for (int i = 0; i < kPU / 8; ++i) tmem_ld8(L.tlane + kColD0 + kPU * L.cs + 8 * i, d + 8 * i);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < kPU; ++i) d[i] = fmaxf(d[i] + cc[i], 0.f);
                    if (L.live) {
                        uint4* m = MAT(fc1 ? MF1 : MF2, t);
                        const uint32_t want = GEN(t) ? kTagS : 0u;
#pragma unroll
                        for (int i = 0; i < kPU / 8; ++i) publish8(m, (kFU / 8) * cta + (kPU / 8) * L.cs + i, L.row, d + 8 * i, kTagS, want);
                    }
                }
                tcgen05_fence_before();
                trace<kTrace>(p, t, 5);
                if (kInl) advance(); else cs_release(t);
            }
        }
    }
#undef MAT
#undef GEN
    // ---- teardown ---------------------------------------------------------------------------------------------------
    dbg<kTrace>(p, 0, 0xFF);
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// per-sample conditioning records of the role-specialised loop: CS[group][step][fold][8][512] fp32
// (c1 r,z,n | c2 r,z,n | c3 | c4), same interpolation and FMA order as cond_expand.cuh.  grid = (folds, step blocks),
// 512 threads = hidden units.
__global__ void __launch_bounds__(512) expand_cond_rs_kernel(const float4* __restrict__ TA1, const float4* __restrict__ TA2,
                                                             const float4* __restrict__ TQ1, const float4* __restrict__ TQ2,
                                                             const float* __restrict__ coef, const FoldDesc* __restrict__ folds, int S, int Ng,
                                                             int cs_steps, int steps_per_block, float* __restrict__ CS) {
    const int b = blockIdx.x;
    const int t0 = blockIdx.y * steps_per_block;
    expand_item_rs(TA1, TA2, TQ1, TQ2, coef, folds[b], b / Ng, b % Ng, t0, min(S, t0 + steps_per_block), cs_steps, Ng, CS, threadIdx.x);
}

// Tables of the inline-conditioning form (kInl), one block per row of the utterances' common row space (T + 4 rows each):
//   FR[row][8][512]     = the aux share + biases of a record (TA1 / TA2 in plane order c1 r,z,n | c2 r,z,n | c3 | c4): frames and the bias-only row T
//   M16[row*200+ph][80] = fp16 of the upsampled mel of sample (frame = row - tq_row0, phase ph): sum_d coef[ph][d] melpad[frame + d]
//                         (same taps and order as the per-frame tables of cond.cu; rows of frames >= T are never read)
// and the all-zero row M16[rows * 200] (+ slack for the prefetch ahead).
__global__ void __launch_bounds__(256) rs_inline_tables_kernel(const float4* __restrict__ TA1, const float4* __restrict__ TA2, const float* __restrict__ mel,
                                                               const UttDesc* __restrict__ utts, int n_utts, const float* __restrict__ coef, int rows,
                                                               float* __restrict__ FR, __half* __restrict__ M16) {
    __shared__ float mp[kTaps][kFeat];
    __shared__ float cf[kHop * kTaps];
    const int row = blockIdx.x, tid = threadIdx.x;
    if (row >= rows) {          // the zero rows behind the table
        for (int i = tid; i < 16 * kFeat; i += 256) M16[(size_t)rows * kHop * kFeat + i] = __float2half_rn(0.f);
        return;
    }
    int lo = 0, hi = n_utts - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (utts[mid].tq_row0 <= row) lo = mid; else hi = mid - 1; }
    const UttDesc u = utts[lo];
    const int f = row - u.tq_row0;
    for (int i = tid; i < kRnn; i += 256) {
        const float4 a = TA1[(size_t)row * kRnn + i], b = TA2[(size_t)row * kRnn + i];
        float* o = FR + (size_t)row * 4096 + i;
        o[0] = a.x; o[512] = a.y; o[1024] = a.z; o[1536] = b.x; o[2048] = b.y; o[2560] = b.z; o[3072] = a.w; o[3584] = b.w;
    }
    for (int i = tid; i < kTaps * kFeat; i += 256) {
        const int d = i / kFeat, c = i - d * kFeat, t = f + d - kPad;
        mp[d][c] = (t >= 0 && t < u.T) ? mel[u.mel_off + (long long)c * u.T + t] : 0.f;
    }
    for (int i = tid; i < kHop * kTaps; i += 256) cf[i] = coef[i];
    __syncthreads();
    __half2* out = reinterpret_cast<__half2*>(M16 + (size_t)row * kHop * kFeat);
    for (int i = tid; i < kHop * kFeat / 2; i += 256) {
        const int ph = i / (kFeat / 2), c = 2 * (i - ph * (kFeat / 2));
        float v0 = 0.f, v1 = 0.f;
#pragma unroll
        for (int d = 0; d < kTaps; ++d) { const float w = cf[ph * kTaps + d]; v0 = fmaf(w, mp[d][c], v0); v1 = fmaf(w, mp[d][c + 1], v1); }
        out[i] = __floats2half2_rn(v0, v1);
    }
}
cudaError_t launch_rs_inline_tables(const float4* TA1, const float4* TA2, const float* mel, const UttDesc* utts, int n_utts, const float* coef,
                                    int rows, float* FR, __half* M16, cudaStream_t stream) {
    rs_inline_tables_kernel<<<rows + 1, 256, 0, stream>>>(TA1, TA2, mel, utts, n_utts, coef, rows, FR, M16);
    return cudaGetLastError();
}
size_t loop_rs_ximage_bytes(int role) { return (size_t)(role == 0 ? 128 : (role == 1 ? 96 : kFU)) * 256; }

cudaError_t set_rs_deadline(long long cycles) { return cudaMemcpyToSymbol(g_rs_deadline, &cycles, sizeof(cycles)); }
size_t loop_rs_image_bytes(int role) { return role == 0 ? (size_t)kW1 + 32 * 128 * 8 : (role == 1 ? (size_t)kWEnd : (size_t)kFU * 128 * 8); }
size_t loop_rs_exchange_bytes(int groups) { return (size_t)groups * kMats * kRsBufs * kMatChunks * 16; }

cudaError_t launch_expand_cond_rs(const float4* TA1, const float4* TA2, const float4* TQ1, const float4* TQ2, const float* coef,
                                  const FoldDesc* folds, int B, int S, int Ng, int cs_steps, float* CS, cudaStream_t stream) {
    const int spb = 64;
    dim3 grid(B, (S + spb - 1) / spb);
    expand_cond_rs_kernel<<<grid, 512, 0, stream>>>(TA1, TA2, TQ1, TQ2, coef, folds, S, Ng, cs_steps, spb, CS);
    return cudaGetLastError();
}

cudaError_t launch_loop_rs(const RsParams& p, cudaStream_t stream) {
    // the instrumented instantiation only when a timeline or checkpoints were asked for
    const bool tr = p.trace || p.dbg;
    const void* fn = p.inl ? (tr ? (const void*)wrnn_loop_rs_kernel<true, true> : (const void*)wrnn_loop_rs_kernel<false, true>)
                           : (tr ? (const void*)wrnn_loop_rs_kernel<true, false> : (const void*)wrnn_loop_rs_kernel<false, false>);
    cudaError_t err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (err != cudaSuccess) return err;
    RsParams pp = p;
    void* args[] = {&pp};
    // inline conditioning: n_expanders CTAs that exit at once pad the grid to the SM count, so that the groups sit on the SMs they
    // sat on with the expanders behind them (the step time follows the placement: +-1 us between grid sizes on one GPU, measured)
    const int grid = p.G * p.ctas + ((p.cs_done || p.inl) ? p.n_expanders : 0);
    return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(NT), args, kSmemBytes, stream);
}

}  // namespace wrnn
