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

constexpr int NW = 16;                        // ingest / epilogue warps: warp w = (lane quadrant q = w & 3, K quarter / column slice w >> 2)
constexpr int NT = (NW + 4) * 32;             // + a service warpgroup: the MMA warp and three idle warps (setmaxnreg works on whole warpgroups)
constexpr int kRegsEpi = 104, kRegsSvc = 64;  // after the hand-over: 16 x 104 + 4 x 64 = 20 x 96, the launch allocation (the SM's spare registers are NOT available to setmaxnreg.inc: measured, it blocks for ever)
constexpr int kChunks = kRnn / 8;             // 64 chunks of 8 fp16 per activation row
constexpr size_t kMatChunks = (size_t)kChunks * 128;     // one buffer of one exchange matrix: [chunk][fold] x 16 bytes = 128 KB
enum { MH1 = 0, MH2, MS2, MF1, MF2, kMats };
constexpr uint32_t kTagE = 0x40004000u;       // generation bit of matrices with |v| < 2: bit 14 of every half
constexpr uint32_t kTagS = 0x80008000u;       // of non-negative matrices (ReLU outputs): the sign bit
// TMEM columns
constexpr uint32_t kColA = 0;                 // activation operand, 128 lanes x 256 columns (512 fp16 per lane)
constexpr uint32_t kColD0 = 256;              // on-path accumulator (T1: W_hh1 h1 [96] lives here too, see below)
constexpr uint32_t kColD1 = 352;              // second accumulator
// shared memory: weight tiles (K-major SWIZZLE_128B, [k-block 8][rows N][128 B]) then constants and the control block
constexpr int kW0 = 0;                                   // first tile: T1 W_hh1 (96 rows), T2 W_ih2a (96), T3 fc1a (64), T4 fc2 (64)
constexpr int kW1 = 96 * 128 * 8;                        // second tile: T1 fc3 (32 rows), T2 W_hh2 (96)
constexpr int kWEnd = kW1 + 96 * 128 * 8;                // 196608
constexpr int kConstOfs = kWEnd;                         // per-unit constants, <= 5 x 64 floats
constexpr int kCtlOfs = kConstOfs + 2048;
constexpr int kSmemBytes = kCtlOfs + 256;

struct Ctl {
    uint64_t abar[4];      // K quarter kq of the A operand is in TMEM (4 warps arrive)
    uint64_t dbar[2];      // accumulator complete (tcgen05.commit): [0] on-path job, [1] recurrent (off-path) job
    uint64_t ebar;         // all 16 epilogue warps have read the recurrent accumulator of the previous step
    uint32_t tmem;
    int abort_local;
};

__device__ __forceinline__ bool aborted_local(Ctl* c) { return *reinterpret_cast<volatile int*>(&c->abort_local) != 0; }
// warp-uniform view of the abort flag (lane 0's): the step loops end together for all lanes of a warp
__device__ __forceinline__ bool warp_aborted(Ctl* c) { return __shfl_sync(0xffffffffu, aborted_local(c) ? 1 : 0, 0) != 0; }
__device__ __noinline__ bool spin_check(const RsParams& p, Ctl* c, long long& t0) {
    if (aborted_local(c)) return true;
    if (ld_volatile_i32(p.abort_flag) != 0) { *reinterpret_cast<volatile int*>(&c->abort_local) = 1; return true; }
    if (t0 == 0) t0 = clock64();
    if (clock64() - t0 > g_rs_deadline) {
        *reinterpret_cast<volatile int*>(&c->abort_local) = 1;
        atomicExch(p.abort_flag, 1);
        return true;
    }
    return false;
}
// kSleep: the 16 ingest / epilogue warps wait for the accumulator while the MMA warp issues: every instruction they spend
// polling is an issue slot the MMA warp does not get (measured: +30 clocks per MMA), so they back off between tries
template <bool kSleep = false>
__device__ __forceinline__ bool wait_mbar(const RsParams& p, Ctl* c, uint64_t* bar, uint32_t parity) {
    long long t0 = 0;
    int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (kSleep) __nanosleep(100);
        if (((++spins) & 15) == 0 && aborted_local(c)) return false;
        if ((spins & 1023) == 0 && spin_check(p, c, t0)) return false;
    }
    return true;
}

// optional checkpoints into mapped host memory (WRNN_RS_DEBUG=1): [CTA][32 warps] last (step << 8 | code) of lane 0
__device__ __forceinline__ void dbg(const RsParams& p, int t, int code) {
    if (p.dbg && (threadIdx.x & 31) == 0) {
        *reinterpret_cast<volatile int*>(p.dbg + blockIdx.x * 32 + (threadIdx.x >> 5)) = (t << 8) | code;
    }
}

// optional timeline (WRNN_RS_TRACE=path): %globaltimer (ns, common to all SMs) of steps [kTraceStep0, +kTraceSteps) per CTA,
// written by lane 0 of warp 0 (events 0..8) and of the MMA warp (9..12): [CTA][step][16]
constexpr int kTraceStep0 = 96, kTraceSteps = 8;
__device__ __forceinline__ void trace(const RsParams& p, int t, int ev) {
    if (p.trace && (threadIdx.x == 0 || threadIdx.x == NW * 32) && t >= kTraceStep0 && t < kTraceStep0 + kTraceSteps) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        p.trace[((size_t)blockIdx.x * kTraceSteps + (t - kTraceStep0)) * 48 + ev] = ns;
    }
}

// SM-clock stamps of warp 0 / lane 0 inside an epilogue (slots 32 + k): the globaltimer ticks too coarsely (32-256 ns) for these
__device__ __forceinline__ void ctrace(const RsParams& p, int t, int k) {
    if (p.trace && threadIdx.x == 0 && t >= kTraceStep0 && t < kTraceStep0 + kTraceSteps)
        p.trace[((size_t)blockIdx.x * kTraceSteps + (t - kTraceStep0)) * 48 + 32 + k] = (unsigned long long)clock64();
}

// per-K-quarter events of the first ingest of a step: lane 0 of warps (q = 0, cs): slots 16 + 4 cs + {0 canaries, 1 loaded, 2 in TMEM}
__device__ __forceinline__ void trace_kq(const RsParams& p, int t, int ev0, int cs, int k) {
    if (p.trace && ev0 == 1 && (threadIdx.x & 127) == 0 && threadIdx.x < NW * 32 && t >= kTraceStep0 && t < kTraceStep0 + kTraceSteps) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        p.trace[((size_t)blockIdx.x * kTraceSteps + (t - kTraceStep0)) * 48 + 16 + 4 * cs + k] = ns;
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
__device__ __forceinline__ bool tags_ok(uint4 v, uint32_t tb, uint32_t want) {
    return ((((v.x ^ want) | (v.y ^ want) | (v.z ^ want) | (v.w ^ want)) & tb) == 0u);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint4& a, const uint4& b, const uint4& c, const uint4& d) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w),
        "r"(d.x), "r"(d.y), "r"(d.z), "r"(d.w)
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

template <bool kAcc>
__device__ __forceinline__ void umma_ts_c(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc) {
    if (kAcc)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
                     "r"(tmem_a), "l"(bdesc), "r"(idesc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
                     "r"(tmem_a), "l"(bdesc), "r"(idesc) : "memory");
}

// exp / reciprocal as ONE special-function instruction each (ex2.approx.ftz / rcp.approx.ftz): __expf and __fdividef wrap
// the same MUFU operations in three more instructions of denormal handling, and the GRU epilogues are instruction-bound
// (24 exp + 24 reciprocals per thread and step).  Results differ from __expf only below 1e-38, where a sigmoid is 0 or 1.
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_ftz(1.0f + ex2_ftz(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(-2.0f, rcp_ftz(1.0f + ex2_ftz(2.8853900817779268f * x)), 1.0f); }
// per-unit constants live in shared memory: read them with ld.shared (the generic pointer would cost a generic LD each)
__device__ __forceinline__ float4 lds4(const float* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ void lds8(const float* p, float* out) {
    const float4 a = lds4(p), b = lds4(p + 4);
    out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w; out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
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
    int q, cs, lane, row;         // lane quadrant, column slice / K quarter, lane, fold row inside the group
    bool live;
    uint32_t tlane;               // TMEM address of my lane quadrant, column 0
};

// Receive one activation matrix (this step's buffer) into the A operand in TMEM: warp (q, kq) takes the 16 chunks
// [16 kq, 16 kq + 16) of its 32 folds.  `extra` (optional): one more chunk of my row, returned to the caller (T2: my own
// units of h1).  Lanes 0..15 first poll one canary chunk each (chunk 16 kq + lane of row 32 q + lane), then every lane
// loads its 16 chunks once and re-polls the ones whose generation bits do not match yet.
// (inlined on purpose: as a real call the ABI spills around it cost more than the code size saves -- 25.8 vs 19.2 us per step)
struct IngestOut { uint4 extra; float x; };
__device__ __forceinline__ IngestOut ingest(const RsParams& p, Ctl* ctl, const Lane& L, const uint4* mat, uint32_t tb, uint32_t want,
                                         int extra_chunk, int dbg_t, int ev0, const unsigned long long* xw, uint32_t xtag) {
    uint4 extra;
    float xval = 0.f;
    const uint4* base = mat + (size_t)(L.cs * 16) * 128 + L.row;
    dbg(p, dbg_t, 0x10);
    {   // phase 1: one canary chunk per lane (chunk 16 kq + (lane & 15) of my own row) until the first producer of this K
        // quarter shows this step's generation (p.canary_all: until all do) -- one light load per lane and pass while the data
        // is still far away; the full passes of phase 2 then overlap the arrival of the remaining producers
        const uint4* cp = mat + (size_t)(L.cs * 16 + (L.lane & 15)) * 128 + L.row;
        long long t0 = 0;
        int spins = 0;
        const bool any_live = __any_sync(0xffffffffu, L.live);
        bool ok = !L.live;
        while (any_live) {
            if (!ok || !p.canary_all) ok = L.live ? tags_ok(ld_chunk(cp), tb, want) : p.canary_all != 0;
            if (p.canary_all ? __all_sync(0xffffffffu, ok) : __any_sync(0xffffffffu, ok)) break;
            const bool quit = ((++spins) & 255) == 0 && spin_check(p, ctl, t0);     // (spins is warp-uniform)
            if (__any_sync(0xffffffffu, quit)) break;
        }
    }
    dbg(p, dbg_t, 0x11);
    trace(p, dbg_t, ev0);
    trace_kq(p, dbg_t, ev0, L.cs, 0);
    if (ev0 == 1) ctrace(p, dbg_t, 8);
    unsigned long long xword = 0ull;
    if (xw) xword = ll_load(xw);            // the sample word was published before this matrix: its load rides along with phase 2
    // phase 2: full passes (16 loads at immediate offsets, one OR-reduction of the generation bits) until everything matches.
    // Lean on purpose: at 16 warps per SM every instruction of this path costs ~4 clocks of the step.
    uint4 v[16];
    const uint4* xp = mat + (size_t)(extra_chunk >= 0 ? extra_chunk : 0) * 128 + L.row;
    extra = make_uint4(0u, 0u, 0u, 0u);
    int passes = 0;
    {
        long long t0 = 0;
        int spins = 0;
        uint32_t bad;
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = make_uint4(0u, 0u, 0u, 0u);
        do {
            ++passes;
            if (L.live) {          // (padding rows are not fetched: the pass is bound by the SM's ~90 B/clk from L2)
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = ld_chunk(base + i * 128);
                if (extra_chunk >= 0) extra = ld_chunk(xp);
            }
            bad = (extra_chunk >= 0) ? ((extra.x ^ want) | (extra.y ^ want) | (extra.z ^ want) | (extra.w ^ want)) : 0u;
#pragma unroll
            for (int i = 0; i < 16; ++i) bad |= (v[i].x ^ want) | (v[i].y ^ want) | (v[i].z ^ want) | (v[i].w ^ want);
            bad = L.live ? (bad & tb) : 0u;
        } while (bad != 0u && !(((++spins) & 255) == 0 && spin_check(p, ctl, t0)));
    }
    __syncwarp();
    dbg(p, dbg_t, 0x12);
    trace(p, dbg_t, ev0 + 1);
    trace_kq(p, dbg_t, ev0, L.cs, 1);
    if (ev0 == 1) ctrace(p, dbg_t, 9);
    if (p.trace && threadIdx.x == 0 && dbg_t >= kTraceStep0 && dbg_t < kTraceStep0 + kTraceSteps)
        p.trace[((size_t)blockIdx.x * kTraceSteps + (dbg_t - kTraceStep0)) * 48 + (ev0 == 1 ? 13 : 14)] = (unsigned long long)passes;
    if (xw) {
        long long t0 = 0;
        int spins = 0;
        while (L.live && ll_tag(xword) != xtag) {
            xword = ll_load(xw);
            if (((++spins) & 255) == 0 && spin_check(p, ctl, t0)) break;
        }
        xval = L.live ? ll_val(xword) : 0.f;
    }
    if (ev0 == 1) ctrace(p, dbg_t, 10);
    if (want != 0u) {          // generation 1: the bit is set in every half; take it out (generation 0 needs nothing)
#pragma unroll
        for (int i = 0; i < 16; ++i) { v[i].x ^= tb; v[i].y ^= tb; v[i].z ^= tb; v[i].w ^= tb; }
        extra.x ^= tb; extra.y ^= tb; extra.z ^= tb; extra.w ^= tb;
    }
    if (__any_sync(0xffffffffu, L.live)) {       // (a quadrant of padding rows keeps whatever it holds: its accumulator rows are never read)
#pragma unroll
        for (int j = 0; j < 4; ++j) tmem_st16(L.tlane + kColA + (uint32_t)(L.cs * 64 + j * 16), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    tmem_st_wait();
    tcgen05_fence_before();
    __syncwarp();
    if (L.lane == 0) mbar_arrive(&ctl->abar[L.cs]);
    dbg(p, dbg_t, 0x13);
    trace(p, dbg_t, ev0 + 2);
    trace_kq(p, dbg_t, ev0, L.cs, 2);
    if (ev0 == 1) ctrace(p, dbg_t, 11);
    IngestOut o;
    o.extra = extra; o.x = xval;
    return o;
}

// Publish 8 values of my fold as one chunk (generation bit in every half).
__device__ __forceinline__ void publish8(uint4* mat, int chunk, int row, const float* v, uint32_t tb, uint32_t want) {
    uint4 w;
    w.x = (pack2(v[0], v[1]) & ~tb) | want; w.y = (pack2(v[2], v[3]) & ~tb) | want;
    w.z = (pack2(v[4], v[5]) & ~tb) | want; w.w = (pack2(v[6], v[7]) & ~tb) | want;
    st_chunk(mat + (size_t)chunk * 128 + row, w);
}

// the MOL draw of one fold from its 30 outputs (vocoder/distribution.py:104-140; same arithmetic as loop_tc.cu), in two
// halves: the noise depends on (step, fold) only and is drawn while the step's activations are still travelling
struct MolNoise { float gum[10]; float lnoise; };
__device__ __forceinline__ void mol_noise(MolNoise& nz, uint32_t t, uint32_t fold, uint32_t utt, uint2 key) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        const uint4 r = philox4x32_10(make_uint4(t, fold, utt, (uint32_t)b), key);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int i = 4 * b + w;
            if (i < 10) nz.gum[i] = -__logf(-__logf(1e-5f + u01(word_of(r, w)) * (1.0f - 2e-5f)));
        }
        if (b == 2) {
            const float ul = 1e-5f + u01(r.z) * (1.0f - 2e-5f);
            nz.lnoise = __logf(ul) - __logf(1.0f - ul);
        }
    }
}
__device__ __forceinline__ float mol_draw(const float* lg, const float* sbias, const MolNoise& nz) {
    float best = -INFINITY;
    int kbest = 0;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const float sc = lg[i] + sbias[i] + nz.gum[i];
        if (sc > best) { best = sc; kbest = i; }
    }
    float mean = 0.f, lsc = 0.f;
#pragma unroll
    for (int i = 0; i < 10; ++i)
        if (i == kbest) { mean = lg[10 + i] + sbias[10 + i]; lsc = lg[20 + i] + sbias[20 + i]; }
    lsc = fmaxf(lsc, -32.23619130191664f);
    const float xs = mean + __expf(lsc) * nz.lnoise;
    return fminf(fmaxf(xs, -1.0f), 1.0f);
}

// One MMA job: D[128 folds x N] = A (TMEM, 512 fp16 per lane) x W^T (shared memory tile [k-block][N][64]); issued K quarter by K
// quarter as the ingest warps deliver them.  Whole warp in the loop, one elected lane issues (tc_common.cuh: elect_one); a K
// step is one add on the descriptor and one UTCHMMA.  One rolled copy of the code for every job of every role.
__device__ __forceinline__ void mma_job(const RsParams& p, Ctl* ctl, uint32_t w_smem, uint32_t N, uint32_t d, uint32_t a0, uint32_t a_par,
                                     uint64_t* done, bool wait_e, uint32_t e_par, int tt, int ev) {
    const uint32_t idesc = umma_idesc_f16(128, (int)N);
    const uint64_t bd0 = umma_desc_sw128(w_smem);
    const uint32_t kb_step = N * 8u;                      // one k-block of the tile, in descriptor units of 16 bytes
    // The four K quarters reach TMEM within ~0.3 us of each other, and one wait + fence + elect round costs as much as
    // eight MMAs: wait for all four, then issue the 32 K steps in one go.
#pragma unroll
    for (int kq = 0; kq < 4; ++kq) wait_mbar(p, ctl, &ctl->abar[kq], a_par);
    if (wait_e) wait_mbar(p, ctl, &ctl->ebar, e_par);
    tcgen05_fence_after();
    trace(p, tt, ev);
    if (elect_one()) {
        uint64_t bd = bd0;
        uint32_t a = a0;
#pragma unroll 1
        for (int kb = 0; kb < 8; ++kb) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                umma_ts(d, a, bd + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                a += 8u;
            }
            bd += kb_step;
        }
        umma_commit(done);
    }
    __syncwarp();
    trace(p, tt, ev + 1);
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

}  // namespace

__global__ void __launch_bounds__(NT, 1) wrnn_loop_rs_kernel(const __grid_constant__ RsParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    Ctl* ctl = reinterpret_cast<Ctl*>(smem + kCtlOfs);
    float* cst = reinterpret_cast<float*>(smem + kConstOfs);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = (int)blockIdx.x / kRsCtas, rc = (int)blockIdx.x % kRsCtas;
    const int role = rc < kRsT1 ? 0 : (rc < kRsT1 + kRsT2 ? 1 : (rc < kRsT1 + kRsT2 + kRsT3 ? 2 : 3));
    const int cta = role == 0 ? rc : (role == 1 ? rc - kRsT1 : (role == 2 ? rc - kRsT1 - kRsT2 : rc - kRsT1 - kRsT2 - kRsT3));
    const int fold0 = g * p.Ng, nrows = max(0, min(p.Ng, p.B - fold0));
    const int S = p.S;
    const bool expander = (int)blockIdx.x >= p.G * kRsCtas;        // CTAs past the groups produce the conditioning records
    const unsigned int consumers = (unsigned int)(p.G * kRsCtas * NW);   // warps that read every record chunk

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
            const int e_idx = (int)blockIdx.x - p.G * kRsCtas;
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
                            if (((++spins) & 63) == 0 && spin_check(p, ctl, t0)) break;
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
    if ((int)blockIdx.x < p.G * kRsCtas) {
        const unsigned char* img = role == 0 ? p.w1 + (size_t)cta * (kW1 + 32 * 128 * 8)
                                 : role == 1 ? p.w2 + (size_t)cta * kWEnd
                                 : role == 2 ? p.w3 + (size_t)cta * (64 * 128 * 8) : p.w4 + (size_t)cta * (64 * 128 * 8);
        const int bytes = role == 0 ? kW1 + 32 * 128 * 8 : (role == 1 ? kWEnd : 64 * 128 * 8);
        const uint4* src = reinterpret_cast<const uint4*>(img);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        for (int i = tid; i < bytes / 16; i += NT) dst[i] = src[i];
        fence_proxy_async_smem();
    }
    if (role == 0) {            // [v1 r,z,n | b_hn1] x 32 units, fc3 bias (32)
        if (tid < 128) { const int a = tid >> 5, u = tid & 31, j = 32 * cta + u; cst[tid] = a < 3 ? p.v1[a * kRnn + j] : p.bhn1[j]; }
        else if (tid < 160) cst[tid] = (tid - 128) < 30 ? p.bfc3[tid - 128] : 0.f;
    } else if (role == 1) {     // [v2 r,z,n | b_hn2] x 32 units
        if (tid < 128) { const int a = tid >> 5, u = tid & 31, j = 32 * cta + u; cst[tid] = a < 3 ? p.v2[a * kRnn + j] : p.bhn2[j]; }
    } else if (role == 2) {     // v3 x 64 units
        if (tid < 64) cst[tid] = p.v3[64 * cta + tid];
    }
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&ctl->abar[i], 4);
        mbar_init(&ctl->dbar[0], 1); mbar_init(&ctl->dbar[1], 1);
        mbar_init(&ctl->ebar, NW);
        ctl->abort_local = 0;
        mbar_fence_init();
    }
    if (warp == 0 && (int)blockIdx.x < p.G * kRsCtas) tmem_alloc(&ctl->tmem, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = ctl->tmem;
    dbg(p, 0, 1);
    // (register hand-over, first statement of each role's branch below: the service warpgroup gives its registers to the
    //  16 ingest / epilogue warps, so nothing on the chain spills)

    uint4* const X = p.X + (size_t)g * kMats * kRsBufs * kMatChunks;
#define MAT(m, t) (X + ((size_t)(m) * kRsBufs + ((t) % kRsBufs)) * kMatChunks)
#define GEN(t) ((((t) / kRsBufs) & 1) != 0)

    if (warp >= NW) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsSvc));     // (one instruction for the whole warpgroup)
        dbg(p, 0, 2);
        if (warp == NW) {
        // =================================== MMA issuer ====================================================================
        // whole warp in the loop, one elected lane issues (tc_common.cuh: elect_one).  A job = 32 K-steps over the A operand
        // in TMEM, issued K quarter by K quarter as the ingest warps deliver them.
        uint32_t n_ingest = 0;
        auto job = [&](uint32_t N, uint32_t dcol, uint32_t wofs, uint64_t* done, bool wait_e, uint32_t e_par, int tt, int ev) {
            mma_job(p, ctl, smem_u32(smem + wofs), N, tmem + dcol, tmem + kColA, n_ingest & 1u, done, wait_e, e_par, tt, ev);
            ++n_ingest;
        };
        if (role == 0) {
            for (int t = 0; t <= S && !warp_aborted(ctl); ++t) {
                if (t > 0) job(32u, kColD1, kW1, &ctl->dbar[0], false, 0, t, 9);                       // fc3 f2(t-1)
                if (t < S) job(96u, kColD0, kW0, &ctl->dbar[1], true, (uint32_t)t & 1u, t, 11);         // W_hh1 h1(t)
            }
        } else if (role == 1) {
            for (int t = 0; t < S && !warp_aborted(ctl); ++t) {
                job(96u, kColD0, kW0, &ctl->dbar[0], false, 0, t, 9);                             // W_ih2a h1(t)
                job(96u, kColD1, kW1, &ctl->dbar[1], true, (uint32_t)t & 1u, t, 11);                    // W_hh2 h2(t)
            }
        } else {
            for (int t = 0; t < S && !warp_aborted(ctl); ++t) job(64u, kColD0, kW0, &ctl->dbar[0], false, 0, t, 9);          // fc1a s2(t) / fc2 f1(t)
        }
        }
    } else {
        // =================================== ingest + epilogue warps ========================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpi));
        dbg(p, 0, 3);
        Lane L;
        L.q = warp & 3; L.cs = warp >> 2; L.lane = lane; L.row = 32 * L.q + lane; L.live = L.row < nrows;
        L.tlane = tmem + ((uint32_t)(32 * L.q) << 16);
        const bool warp_live = 32 * L.q < nrows;       // a warp whose 32 rows are all padding skips the arithmetic (frees issue slots)
        const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
        const FoldDesc fd = p.folds[L.live ? fold0 + L.row : 0];
        const size_t srow = (size_t)(fold0 + L.row) * S;                     // my fold's row of samples / forced
        unsigned long long* const xw = p.bX + (size_t)g * 128 + L.row;
        const float* const csrow = p.CS + ((size_t)g * p.cs_steps * p.Ng + L.row) * 4096;   // + (t % cs_steps) * Ng * 4096
        const size_t cs_step = (size_t)p.Ng * 4096;
        // the expanders run a few chunks ahead: before the first step of a chunk, lane 0 acquires its counter
        auto cs_wait = [&](int t) {
            if (p.cs_done && (t % kRsChunk) == 0) {
                if (lane == 0) {
                    long long t0 = 0;
                    int spins = 0;
                    while (true) {
                        unsigned int v;
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.cs_done + t / kRsChunk) : "memory");
                        if (v >= (unsigned int)p.B) break;
                        if (((++spins) & 63) == 0 && spin_check(p, ctl, t0)) break;
                    }
                }
                __syncwarp();
            }
        };
        // ... and after the last step of a chunk (its records are in registers and used) the warp hands the ring slot back
        auto cs_release = [&](int t) {
            if (p.cs_done && ((t % kRsChunk) == kRsChunk - 1 || t == S - 1)) {
                __syncwarp();
                if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.cs_consumed + t / kRsChunk) : "memory");
            }
        };
        if (role == 0) {
            // ---- T1: fc3 + draw of step t-1, GRU1 of step t, then the recurrent product for step t+1 -----------------------
            float h1[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) h1[i] = 0.f;
            const float* v1 = cst + L.cs * 8;                    // [a][32]: + 32 a
            const float* sbias = cst + 128;
            for (int t = 0; t <= S && !warp_aborted(ctl); ++t) {
                trace(p, t, 0);
                float c1[24];
                if (t < S) cs_wait(t);
                if (t < S && L.live) {
                    const float* c = csrow + (size_t)(t % p.cs_steps) * cs_step + 32 * cta + 8 * L.cs;
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const float4 lo = __ldcg(reinterpret_cast<const float4*>(c + a * 512)), hi = __ldcg(reinterpret_cast<const float4*>(c + a * 512) + 1);
                        c1[8 * a + 0] = lo.x; c1[8 * a + 1] = lo.y; c1[8 * a + 2] = lo.z; c1[8 * a + 3] = lo.w;
                        c1[8 * a + 4] = hi.x; c1[8 * a + 5] = hi.y; c1[8 * a + 6] = hi.z; c1[8 * a + 7] = hi.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 24; ++i) c1[i] = 0.f;
                }
                float x = 0.f;
                if (t > 0) {
                    MolNoise nz;
                    if (warp_live) mol_noise(nz, (uint32_t)(t - 1), (uint32_t)fd.fold, (uint32_t)fd.utt, key);
                    // the A buffer is free once the recurrent job of step t-1 has completed
                    wait_mbar(p, ctl, &ctl->dbar[1], (uint32_t)(t - 1) & 1u);
                    tcgen05_fence_after();
                    ingest(p, ctl, L, MAT(MF2, t - 1), kTagS, GEN(t - 1) ? kTagS : 0u, -1, t, 1, nullptr, 0u);
                    ctrace(p, t, 0);
                    wait_mbar<true>(p, ctl, &ctl->dbar[0], (uint32_t)(t - 1) & 1u);
                    tcgen05_fence_after();
                    trace(p, t, 4);
                    ctrace(p, t, 1);
                    float lg[32];
                    tmem_ld8(L.tlane + kColD1 + 0, lg); tmem_ld8(L.tlane + kColD1 + 8, lg + 8);
                    tmem_ld8(L.tlane + kColD1 + 16, lg + 16); tmem_ld8(L.tlane + kColD1 + 24, lg + 24);
                    tmem_ld_wait();
                    ctrace(p, t, 2);
                    const float xs = warp_live ? mol_draw(lg, sbias, nz) : 0.f;
                    x = xs;
                    if (L.live) {
                        if (p.forced) x = p.forced[srow + t - 1];
                        if (cta == 0 && L.cs == 0) {
                            p.samples[srow + t - 1] = xs;
                            ll_store(xw, x, (uint32_t)t);
                            if (p.logits_out)
                                for (int i = 0; i < 30; ++i) p.logits_out[(srow + t - 1) * 30 + i] = lg[i] + sbias[i];
                        }
                    }
                }
                if (t == S) break;
                ctrace(p, t, 3);
                float gh[24];
                if (t > 0) {
                    tmem_ld8(L.tlane + kColD0 + 0 + 8 * L.cs, gh); tmem_ld8(L.tlane + kColD0 + 32 + 8 * L.cs, gh + 8);
                    tmem_ld8(L.tlane + kColD0 + 64 + 8 * L.cs, gh + 16);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 24; ++i) gh[i] = 0.f;
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->ebar);
                ctrace(p, t, 4);
                if (warp_live) {
                float kr[8], kz[8], kn[8], kb[8];
                lds8(v1, kr); lds8(v1 + 32, kz); lds8(v1 + 64, kn); lds8(v1 + 96, kb);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float r = sigmoid_fast(fmaf(kr[i], x, c1[i]) + gh[i]);
                    const float z = sigmoid_fast(fmaf(kz[i], x, c1[8 + i]) + gh[8 + i]);
                    const float n = tanh_fast(fmaf(kn[i], x, c1[16 + i]) + r * (gh[16 + i] + kb[i]));
                    h1[i] = fmaf(z, h1[i] - n, n);
                }
                ctrace(p, t, 5);
                if (L.live) publish8(MAT(MH1, t), 4 * cta + L.cs, L.row, h1, kTagE, GEN(t) ? kTagE : 0u);
                }
                trace(p, t, 5);
                ctrace(p, t, 6);
                // recurrent product for step t+1: the full h1(t) -> A buffer (the fc3 job has completed: dbar[0] was waited).
                // Off the critical path: wait until the T2 CTAs have read the same lines for the on-path product.
                __nanosleep(p.offpath_delay_ns);
                ingest(p, ctl, L, MAT(MH1, t), kTagE, GEN(t) ? kTagE : 0u, -1, t, 6, nullptr, 0u);
                cs_release(t);
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
            const float* v2 = cst + L.cs * 8;
            for (int t = 0; t < S && !warp_aborted(ctl); ++t) {
                trace(p, t, 0);
                float c2[24];
                cs_wait(t);
                if (L.live) {
                    const float* c = csrow + (size_t)(t % p.cs_steps) * cs_step + 3 * 512 + 32 * cta + 8 * L.cs;
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const float4 lo = __ldcg(reinterpret_cast<const float4*>(c + a * 512)), hi = __ldcg(reinterpret_cast<const float4*>(c + a * 512) + 1);
                        c2[8 * a + 0] = lo.x; c2[8 * a + 1] = lo.y; c2[8 * a + 2] = lo.z; c2[8 * a + 3] = lo.w;
                        c2[8 * a + 4] = hi.x; c2[8 * a + 5] = hi.y; c2[8 * a + 6] = hi.z; c2[8 * a + 7] = hi.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 24; ++i) c2[i] = 0.f;
                }
                if (t > 0) {
                    wait_mbar(p, ctl, &ctl->dbar[1], (uint32_t)(t - 1) & 1u);
                    tcgen05_fence_after();
                }
                const IngestOut io = ingest(p, ctl, L, MAT(MH1, t), kTagE, GEN(t) ? kTagE : 0u, 4 * cta + L.cs, t, 1, t > 0 ? xw : nullptr, (uint32_t)t);
                const float x = io.x;
                const uint4 extra = io.extra;
                dbg(p, t, 0x30);
                ctrace(p, t, 0);
                wait_mbar<true>(p, ctl, &ctl->dbar[0], (uint32_t)t & 1u);
                tcgen05_fence_after();
                dbg(p, t, 0x31);
                trace(p, t, 4);
                ctrace(p, t, 1);
                float pb[24], gh[24];
                tmem_ld8(L.tlane + kColD0 + 0 + 8 * L.cs, pb); tmem_ld8(L.tlane + kColD0 + 32 + 8 * L.cs, pb + 8);
                tmem_ld8(L.tlane + kColD0 + 64 + 8 * L.cs, pb + 16);
                if (t > 0) {
                    tmem_ld8(L.tlane + kColD1 + 0 + 8 * L.cs, gh); tmem_ld8(L.tlane + kColD1 + 32 + 8 * L.cs, gh + 8);
                    tmem_ld8(L.tlane + kColD1 + 64 + 8 * L.cs, gh + 16);
                } else {
#pragma unroll
                    for (int i = 0; i < 24; ++i) gh[i] = 0.f;
                }
                tmem_ld_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->ebar);
                ctrace(p, t, 4);
                float s2[8];
                const float2 e0 = unpack2(extra.x), e1 = unpack2(extra.y), e2 = unpack2(extra.z), e3 = unpack2(extra.w);
                const float h1o[8] = {e0.x, e0.y, e1.x, e1.y, e2.x, e2.y, e3.x, e3.y};
                if (warp_live) {
                float kr[8], kz[8], kn[8], kb[8];
                lds8(v2, kr); lds8(v2 + 32, kz); lds8(v2 + 64, kn); lds8(v2 + 96, kb);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float r = sigmoid_fast(pb[i] + fmaf(kr[i], x, c2[i]) + gh[i]);
                    const float z = sigmoid_fast(pb[8 + i] + fmaf(kz[i], x, c2[8 + i]) + gh[8 + i]);
                    const float n = tanh_fast(pb[16 + i] + fmaf(kn[i], x, c2[16 + i]) + r * (gh[16 + i] + kb[i]));
                    h2[i] = fmaf(z, h2[i] - n, n);
                    s2[i] = fminf(fmaxf(h1o[i] + h2[i], -1.9990234375f), 1.9990234375f);
                }
                ctrace(p, t, 5);
                if (L.live) {
                    publish8(MAT(MS2, t), 4 * cta + L.cs, L.row, s2, kTagE, GEN(t) ? kTagE : 0u);
                    publish8(MAT(MH2, t), 4 * cta + L.cs, L.row, h2, kTagE, GEN(t) ? kTagE : 0u);
                }
                }
                trace(p, t, 5);
                ctrace(p, t, 6);
                ingest(p, ctl, L, MAT(MH2, t), kTagE, GEN(t) ? kTagE : 0u, -1, t, 6, nullptr, 0u);
                cs_release(t);
            }
        } else {
            // ---- T3 / T4: fc1 on s2(t) (+ the sample's rank-1 term) / fc2 on f1(t); ReLU; publish ------------------------------
            const bool fc1 = role == 2;
            const float* v3 = cst + L.cs * 16;
            for (int t = 0; t < S && !warp_aborted(ctl); ++t) {
                trace(p, t, 0);
                float cc[16];
                cs_wait(t);
                if (L.live) {
                    const float4* c = reinterpret_cast<const float4*>(csrow + (size_t)(t % p.cs_steps) * cs_step + (fc1 ? 6 : 7) * 512 + 64 * cta + 16 * L.cs);
#pragma unroll
                    for (int i = 0; i < 4; ++i) { const float4 q4 = __ldcg(c + i); cc[4 * i] = q4.x; cc[4 * i + 1] = q4.y; cc[4 * i + 2] = q4.z; cc[4 * i + 3] = q4.w; }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) cc[i] = 0.f;
                }
                const float x = ingest(p, ctl, L, MAT(fc1 ? MS2 : MF1, t), fc1 ? kTagE : kTagS, GEN(t) ? (fc1 ? kTagE : kTagS) : 0u, -1, t, 1,
                                       (fc1 && t > 0) ? xw : nullptr, (uint32_t)t).x;
                wait_mbar<true>(p, ctl, &ctl->dbar[0], (uint32_t)t & 1u);
                tcgen05_fence_after();
                trace(p, t, 4);
                float d[16];
                tmem_ld8(L.tlane + kColD0 + 16 * L.cs, d); tmem_ld8(L.tlane + kColD0 + 16 * L.cs + 8, d + 8);
                tmem_ld_wait();
                tcgen05_fence_before();
#pragma unroll
                for (int i = 0; i < 16; ++i) d[i] = fmaxf(d[i] + (fc1 ? fmaf(v3[i], x, cc[i]) : cc[i]), 0.f);
                if (L.live) {
                    uint4* m = MAT(fc1 ? MF1 : MF2, t);
                    const uint32_t want = GEN(t) ? kTagS : 0u;
                    publish8(m, 8 * cta + 2 * L.cs, L.row, d, kTagS, want);
                    publish8(m, 8 * cta + 2 * L.cs + 1, L.row, d + 8, kTagS, want);
                }
                trace(p, t, 5);
                cs_release(t);
            }
        }
    }
#undef MAT
#undef GEN
    // ---- teardown ---------------------------------------------------------------------------------------------------
    dbg(p, 0, 0xFF);
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

cudaError_t set_rs_deadline(long long cycles) { return cudaMemcpyToSymbol(g_rs_deadline, &cycles, sizeof(cycles)); }
size_t loop_rs_image_bytes(int role) { return role == 0 ? (size_t)kW1 + 32 * 128 * 8 : (role == 1 ? (size_t)kWEnd : (size_t)64 * 128 * 8); }
size_t loop_rs_exchange_bytes(int groups) { return (size_t)groups * kMats * kRsBufs * kMatChunks * 16; }

cudaError_t launch_expand_cond_rs(const float4* TA1, const float4* TA2, const float4* TQ1, const float4* TQ2, const float* coef,
                                  const FoldDesc* folds, int B, int S, int Ng, int cs_steps, float* CS, cudaStream_t stream) {
    const int spb = 64;
    dim3 grid(B, (S + spb - 1) / spb);
    expand_cond_rs_kernel<<<grid, 512, 0, stream>>>(TA1, TA2, TQ1, TQ2, coef, folds, S, Ng, cs_steps, spb, CS);
    return cudaGetLastError();
}

cudaError_t launch_loop_rs(const RsParams& p, cudaStream_t stream) {
    cudaError_t err = cudaFuncSetAttribute(wrnn_loop_rs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes + 1024);
    if (err != cudaSuccess) return err;
    RsParams pp = p;
    void* args[] = {&pp};
    const int grid = p.G * kRsCtas + (p.cs_done ? p.n_expanders : 0);
    return cudaLaunchCooperativeKernel((const void*)wrnn_loop_rs_kernel, dim3(grid), dim3(NT), args, kSmemBytes + 1024, stream);
}

}  // namespace wrnn
