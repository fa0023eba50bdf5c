// cond_expand.cuh -- per-sample conditioning records for the tensor-core loop (shared by the stand-alone expansion
// kernel in cond.cu and by the expander CTAs inside the loop kernel, loop_tc.cu).
//
// Interpolates the per-frame tables into CS[virtual group][t][row][unit pair][16 floats], so that a loop thread reads one
// contiguous 64-byte record per step:
//   {c1_r[2], c1_z[2], c1_n[2], c2_r[2], c2_z[2], c2_n[2], c3[2], c4[2]}   (two hidden units per record)
// Positions past the utterance (fold tail padding, Q9) take the bias-only row and no mel share.
#pragma once
#include "common.cuh"

namespace wrnn {

// one work item = (fold b, steps [t0, t1)); executed by 256 threads, `tx` = 0..255 = unit pair
__device__ __forceinline__ void expand_cond_item(const float4* __restrict__ TA1, const float4* __restrict__ TA2,
                                                 const float4* __restrict__ TQ1, const float4* __restrict__ TQ2,
                                                 const float* __restrict__ coef, const FoldDesc& fd, int b, int t0, int t1, int cs_steps, int Mg,
                                                 float4* __restrict__ CS, int tx) {
    const int g = b / Mg, row = b - g * Mg;          // virtual group, row
    const int j = tx * 2;
    for (int t = t0; t < t1; ++t) {
        const int n = fd.n0 + t;
        const bool valid = n < fd.N;
        const int q0 = valid ? n / kHop : 0;
        const size_t ra = (size_t)(fd.ta_row0 + (valid ? q0 : fd.T)) * kRnn + j;
        float4 a1[2] = {__ldg(TA1 + ra), __ldg(TA1 + ra + 1)}, a2[2] = {__ldg(TA2 + ra), __ldg(TA2 + ra + 1)};
        if (valid) {
            const float* cf = coef + (n - q0 * kHop) * kTaps;
#pragma unroll
            for (int d = 0; d < kTaps; ++d) {
                const float c = __ldg(cf + d);
                if (c != 0.f) {
                    const size_t rq = (size_t)(fd.tq_row0 + q0 + d) * kRnn + j;
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const float4 q1 = __ldg(TQ1 + rq + u), q2 = __ldg(TQ2 + rq + u);
                        a1[u].x = fmaf(c, q1.x, a1[u].x); a1[u].y = fmaf(c, q1.y, a1[u].y);
                        a1[u].z = fmaf(c, q1.z, a1[u].z); a1[u].w = fmaf(c, q1.w, a1[u].w);
                        a2[u].x = fmaf(c, q2.x, a2[u].x); a2[u].y = fmaf(c, q2.y, a2[u].y); a2[u].z = fmaf(c, q2.z, a2[u].z);
                    }
                }
            }
        }
        float4* out = CS + ((((size_t)g * cs_steps + t) * Mg + row) * 256 + tx) * 4;
        __stcs(out + 0, make_float4(a1[0].x, a1[1].x, a1[0].y, a1[1].y));
        __stcs(out + 1, make_float4(a1[0].z, a1[1].z, a2[0].x, a2[1].x));
        __stcs(out + 2, make_float4(a2[0].y, a2[1].y, a2[0].z, a2[1].z));
        __stcs(out + 3, make_float4(a1[0].w, a1[1].w, a2[0].w, a2[1].w));
    }
}


// The same work item with the table rows of the current frame kept in shared memory: a frame lasts 200 steps, so the
// 24 float4 a thread needs per step are fetched from L2 once per frame instead of once per step (the expander CTAs of the
// loop kernel have almost no L1: the loop's shared-memory carve-out takes it).  `cache` = this thread group's private
// area of kExpandCacheFloats x 256 floats, laid out [slot][thread]: no barriers, no bank conflicts.
constexpr int kExpandCacheFloats = 16 + kTaps * 14;
__device__ __forceinline__ void expand_cond_item_cached(const float4* __restrict__ TA1, const float4* __restrict__ TA2,
                                                        const float4* __restrict__ TQ1, const float4* __restrict__ TQ2,
                                                        const float* __restrict__ coef, const FoldDesc& fd, int b, int t0, int t1, int cs_steps, int Mg,
                                                        float4* __restrict__ CS, int tx, float* __restrict__ cache, int& key) {
    const int g = b / Mg, row = b - g * Mg;
    const int j = tx * 2;
    float* c = cache + tx;
    for (int t = t0; t < t1; ++t) {
        const int n = fd.n0 + t;
        const bool valid = n < fd.N;
        const int q0 = valid ? n / kHop : 0;
        const int want = valid ? (fd.tq_row0 + q0) : -2 - fd.ta_row0;       // which table rows the cache must hold
        if (want != key) {
            key = want;
            const size_t ra = (size_t)(fd.ta_row0 + (valid ? q0 : fd.T)) * kRnn + j;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float4 x1 = __ldg(TA1 + ra + u), x2 = __ldg(TA2 + ra + u);
                c[(u * 8 + 0) * 256] = x1.x; c[(u * 8 + 1) * 256] = x1.y; c[(u * 8 + 2) * 256] = x1.z; c[(u * 8 + 3) * 256] = x1.w;
                c[(u * 8 + 4) * 256] = x2.x; c[(u * 8 + 5) * 256] = x2.y; c[(u * 8 + 6) * 256] = x2.z; c[(u * 8 + 7) * 256] = x2.w;
            }
            if (valid) {
#pragma unroll
                for (int d = 0; d < kTaps; ++d) {
                    const size_t rq = (size_t)(fd.tq_row0 + q0 + d) * kRnn + j;
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const float4 q1 = __ldg(TQ1 + rq + u), q2 = __ldg(TQ2 + rq + u);
                        float* q = c + (16 + d * 14 + u * 7) * 256;
                        q[0] = q1.x; q[256] = q1.y; q[512] = q1.z; q[768] = q1.w; q[1024] = q2.x; q[1280] = q2.y; q[1536] = q2.z;
                    }
                }
            }
        }
        float a[2][8];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[u][i] = c[(u * 8 + i) * 256];
        if (valid) {
            const float* cf = coef + (n - q0 * kHop) * kTaps;
#pragma unroll
            for (int d = 0; d < kTaps; ++d) {
                const float w = __ldg(cf + d);
                if (w != 0.f) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const float* q = c + (16 + d * 14 + u * 7) * 256;
#pragma unroll
                        for (int i = 0; i < 7; ++i) a[u][i] = fmaf(w, q[i * 256], a[u][i]);
                    }
                }
            }
        }
        float4* out = CS + ((((size_t)g * cs_steps + (t % cs_steps)) * Mg + row) * 256 + tx) * 4;   // CS is a ring of cs_steps steps
        __stcs(out + 0, make_float4(a[0][0], a[1][0], a[0][1], a[1][1]));
        __stcs(out + 1, make_float4(a[0][2], a[1][2], a[0][4], a[1][4]));
        __stcs(out + 2, make_float4(a[0][5], a[1][5], a[0][6], a[1][6]));
        __stcs(out + 3, make_float4(a[0][3], a[1][3], a[0][7], a[1][7]));
    }
}


// The same work item by 512 threads, thread = ONE hidden unit `j`: the 8 + 5 x 7 table values of the current frame stay in
// registers (no shared-memory traffic in the step loop: ~2x fewer instructions per record than the cached variant), the
// interpolation weights come from a copy of `coef` in shared memory (a broadcast read), and the two threads of a unit pair
// swap four values by shuffle so that each stores 32 contiguous bytes of the pair's 64-byte record.  Same FMA order as
// the variants above: bit-identical records.
__device__ __forceinline__ void expand_cond_item_regs(const float4* __restrict__ TA1, const float4* __restrict__ TA2,
                                                      const float4* __restrict__ TQ1, const float4* __restrict__ TQ2,
                                                      const float* __restrict__ coef_s, const FoldDesc& fd, int b, int t0, int t1, int cs_steps, int Mg,
                                                      float4* __restrict__ CS, int j) {
    const int g = b / Mg, row = b - g * Mg;
    const bool odd = (j & 1) != 0;
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
            const float* cf = coef_s + (n - q0 * kHop) * kTaps;
#pragma unroll
            for (int d = 0; d < kTaps; ++d) {
                const float w = cf[d];
                if (w != 0.f) {
#pragma unroll
                    for (int i = 0; i < 7; ++i) a[i] = fmaf(w, tq[d][i], a[i]);
                }
            }
        }
        // record of the pair (a0 = even unit, a1 = odd unit):
        //   {a0[0],a1[0],a0[1],a1[1]} {a0[2],a1[2],a0[4],a1[4]} | {a0[5],a1[5],a0[6],a1[6]} {a0[3],a1[3],a0[7],a1[7]}
        // the even thread stores the first half, the odd thread the second
        const float r0 = __shfl_xor_sync(0xffffffffu, odd ? a[0] : a[5], 1);
        const float r1 = __shfl_xor_sync(0xffffffffu, odd ? a[1] : a[6], 1);
        const float r2 = __shfl_xor_sync(0xffffffffu, odd ? a[2] : a[3], 1);
        const float r3 = __shfl_xor_sync(0xffffffffu, odd ? a[4] : a[7], 1);
        float4* out = CS + ((((size_t)g * cs_steps + (t % cs_steps)) * Mg + row) * 256 + (j >> 1)) * 4 + (odd ? 2 : 0);
        if (!odd) {
            __stcs(out + 0, make_float4(a[0], r0, a[1], r1));
            __stcs(out + 1, make_float4(a[2], r2, a[4], r3));
        } else {
            __stcs(out + 0, make_float4(r0, a[5], r1, a[6]));
            __stcs(out + 1, make_float4(r2, a[3], r3, a[7]));
        }
    }
}

}  // namespace wrnn
