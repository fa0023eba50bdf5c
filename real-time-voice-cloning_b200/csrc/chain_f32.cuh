// chain_f32.cuh -- building blocks of the fp32 weight-stationary loops of the small topologies (loop_rr.cu: runtimeracer,
// loop_gn.cu: geneing): the gather of an exchanged activation vector ({value, step-tag} words through L2, common.cuh) and the
// GEMV of a CTA's resident weight rows against the gathered vectors.  K = length of the vector (256 or 128).
#pragma once
#include "common.cuh"

namespace wrnn {
namespace chain {

constexpr int NT = 512;         // threads per CTA of these loops
constexpr int NW = NT / 32;

// Spin until every {value, tag} word of `nb` rows of `buf` carries `tag`; values into act[nb][H] and, per `mode`, into the running
// sum (0: none, 1: sum = act, 2: sum += act).  Returns nonzero (CTA-uniform) if the deadline passed or another CTA aborted.
template <int K>
__device__ __noinline__ int gather(const unsigned long long* __restrict__ buf, int nb, float* __restrict__ act, float* __restrict__ sum,
                                   int mode, uint32_t tag, int* abort_flag, long long deadline) {
    constexpr int MAXI = 8;                  // word pairs in flight per thread (one L2 round trip serves all of them)
    const int tid = threadIdx.x;
    const int npairs = nb * (K / 2);
    int failed = 0;
    if (npairs <= NT) {                      // at most one pair per thread (<= 4 folds): the plain spin is the shortest path
        if (tid < npairs) {
            unsigned long long a, b;
            long long t0 = 0;
            int spins = 0;
            while (true) {
                ll_load2(buf + 2 * (size_t)tid, a, b);
                if (ll_tag(a) == tag && ll_tag(b) == tag) break;
                if (((++spins) & 63) == 0) {
                    if (t0 == 0) t0 = clock64();
                    if (clock64() - t0 > deadline || ld_volatile_i32(abort_flag) != 0) { failed = 1; break; }
                }
            }
            const float2 v = make_float2(ll_val(a), ll_val(b));
            *reinterpret_cast<float2*>(act + 2 * (size_t)tid) = v;
            if (mode == 1) *reinterpret_cast<float2*>(sum + 2 * (size_t)tid) = v;
            else if (mode == 2) {
                float2 s2 = *reinterpret_cast<float2*>(sum + 2 * (size_t)tid);
                s2.x += v.x; s2.y += v.y;
                *reinterpret_cast<float2*>(sum + 2 * (size_t)tid) = s2;
            }
        }
        return __syncthreads_or(failed);
    }
    for (int base = 0; base < npairs && !failed; base += NT * MAXI) {
        uint32_t pending = 0;
#pragma unroll
        for (int i = 0; i < MAXI; ++i)
            if (base + tid + i * NT < npairs) pending |= 1u << i;
        long long t0 = 0;
        int spins = 0;
        while (pending) {
            unsigned long long a[MAXI], b[MAXI];
#pragma unroll
            for (int i = 0; i < MAXI; ++i)
                if ((pending >> i) & 1u) ll_load2(buf + 2 * (size_t)(base + tid + i * NT), a[i], b[i]);
#pragma unroll
            for (int i = 0; i < MAXI; ++i)
                if (((pending >> i) & 1u) && ll_tag(a[i]) == tag && ll_tag(b[i]) == tag) {
                    const size_t o = 2 * (size_t)(base + tid + i * NT);
                    const float2 v = make_float2(ll_val(a[i]), ll_val(b[i]));
                    *reinterpret_cast<float2*>(act + o) = v;
                    if (mode == 1) *reinterpret_cast<float2*>(sum + o) = v;
                    else if (mode == 2) {
                        float2 s2 = *reinterpret_cast<float2*>(sum + o);
                        s2.x += v.x; s2.y += v.y;
                        *reinterpret_cast<float2*>(sum + o) = s2;
                    }
                    pending &= ~(1u << i);
                }
            if (pending && ((++spins) & 63) == 0) {
                if (t0 == 0) t0 = clock64();
                if (clock64() - t0 > deadline || ld_volatile_i32(abort_flag) != 0) { failed = 1; break; }
            }
        }
    }
    return __syncthreads_or(failed);
}

// out[f][o0 + r] = sum_k W[r][k] * in[f][k] for r < R, f < nb.  One or two folds: one warp per (row, fold pair), lanes split K.
// More: warp tasks of 2 rows x 4 folds; the 8 partial sums are reduced with a halving butterfly (9 shuffles instead of 40).
template <int K>
__device__ __forceinline__ void dots(const float* __restrict__ sW, int R, const float* __restrict__ in, int nb, float* __restrict__ out, int ldo,
                                     int o0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (nb <= 2) {
        const int f1 = nb - 1;
        for (int r = warp; r < R; r += NW) {
            float sa = 0.f, sb = 0.f;
#pragma unroll
            for (int half = 0; half < K / 128; ++half) {
                const int k = half * 128 + lane * 4;
                const float4 w = *reinterpret_cast<const float4*>(sW + r * K + k);
                const float4 a = *reinterpret_cast<const float4*>(in + k), b = *reinterpret_cast<const float4*>(in + f1 * K + k);
                sa = fmaf(w.x, a.x, sa); sa = fmaf(w.y, a.y, sa); sa = fmaf(w.z, a.z, sa); sa = fmaf(w.w, a.w, sa);
                sb = fmaf(w.x, b.x, sb); sb = fmaf(w.y, b.y, sb); sb = fmaf(w.z, b.z, sb); sb = fmaf(w.w, b.w, sb);
            }
            sa = warp_sum(sa);
            sb = warp_sum(sb);
            if (lane == 0) {
                out[o0 + r] = sa;
                if (f1 != 0) out[f1 * ldo + o0 + r] = sb;
            }
        }
        return;
    }
    const int nrg = (R + 1) / 2, nfg = (nb + 3) / 4;
    for (int task = warp; task < nrg * nfg; task += NW) {
        const int r0 = (task % nrg) * 2, f0 = (task / nrg) * 4;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
        for (int half = 0; half < K / 128; ++half) {
            const int k = half * 128 + lane * 4;
            float4 w[2], a[4];
#pragma unroll
            for (int r = 0; r < 2; ++r) w[r] = *reinterpret_cast<const float4*>(sW + min(r0 + r, R - 1) * K + k);
#pragma unroll
            for (int f = 0; f < 4; ++f) a[f] = *reinterpret_cast<const float4*>(in + min(f0 + f, nb - 1) * K + k);
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    float v = acc[r * 4 + f];
                    v = fmaf(w[r].x, a[f].x, v); v = fmaf(w[r].y, a[f].y, v); v = fmaf(w[r].z, a[f].z, v); v = fmaf(w[r].w, a[f].w, v);
                    acc[r * 4 + f] = v;
                }
        }
        // butterfly: after the three halving stages lane l holds the sum over its 4-lane-stride class of value (l >> 2) & 7
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool up = (lane & 16) != 0;
            const float send = up ? acc[i] : acc[i + 4], keep = up ? acc[i + 4] : acc[i];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const bool up = (lane & 8) != 0;
            const float send = up ? acc[i] : acc[i + 2], keep = up ? acc[i + 2] : acc[i];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        {
            const bool up = (lane & 4) != 0;
            const float send = up ? acc[0] : acc[1], keep = up ? acc[1] : acc[0];
            acc[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 2);
        acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);
        const int idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);     // which of the 8 values this lane ended with
        const int r = idx >> 2, f = idx & 3;
        if ((lane & 3) == 0 && f0 + f < nb && r0 + r < R) out[(f0 + f) * ldo + o0 + r0 + r] = acc[0];
    }
}

}  // namespace chain
}  // namespace wrnn
