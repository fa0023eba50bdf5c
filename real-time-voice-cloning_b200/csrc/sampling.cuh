// sampling.cuh -- fused sampling of the loop kernels (reference: fatchord_version.py:215-230 and
// vocoder/distribution.py:104-140; build-defined RAW rule = inverse CDF on one Philox uniform, see oracle/).
// Logits arrive as {value, tag} exchange words; every spin has a deadline (cycles) and an abort flag.
#pragma once
#include "common.cuh"

namespace wrnn {

// spin on one word
__device__ __forceinline__ bool wait_word(const unsigned long long* p, uint32_t tag, float& v, int* abort_flag, long long kSpinDeadline) {
    long long t0 = 0;
    int spins = 0;
    while (true) {
        unsigned long long w = ll_load(p);
        if (ll_tag(w) == tag) { v = ll_val(w); return true; }
        if (((++spins) & 63) == 0) {
            if (t0 == 0) t0 = clock64();
            if (clock64() - t0 > kSpinDeadline || ld_volatile_i32(abort_flag) != 0) return false;
        }
    }
}

// RAW: softmax + inverse CDF of one fold's logits row by one warp (rule: oracle sample_raw).
// Lane l owns classes [l*SEG, (l+1)*SEG).  Returns class index (warp-uniform) or -1 on timeout.
template <int SEG>
static __device__ __noinline__ int sample_raw_warp(const unsigned long long* __restrict__ row, uint32_t tag, float u, int* abort_flag,
                                           long long kSpinDeadline) {
    const int lane = threadIdx.x & 31;
    float l[SEG];
    bool ok = true;
    {
        const unsigned long long* src = row + lane * SEG;
        uint32_t pending = (SEG / 2 >= 32) ? 0xffffffffu : ((1u << (SEG / 2)) - 1u);
        long long t0 = 0;
        int spins = 0;
        while (pending) {
#pragma unroll
            for (int i = 0; i < SEG / 2; ++i)
                if ((pending >> i) & 1u) {
                    unsigned long long a, b;
                    ll_load2(src + 2 * i, a, b);
                    if (ll_tag(a) == tag && ll_tag(b) == tag) {
                        l[2 * i] = ll_val(a);
                        l[2 * i + 1] = ll_val(b);
                        pending &= ~(1u << i);
                    }
                }
            if (pending && ((++spins) & 63) == 0) {
                if (t0 == 0) t0 = clock64();
                if (clock64() - t0 > kSpinDeadline || ld_volatile_i32(abort_flag) != 0) { ok = false; break; }
            }
        }
    }
    if (!__all_sync(0xffffffffu, ok)) return -1;
    float m = l[0];
#pragma unroll
    for (int i = 1; i < SEG; ++i) m = fmaxf(m, l[i]);
    m = warp_max(m);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < SEG; ++i) { l[i] = expf(l[i] - m); s += l[i]; }
    const float total = warp_sum(s);
    float ls = 0.f;
#pragma unroll
    for (int i = 0; i < SEG; ++i) { l[i] = l[i] / total; ls += l[i]; }
    float incl = ls;   // inclusive scan over lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const float excl = incl - ls;
    const unsigned hit = __ballot_sync(0xffffffffu, incl >= u);
    int k = SEG * 32 - 1;
    if (hit) {
        const int L = __ffs(hit) - 1;
        int kk = SEG - 1;
        if (lane == L) {
            float c = excl;
#pragma unroll
            for (int i = 0; i < SEG; ++i) {
                c += l[i];
                if (c >= u) { kk = i; break; }
            }
            kk += L * SEG;
        }
        k = __shfl_sync(0xffffffffu, kk, L);
    }
    return k;
}

// MOL: vocoder/distribution.py:104-140 on one fold's 30 outputs by one warp.
static __device__ __noinline__ float sample_mol_warp(const unsigned long long* __restrict__ row, uint32_t tag, uint2 key, uint32_t step,
                                 const FoldDesc& fd, int* abort_flag, long long kSpinDeadline, bool& ok_out) {
    const int lane = threadIdx.x & 31;
    float lg = 0.f;
    bool ok = true;
    if (lane < 30) ok = wait_word(row + lane, tag, lg, abort_flag, kSpinDeadline);
    ok_out = __all_sync(0xffffffffu, ok);
    if (!ok_out) return 0.f;
    float score = -INFINITY;
    if (lane < 10) {
        uint4 r = philox4x32_10(make_uint4(step, (uint32_t)fd.fold, (uint32_t)fd.utt, (uint32_t)(lane >> 2)), key);
        float um = 1e-5f + u01(word_of(r, lane & 3)) * (1.0f - 2e-5f);
        score = lg - logf(-logf(um));
    }
    int idx = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float os = __shfl_xor_sync(0xffffffffu, score, o);
        int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (os > score || (os == score && oi < idx)) { score = os; idx = oi; }
    }
    const float mean = __shfl_sync(0xffffffffu, lg, 10 + idx);
    const float lsc = fmaxf(__shfl_sync(0xffffffffu, lg, 20 + idx), -32.23619130191664f);
    uint4 r2 = philox4x32_10(make_uint4(step, (uint32_t)fd.fold, (uint32_t)fd.utt, 2u), key);
    const float ul = 1e-5f + u01(r2.z) * (1.0f - 2e-5f);
    float x = mean + expf(lsc) * (logf(ul) - logf(1.0f - ul));
    return fminf(fmaxf(x, -1.0f), 1.0f);
}


}  // namespace wrnn
